#!/bin/bash
# 2-GPU box: scripts.train under torchrun (data from nabu directories, sharded per rank) vs the single-GPU run
mkdir -p gpurun_out
R=/tmp/dpcheck; rm -rf $R
{
python tools/dp_train_check.py prepare $R
timeout 200 python -m nabu_b200.scripts.train --expdir $R/one/exp | grep -c "step" 
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  -m nabu_b200.scripts.train --expdir $R/two/exp | grep -E "WORKER [01]: step (0|7)/"
python tools/dp_train_check.py compare $R
} > gpurun_out/dp_check.log 2>&1
tail -15 gpurun_out/dp_check.log
