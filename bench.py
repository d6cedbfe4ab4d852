#!/usr/bin/env python
"""Headline benchmark: acoustic frames/sec of one full train step (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload dblstm_ctc|las] [--impl b200|reference]

A "step" is Trainer.update on one synthetic minibatch: encoder forward, decoder/output layer, loss,
backward, gradient all-reduce (N > 1), clip + Adam.  One process per GPU (torchrun for N > 1), weak
scaling: every rank keeps the per-GPU batch of the named config.  Prints ONE JSON line on rank 0.

 * value   : frames/s with the batch already resident in HBM (CUDA events, max over ranks)
 * e2e     : frames/s through the same public call with pinned HOST inputs copied in every step and the
             loss read back every step
 * roofline: the dominant kernel of the timed region, timed live with CUDA events inside the library
 * cpu_baseline: the NumPy oracle port of the same step on the host cores, on a bounded sample
`--impl reference` runs only that CPU arm (the reference itself is Python-2/TF-1.8 and cannot run here;
see DESIGN.md), multi-threaded BLAS, same metric/config keys.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[2]: DBLSTM 5x512 + CTC, 128 x 1500 x 40 synthetic fbank, 29 labels
    'dblstm_ctc': dict(kind='ctc', B=128, T=1500, D=40, H=512, layers=5, V=29, L=150,
                       name='DBLSTM 5x512 + CTC, batch 128x1500x40 synthetic fbank (BASELINE configs[2])',
                       sample=dict(B=16, T=150, L=15)),
    # BASELINE.json configs[1]: Listener 3 pBLSTM-256 + 1 BLSTM + Speller 2x256 location-aware, 64x1000x40
    'las': dict(kind='las', B=64, T=1000, D=40, H=256, layers=3, V=30, U=100, dec_H=256, dec_layers=2,
                numfilt=10, filtersize=201,
                name='Listener 3xpBLSTM-256 + Speller 2x256 location_aware, batch 64x1000x40 (BASELINE configs[1])',
                sample=dict(B=8, T=200, U=20)),
    # BASELINE.json configs[0]: the reference's CPU-runnable plumbing case
    'dblstm_small': dict(kind='ctc', B=32, T=200, D=40, H=256, layers=2, V=29, L=20,
                         name='DBLSTM 2x256 + CTC, 32x200x40 (BASELINE configs[0])', sample=dict(B=32, T=200, L=20)),
}


def model_conf(w):
    import configparser
    c = configparser.ConfigParser()
    if w['kind'] == 'ctc':
        c.read_string('[io]\ninputs = features\noutputs = text\noutput_dims = %d\n'
                      '[encoder]\nencoder = dblstm\nnum_units = %d\nnum_layers = %d\ninput_noise = 0\ndropout = 1\n'
                      '[decoder]\ndecoder = dnn_decoder\nnum_layers = 0\n' % (w['V'] - 1, w['H'], w['layers']))
    else:
        c.read_string('[io]\ninputs = features\noutputs = text\noutput_dims = %d\n'
                      '[encoder]\nencoder = listener\nnum_units = %d\nnum_layers = %d\npyramid_steps = 2\n'
                      'input_noise = 0\ndropout = 1\n'
                      '[decoder]\ndecoder = speller\nnum_layers = %d\nnum_units = %d\ndropout = 1\n'
                      'attention = location_aware\nnumfilt = %d\nfiltersize = %d\nsample_prob = 0\n'
                      % (w['V'] - 1, w['H'], w['layers'], w['dec_layers'], w['dec_H'], w['numfilt'], w['filtersize']))
    return c


def trainer_conf(w):
    import configparser
    c = configparser.ConfigParser()
    c.read_string('[trainer]\ntrainer = standard\nloss = %s\ntrainlabels = 1\ntargets = text\nnum_epochs = 1\n'
                  'batch_size = %d\n' % ('CTC' if w['kind'] == 'ctc' else 'average_cross_entropy', w['B']))
    return c


def synth_batch(w, rank, B=None, T=None):
    """x ~ N(0,1) [B,T,40], full lengths, labels U{0..V-2}."""
    B = B or w['B']
    T = T or w['T']
    rng = np.random.default_rng(1234 + rank)
    x = rng.standard_normal((B, T, w['D']), dtype=np.float32)
    lens = np.full(B, T, np.int32)
    lab_rng = np.random.default_rng(99 + rank)
    if w['kind'] == 'ctc':
        L = max(1, T // 10)
        targets = lab_rng.integers(0, w['V'] - 1, size=(B, L)).astype(np.int32)
        tlen = np.full(B, L, np.int32)
    else:
        U = w['U'] if T == w['T'] else w['sample']['U']
        targets = lab_rng.integers(0, w['V'] - 1, size=(B, U)).astype(np.int32)
        targets[:, U - 1] = w['V'] - 1
        tlen = np.full(B, U, np.int32)
    return x, lens, targets, tlen


# ------------------------------------------------------------------------------------------------
# algorithmic work for the roofline object
# ------------------------------------------------------------------------------------------------
def ctc_step_work(w):
    B, T, H, V = w['B'], w['T'], w['H'], w['V']
    N = B * T
    gemm_flops = 0.0
    rec_bytes_f = rec_bytes_b = 0.0
    D = w['D']
    for l in range(w['layers']):
        per_dir = 2.0 * N * D * 4 * H
        gemm_flops += 2 * per_dir            # fwd x-projection, both directions
        gemm_flops += 2 * per_dir            # dKx
        if l > 0:
            gemm_flops += 2 * per_dir        # dX
        gemm_flops += 2 * 2.0 * N * H * 4 * H    # dKh
        rec_bytes_f += N * 4.0 * (8 * H + 2 * H + 2 * H)
        rec_bytes_b += N * 4.0 * 22 * H
        D = 2 * H
    gemm_flops += 3 * 2.0 * N * 2 * H * V
    return dict(gemm_flops=gemm_flops, rec_fwd_bytes=rec_bytes_f, rec_bwd_bytes=rec_bytes_b,
                rec_launches=w['layers'])


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the same train step
# ------------------------------------------------------------------------------------------------
def cpu_step_fn(w):
    """Returns (step() -> frames processed, description).  Oracle fp32, multi-threaded BLAS."""
    import oracle as O
    s = w['sample']
    rng = np.random.default_rng(7)
    if w['kind'] == 'ctc':
        x, lens, labels, ll = synth_batch(w, 0, s['B'], s['T'])
        layers, D = [], w['D']
        for _ in range(w['layers']):
            layers.append(O.init_blstm_params(rng, D, w['H']))
            D = 2 * w['H']
        lin = O.init_linear_params(rng, D, w['V'])

        def step():
            enc, _, caches = O.dblstm_fwd(x, lens, layers, np.float32)
            logits = O.linear_fwd(enc, lin, np.float32)
            loss, dlogits = O.ctc_loss_mean(logits, lens, labels, ll, dtype=np.float32)
            denc, glin = O.linear_bwd(enc, lin, dlogits)
            _, grads = O.dblstm_bwd(caches, denc)
            for p, g in zip(layers, grads):
                for k in p:
                    p[k], _, _ = O.tf_adam_clip(p[k], g[k], np.zeros_like(p[k]), np.zeros_like(p[k]), 1e-3, 1)
            return s['B'] * s['T']
        desc = 'full train step of the %s model on a %dx%dx%d slice (oracle port, numpy fp32)' % (
            'DBLSTM %dx%d+CTC' % (w['layers'], w['H']), s['B'], s['T'], w['D'])
        return step, desc
    else:
        x, lens, targets, tl = synth_batch(w, 0, s['B'], s['T'])
        layers, D = [], w['D']
        for _ in range(w['layers']):
            layers.append(O.init_blstm_params(rng, D, w['H']))
            D = 4 * w['H']
        layers.append(O.init_blstm_params(rng, D, w['H']))
        sp = O.init_speller_params(rng, w['V'], 2 * w['H'], w['dec_H'], w['dec_layers'], 'location_aware',
                                   w['numfilt'], w['filtersize'])

        def step():
            enc, elens, caches = O.listener_fwd(x, lens, layers, 2, np.float32)
            logits, ctx = O.speller_fwd(enc, elens, targets, tl, sp, 'location_aware', w['dec_layers'], np.float32)
            loss, dlogits = O.average_cross_entropy(logits, targets, tl, tl, np.float32)
            dmem, _ = O.speller_bwd(ctx, dlogits)
            O.listener_bwd(caches, dmem, 2)
            return s['B'] * s['T']
        desc = 'full train step of the LAS model on a %dx%dx%d slice, U=%d (oracle port, numpy fp32)' % (
            s['B'], s['T'], w['D'], s['U'])
        return step, desc


def ctc_loss_delta(trainer, w, dev):
    """The metric's second half ("CTC loss delta vs the CPU reference"): the CUDA model's CTC loss on a bounded slice
    of the workload against the oracle (fp64) fed the SAME weights and inputs.  Outside every timed region."""
    import torch
    import oracle as O
    s = w['sample']
    x, lens, labels, ll = synth_batch(w, 0, s['B'], s['T'])
    params = trainer.model.store.to_numpy()
    with torch.no_grad():
        t = lambda a: torch.from_numpy(a).to(dev)
        batch = ({'features': t(x)}, {'features': t(lens)}, {'text': t(labels)}, {'text': t(ll)})
        logits, logit_len = trainer.model(batch[0], batch[1], batch[2], batch[3], True)
        cuda_loss = float(trainer.loss_fn(batch[2], logits, logit_len, batch[3]))
    layers = []
    for l in range(w['layers']):
        base = 'DBLSTM/features/layer%d/bidirectional_rnn/%%s/layer_norm_basic_lstm_cell/%%s' % l
        layers.append({'%s_%s' % (d, k): params[base % (d, k)] for d in ('fw', 'bw') for k in ('kernel', 'bias')})
    lin = {'weights': params['DNNDecoder/text/outlayer/weights'], 'biases': params['DNNDecoder/text/outlayer/biases']}
    enc, _, _ = O.dblstm_fwd(x, lens, layers)
    cpu_loss, _ = O.ctc_loss_mean(O.linear_fwd(enc, lin), lens, labels, ll)
    return {'value': float(abs(cuda_loss - cpu_loss) / abs(cpu_loss)), 'cuda': cuda_loss, 'cpu_fp64': float(cpu_loss),
            'tolerance': 1e-4, 'sample': '%dx%dx%d slice, the weights after the timed steps' % (s['B'], s['T'], w['D'])}


def run_cpu(w, steps, warmup, budget_s=None):
    step, desc = cpu_step_fn(w)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    frames = 0
    n = 0
    for _ in range(steps):
        frames += step()
        n += 1
        if budget_s and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return frames / dt, dt / n, desc, n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='dblstm_ctc', choices=sorted(WORKLOADS))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if os.environ.get('NABU_BENCH_T'):       # profiling aid only (ncu captures); never a bench value
        w['T'] = int(os.environ['NABU_BENCH_T'])
        w['name'] += ' [T overridden to %d for profiling]' % w['T']
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    cores = os.cpu_count() or 1
    config = {'workload': w['name'], 'per_gpu_batch': w['B'], 'global_batch': w['B'] * world, 'frames_per_utt': w['T'],
              'parallelism': 'dp%d' % world, 'precision_mode': 'fp32 parity (input_noise=0, dropout=1)',
              'l2': 'per-step working set (saved activations, >20 GB) is far larger than the 126 MB L2'}
    base = {'metric': 'acoustic frames/sec (train step)', 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'config': config}

    if args.impl == 'reference':
        if rank != 0:
            return
        try:
            import torch
            torch.set_num_threads(cores)
        except Exception:
            pass
        fps, sec, desc, n = run_cpu(w, args.steps, min(args.warmup, 1))
        out = dict(base)
        out.update({'impl': 'reference', 'value': fps, 'steps': n, 'ms_per_step': sec * 1e3,
                    'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
                                     'sample': desc},
                    'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                    'gpu_launches': 0})
        print(json.dumps(out))
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the product path has no CPU fallback)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    from nabu_b200 import lib as L
    from nabu_b200.neuralnetworks.trainers import trainer_factory
    lib = L.load()
    trainer = trainer_factory.factory('standard')(trainer_conf(w), None, model_conf(w), None, None, None, rank,
                                                  device=dev, seed=7)
    trainer.num_steps = 10000
    trainer.model.build({'features': w['D']}, dev)

    x, lens, targets, tlen = synth_batch(w, rank)
    hx = torch.from_numpy(x).pin_memory()
    hl, ht, htl = (torch.from_numpy(a).pin_memory() for a in (lens, targets, tlen))

    def to_dev():
        return ({'features': hx.to(dev, non_blocking=True)}, {'features': hl.to(dev, non_blocking=True)},
                {'text': ht.to(dev, non_blocking=True)}, {'text': htl.to(dev, non_blocking=True)})

    batch = to_dev()
    frames_per_step = int(lens.sum())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def rank_max(v):
        if world == 1:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(args.warmup, 3)):
        loss, _ = trainer.update(*batch)
    barrier()

    # ---- timed region 1: inputs resident in HBM -----------------------------------------------------
    clocks = ClockSampler(local_rank)
    clocks.start()
    lib.nabu_profile_enable(1)
    launches0 = lib.nabu_kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss, _ = trainer.update(*batch)
    e1.record()
    barrier()
    ms = rank_max(e0.elapsed_time(e1))
    launches = lib.nabu_kernel_launches() - launches0
    lib.nabu_profile_enable(0)
    buf = (b' ' * 65536)
    import ctypes
    cbuf = ctypes.create_string_buffer(65536)
    lib.nabu_profile_collect(cbuf, 65536)
    prof = json.loads(cbuf.value.decode())
    clk = clocks.stop()
    value = frames_per_step * world * args.steps / (ms * 1e-3)

    # ---- timed region 2: end to end from pinned host buffers ----------------------------------------
    barrier()
    t0 = time.perf_counter()
    last = None
    for _ in range(args.steps):
        b = to_dev()
        loss, _ = trainer.update(*b)
        last = float(loss)           # device -> host read of the step's result
    torch.cuda.synchronize()
    dt = rank_max(time.perf_counter() - t0)
    e2e = frames_per_step * world * args.steps / dt
    h2d = hx.numel() * 4 + hl.numel() * 4 + ht.numel() * 4 + htl.numel() * 4

    # ---- roofline of the dominant kernel --------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm_peak = peaks.get('hbm_gbs', 6650.0)
    tf_peak = peaks.get('bf16_tflops_sustained', 1400.0)
    peak_src = 'measured' if peaks else 'fallback'
    top = max(prof.items(), key=lambda kv: kv[1][1]) if prof else (None, [0, 0.0])
    roofline = None
    step_ms = ms / args.steps
    shares = {k: round(v[1] / max(ms, 1e-9), 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}
    if w['kind'] == 'ctc' and top[0] is not None:
        work = ctc_step_work(w)
        name, (cnt, tot) = top
        if name.startswith('sgemm') or name.startswith('gemm'):
            gemm_ms = sum(v[1] for k, v in prof.items() if k.startswith('sgemm') or k.startswith('gemm'))
            ach = work['gemm_flops'] * args.steps / (gemm_ms * 1e-3) / 1e12
            roofline = {'kernel': 'dense contractions (%s ...)' % name, 'bound': 'tensor', 'achieved': ach,
                        'peak': tf_peak, 'unit': 'TFLOP/s', 'frac': ach / tf_peak, 'traffic': None,
                        'peak_source': peak_src + ' bf16 sustained; fp32-parity contractions'}
        else:
            key = 'rec_fwd_bytes' if 'fwd' in name else 'rec_bwd_bytes'
            per_launch = work[key] / work['rec_launches']
            ach = per_launch / (tot / cnt * 1e-3) / 1e9
            # DRAM bytes per frame of one launch from the ncu --set full capture (profiles/r1d_ncu_full.md, T=96 launch:
            # dram__bytes_read.sum + dram__bytes_write.sum over 12 288 frames), scaled to this launch's frames
            ncu_bytes_per_frame = 37.4e3 if 'fwd' in name else 38.0e3
            roofline = {'kernel': name, 'bound': 'hbm', 'achieved': ach, 'peak': hbm_peak, 'unit': 'GB/s',
                        'frac': ach / hbm_peak, 'traffic': ncu_bytes_per_frame * w['B'] * w['T'],
                        'traffic_unit': 'bytes per launch (ncu capture at T=96, scaled by frames)',
                        'algorithmic_bytes_per_launch': per_launch, 'peak_source': peak_src,
                        'serial_steps_per_launch': w['T'], 'us_per_serial_step': tot / cnt * 1e3 / w['T'],
                        'note': 'latency-bound serial scan: T dependent time steps per launch; see DESIGN.md section 6'}
    elif top[0] is not None:
        name, (cnt, tot) = top
        roofline = {'kernel': name, 'bound': 'hbm', 'achieved': None, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': None,
                    'traffic': None, 'peak_source': peak_src}
    roofline_gemm = None
    if w['kind'] == 'ctc' and prof:
        # second view: the dense contractions (tensor-bound).  fp32-equivalent FLOPs / summed GEMM kernel time; each
        # product costs 3 fp16 MMAs, so the tensor pipe does 3x this rate.  Under the deferred-weight-gradient overlap the
        # TN GEMMs share the GPU with the backward recurrence, which lengthens their event-timed duration.
        work = ctc_step_work(w)
        gemm_ms = sum(v[1] for k, v in prof.items() if k.startswith('gemm_h2') or k.startswith('gemm_tc'))
        if gemm_ms > 0:
            ach = work['gemm_flops'] * args.steps / (gemm_ms * 1e-3) / 1e12
            roofline_gemm = {'kernel': 'gemm_h2 (fp16 hi/lo split, 3 MMAs per product)', 'bound': 'tensor',
                             'achieved': ach, 'peak': tf_peak, 'unit': 'TFLOP/s', 'frac': ach / tf_peak,
                             'mma_rate_tflops': 3 * ach, 'mma_frac': 3 * ach / tf_peak, 'traffic': None,
                             'peak_source': peak_src + ' bf16 sustained (fp16 MMA runs at the same rate)'}
    if roofline is not None:
        roofline['avg_launch_ms'] = top[1][1] / max(top[1][0], 1)
        roofline['kernel_time_shares'] = shares
        roofline['shares_note'] = ('per-kernel CUDA-event time / step time; the weight-gradient GEMMs run on a side '
                                   'stream under the backward recurrences, so shares can sum to more than 1')

    out = dict(base)
    out.update({'value': value, 'ms_per_step': step_ms, 'loss': last, 'clocks': clk, 'gpu_launches': int(launches),
                'e2e': {'value': e2e, 'unit': 'frames/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4},
                'roofline': roofline, 'roofline_gemm': roofline_gemm})
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(cores)
        fps, sec, desc, n = run_cpu(w, 3, 1, budget_s=30)
        out['cpu_baseline'] = {'value': fps, 'unit': 'frames/s', 'cores': cores, 'kind': 'port', 'sample': desc}
        if w['kind'] == 'ctc':
            try:
                out['ctc_loss_delta_vs_cpu'] = ctc_loss_delta(trainer, w, dev)
            except Exception as e:          # a reported extra: it must never cost the bench line
                out['ctc_loss_delta_vs_cpu'] = {'value': None, 'error': '%s: %s' % (type(e).__name__, e)}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
