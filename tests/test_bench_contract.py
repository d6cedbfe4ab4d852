"""bench.py's contract on the CPU: the algorithmic-work model behind `roofline.achieved`, the
synthetic inputs, and the JSON line of the `--impl reference` arm (the only arm that runs without a GPU)."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_work_model_matches_the_survey_figures():
    w = bench.WORKLOADS['dblstm_ctc']
    assert (w['B'], w['T'], w['D'], w['H'], w['layers'], w['V']) == (128, 1500, 40, 512, 5, 29)      # BASELINE configs[2]
    work = bench.ctc_step_work(w)
    frames = w['B'] * w['T']
    # recurrent scan, per BLSTM layer and valid frame: forward 4*(8H + 2H + 2H) bytes, backward 4*22H bytes
    assert work['rec_fwd_bytes'] == frames * w['layers'] * 4 * 12 * w['H']
    assert work['rec_bwd_bytes'] == frames * w['layers'] * 4 * 22 * w['H']
    assert work['rec_launches'] == w['layers']
    # dense contractions of one train step: SURVEY 8d quotes 164.7 MFLOP per frame in total for cfg-3, of which the
    # recurrent matmuls (3 x 16 H^2 per layer) are not GEMM launches; what is left must be what the model counts
    per_frame = work['gemm_flops'] / frames
    recurrent = 3 * 16 * w['H'] ** 2 * w['layers'] - 2 * 2 * w['H'] * 4 * w['H'] * w['layers']   # dKh IS a GEMM launch
    assert abs((per_frame + recurrent) / 164.7e6 - 1) < 0.02, per_frame
    las = bench.WORKLOADS['las']
    assert (las['B'], las['T'], las['D']) == (64, 1000, 40)                                           # configs[1]


def test_synthetic_batch_is_seeded_and_shaped():
    w = bench.WORKLOADS['dblstm_ctc']
    a = bench.synth_batch(w, 0, 4, 50)
    b = bench.synth_batch(w, 0, 4, 50)
    c = bench.synth_batch(w, 1, 4, 50)
    x, lens, labels, ll = a
    assert x.shape == (4, 50, 40) and x.dtype == np.float32 and lens.dtype == np.int32 and labels.dtype == np.int32
    assert all(np.array_equal(p, q) for p, q in zip(a, b)) and not np.array_equal(a[0], c[0])      # seed 1234 + rank
    assert labels.max() <= w['V'] - 2 and (ll <= lens // 10).all()                                 # blank never a label


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, RANK='0', WORLD_SIZE='1')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                          '--warmup', '0'], capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().split('\n') if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'impl', 'cpu_baseline', 'e2e'):
        assert key in d, key
    assert d['impl'] == 'reference' and d['unit'] == 'frames/s' and d['vs_baseline'] is None and d['value'] > 0
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['sample']
    assert d['e2e'] == {'value': d['value'], 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and 'model' not in d['config']
    # ranks other than 0 do no work and print nothing
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2'],
                         capture_output=True, text=True, env=dict(env, RANK='1', WORLD_SIZE='2'), timeout=120, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ''
