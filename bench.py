#!/usr/bin/env python
"""Headline benchmark: acoustic frames/sec of one full train step (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl b200|reference]

A "step" is Trainer.update on one synthetic minibatch: encoder forward, decoder / output layer, loss, backward,
gradient all-reduce (N > 1), clip + Adam.  One process per GPU (torchrun for N > 1).  Rank 0 prints ONE JSON line.

What the line measures (default workload = BASELINE configs[2], DBLSTM 5x512 + CTC, 128 x 1500 x 40):
 * value / ms_per_step : WEAK scaling -- every GPU runs the config's own 128-utterance minibatch (global batch 128 N);
                         inputs resident in HBM, CUDA events, max over ranks, per-launch profiling OFF
 * strong              : (N > 1) the SURVEY 8e split of ONE 128-utterance minibatch, rank r takes utterances r::N
                         (128/N per GPU, global batch 128), same timing rules -- the serial-latency-bound case
 * allreduce           : (N > 1) the step's one collective timed alone on the flat gradient buffer: bytes, ms, bus GB/s
 * e2e                 : the weak step through the same public call with pinned HOST inputs copied in and the loss read
                         back every step
 * roofline            : the dominant kernel, timed in a SEPARATE pass with CUDA events around every launch inside the
                         library; `traffic` = dram bytes of that kernel from the committed ncu capture of this command
                         (profiles/r2_traffic.json), null when no capture of this configuration exists
 * cpu_baseline        : the NumPy oracle port of the same step on the host cores (full batch, T truncated)
 * ctc_loss_delta_vs_cpu : CUDA model vs fp64 oracle CTC loss on utterances of the full T = 1500 length
 * las / decode        : (N = 1, default workload) configs[1] train step and configs[3] decode objects, same keys
`--impl reference` runs only the CPU arm (the reference is Python-2 / TF-1.8 and cannot run here).
"""
import os
import sys

if '--impl' in sys.argv and 'reference' in sys.argv:
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm uses every host core at every N.
    _n = str(os.cpu_count() or 1)
    for _k in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[_k] = _n

import argparse
import ctypes
import json
import subprocess
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[2]: DBLSTM 5x512 + CTC, 128 x 1500 x 40 synthetic fbank, 29 labels
    'dblstm_ctc': dict(kind='ctc', B=128, T=1500, D=40, H=512, layers=5, V=29,
                       name='DBLSTM 5x512 + CTC, batch 128x1500x40 synthetic fbank (BASELINE configs[2])',
                       sample=dict(B=128, T=60)),
    # BASELINE.json configs[1]: Listener 3 pBLSTM-256 + 1 BLSTM + Speller 2x256 location-aware, 64x1000x40
    'las': dict(kind='las', B=64, T=1000, D=40, H=256, layers=3, V=30, U=100, dec_H=256, dec_layers=2,
                numfilt=10, filtersize=201,
                name='Listener 3xpBLSTM-256 + Speller 2x256 location_aware, batch 64x1000x40 (BASELINE configs[1])',
                sample=dict(B=64, T=120, U=12)),
    # BASELINE.json configs[0]: the reference's CPU-runnable plumbing case
    'dblstm_small': dict(kind='ctc', B=32, T=200, D=40, H=256, layers=2, V=29,
                         name='DBLSTM 2x256 + CTC, 32x200x40 (BASELINE configs[0])', sample=dict(B=32, T=200)),
    # BASELINE.json configs[4]: DBLSTM 6x1024 + CTC, 256 x 2000 over 8 GPUs = 32 utterances per GPU.  The saved
    # activations of the whole 256-utterance batch (150 GB) do not fit one GPU next to the workspace, so every GPU runs
    # the 8-GPU shard at any N.
    'dblstm_1024': dict(kind='ctc', B=32, T=2000, D=40, H=1024, layers=6, V=29,
                        name='DBLSTM 6x1024 + CTC, 32x2000x40 per GPU = the 8-GPU shard of batch 256x2000 (BASELINE configs[4])',
                        sample=dict(B=32, T=40)),
    # BASELINE.json configs[3]: LAS BeamSearchDecoder decode, beam 16, batch 32x1000x40 / CTCDecoder on the cfg-3 model
    'las_decode': dict(kind='las_decode', B=32, T=1000, D=40, H=256, layers=3, V=30, U=100, dec_H=256, dec_layers=2,
                       numfilt=10, filtersize=201, beam=16, max_steps=100,
                       name='LAS BeamSearchDecoder, beam 16, max_steps 100, batch 32x1000x40 (BASELINE configs[3])'),
    'ctc_decode': dict(kind='ctc_decode', B=32, T=1000, D=40, H=512, layers=5, V=29,
                       name='DBLSTM 5x512 + CTCDecoder (beam 100, merge_repeated), batch 32x1000x40 (BASELINE configs[3])'),
}


def make_conf(text):
    import configparser
    c = configparser.ConfigParser()
    c.read_string(text)
    return c


def model_conf(w):
    if w['kind'] in ('ctc', 'ctc_decode'):
        return make_conf('[io]\ninputs = features\noutputs = text\noutput_dims = %d\n'
                         '[encoder]\nencoder = dblstm\nnum_units = %d\nnum_layers = %d\ninput_noise = 0\ndropout = 1\n'
                         '[decoder]\ndecoder = dnn_decoder\nnum_layers = 0\n' % (w['V'] - 1, w['H'], w['layers']))
    return make_conf('[io]\ninputs = features\noutputs = text\noutput_dims = %d\n'
                     '[encoder]\nencoder = listener\nnum_units = %d\nnum_layers = %d\npyramid_steps = 2\n'
                     'input_noise = 0\ndropout = 1\n'
                     '[decoder]\ndecoder = speller\nnum_layers = %d\nnum_units = %d\ndropout = 1\n'
                     'attention = location_aware\nnumfilt = %d\nfiltersize = %d\nsample_prob = 0\n'
                     % (w['V'] - 1, w['H'], w['layers'], w['dec_layers'], w['dec_H'], w['numfilt'], w['filtersize']))


def trainer_conf(w):
    return make_conf('[trainer]\ntrainer = standard\nloss = %s\ntrainlabels = 1\ntargets = text\nnum_epochs = 1\n'
                     'batch_size = %d\n' % ('CTC' if w['kind'] == 'ctc' else 'average_cross_entropy', w['B']))


def synth_batch(w, seed_off, B=None, T=None, U=None):
    """x ~ N(0,1) [B,T,40], full lengths, labels U{0..V-2}."""
    B = B or w['B']
    T = T or w['T']
    rng = np.random.default_rng(1234 + seed_off)
    x = rng.standard_normal((B, T, w['D']), dtype=np.float32)
    lens = np.full(B, T, np.int32)
    lab_rng = np.random.default_rng(99 + seed_off)
    if w['kind'] in ('ctc', 'ctc_decode'):
        L = max(1, T // 10)
        targets = lab_rng.integers(0, w['V'] - 1, size=(B, L)).astype(np.int32)
        tlen = np.full(B, L, np.int32)
    else:
        U = U or w['U']
        targets = lab_rng.integers(0, w['V'] - 1, size=(B, U)).astype(np.int32)
        targets[:, U - 1] = w['V'] - 1
        tlen = np.full(B, U, np.int32)
    return x, lens, targets, tlen


# ------------------------------------------------------------------------------------------------
# algorithmic work for the roofline objects
# ------------------------------------------------------------------------------------------------
def blstm_stack_work(B, T_of_layer, D0, H, n_layers, widen):
    """FLOPs of the dense contractions (train) and scan bytes of a stack of BLSTM layers; layer l runs T_of_layer[l]
    frames on an input of D0 (l = 0) or widen * 2H features."""
    gemm_flops, rec_f, rec_b = 0.0, [], []
    D = D0
    for l in range(n_layers):
        N = B * T_of_layer[l]
        per_dir = 2.0 * N * D * 4 * H
        gemm_flops += 2 * per_dir + 2 * per_dir + (2 * per_dir if l > 0 else 0)     # x-projection, dKx, dX
        gemm_flops += 2 * 2.0 * N * H * 4 * H                                          # dKh
        rec_f.append(N * 4.0 * (8 * H + 2 * H + 2 * H))       # read Gx, write h, write c
        rec_b.append(N * 4.0 * 22 * H)
        D = widen * 2 * H
    return gemm_flops, rec_f, rec_b


def ctc_step_work(w, B=None):
    B = B or w['B']
    g, f, b = blstm_stack_work(B, [w['T']] * w['layers'], w['D'], w['H'], w['layers'], 1)
    g += 3 * 2.0 * B * w['T'] * 2 * w['H'] * w['V']
    return dict(gemm_flops=g, rec_fwd_bytes=sum(f), rec_bwd_bytes=sum(b), rec_launches=w['layers'])


def las_attention_bytes(w, B=None):
    """one decoder step streams keys [B,T',A] + values [B,T',2H] once (24.6 MB at cfg-2)."""
    B = B or w['B']
    Tm = w['T'] // (2 ** w['layers'])
    return 4.0 * B * Tm * (w['dec_H'] + 2 * w['H'])


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
# CPU arm: the oracle port of the same train step (full batch, T truncated)
# ------------------------------------------------------------------------------------------------
def cpu_threads():
    cores = os.cpu_count() or 1
    try:
        import torch
        torch.set_num_threads(cores)
    except Exception:
        pass
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=cores)          # numpy's BLAS, whatever OMP_NUM_THREADS said at import time
    except Exception:
        pass
    return cores


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
        slots = {}

        def adam(p, g, key):
            m, v = slots.setdefault(key, (np.zeros_like(p), np.zeros_like(p)))
            th, m2, v2 = O.tf_adam_clip(p, g, m, v, 1e-3, 1)
            slots[key] = (m2, v2)
            return th

        def step():
            enc, _, caches = O.dblstm_fwd(x, lens, layers, np.float32)
            logits = O.linear_fwd(enc, lin, np.float32)
            loss, dlogits = O.ctc_loss_mean(logits, lens, labels, ll, dtype=np.float32)
            denc, glin = O.linear_bwd(enc, lin, dlogits)
            _, grads = O.dblstm_bwd(caches, denc)
            for i, (p, g) in enumerate(zip(layers, grads)):
                for k in p:
                    p[k] = adam(p[k], g[k], (i, k))
            for k in lin:
                lin[k] = adam(lin[k], glin[k], ('lin', k))
            return s['B'] * s['T']
        desc = ('full train step (fwd, CTC, bwd, clip+Adam of every variable) of DBLSTM %dx%d+CTC at the full batch B=%d '
                'with T truncated to %d of %d frames (oracle port, numpy fp32, time is linear in T)'
                % (w['layers'], w['H'], s['B'], s['T'], w['T']))
        return step, desc
    x, lens, targets, tl = synth_batch(w, 0, s['B'], s['T'], s['U'])
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
        dmem, gsp = O.speller_bwd(ctx, dlogits)
        _, glayers = O.listener_bwd(caches, dmem, 2)
        for p, g in zip(layers, glayers):
            for k in p:
                p[k], _, _ = O.tf_adam_clip(p[k], g[k], np.zeros_like(p[k]), np.zeros_like(p[k]), 1e-3, 1)
        for k in gsp:
            sp[k], _, _ = O.tf_adam_clip(sp[k], gsp[k].astype(np.float32), np.zeros_like(sp[k]), np.zeros_like(sp[k]), 1e-3, 1)
        return s['B'] * s['T']
    desc = ('full train step of the LAS model at the full batch B=%d with T truncated to %d of %d frames and U to %d of '
            '%d targets (oracle port, numpy fp32)' % (s['B'], s['T'], w['T'], s['U'], w['U']))
    return step, desc


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


def dblstm_oracle_params(params, n_layers):
    layers = []
    for l in range(n_layers):
        base = 'DBLSTM/features/layer%d/bidirectional_rnn/%%s/layer_norm_basic_lstm_cell/%%s' % l
        layers.append({'%s_%s' % (d, k): params[base % (d, k)] for d in ('fw', 'bw') for k in ('kernel', 'bias')})
    lin = {'weights': params['DNNDecoder/text/outlayer/weights'], 'biases': params['DNNDecoder/text/outlayer/biases']}
    return layers, lin


def ctc_loss_delta(trainer, w, dev, n_utt=4):
    """The metric's second half ("CTC loss delta vs the CPU reference"): the CUDA model's per-utterance CTC loss at the
    workload's FULL length T against the fp64 oracle fed the same weights and inputs, on `n_utt` utterances (the oracle
    needs ~1.7 s per utterance and layer-thousand-frames).  Outside every timed region."""
    import torch
    import oracle as O
    from nabu_b200 import engine
    x, lens, labels, ll = synth_batch(w, 0, n_utt, w['T'])
    params = trainer.model.store.to_numpy()
    with torch.no_grad():
        t = lambda a: torch.from_numpy(a).to(dev)
        logits, logit_len = trainer.model({'features': t(x)}, {'features': t(lens)}, None, None, False)
        per_utt, _ = engine.ctc_loss_per_utt(logits['text'], t(lens), t(labels), t(ll))
        cuda = per_utt.cpu().numpy().astype(np.float64)
    layers, lin = dblstm_oracle_params(params, w['layers'])
    enc, _, _ = O.dblstm_fwd(x, lens, layers)
    cpu, _ = O.ctc_loss_and_grad(O.linear_fwd(enc, lin), lens, labels, ll)
    return {'value': float(np.abs(cuda / cpu - 1).max()), 'cuda': [float(v) for v in cuda], 'cpu_fp64': [float(v) for v in cpu],
            'tolerance': 1e-4, 'sample': '%d utterances x %d frames (the full length), the weights the timed steps left; '
            'worst per-utterance relative difference' % (n_utt, w['T'])}


# ------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------
class Dist(object):
    def __init__(self):
        self.rank = int(os.environ.get('RANK', '0'))
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))

    def barrier(self):
        import torch
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def rank_max(self, v, dev):
        if self.world == 1:
            return v
        import torch
        import torch.distributed as dist
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())


def make_trainer(w, dev, rank, seed=7):
    from nabu_b200.neuralnetworks.trainers import trainer_factory
    trainer = trainer_factory.factory('standard')(trainer_conf(w), None, model_conf(w), None, None, None, rank,
                                                  device=dev, seed=seed)
    trainer.num_steps = 10000
    trainer.model.build({'features': w['D']}, dev)
    return trainer


class HostBatch(object):
    """One synthetic minibatch in pinned host memory + its device copy."""

    def __init__(self, arrays, dev):
        import torch
        self.dev = dev
        self.host = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in arrays]
        self.frames = int(arrays[1].sum())
        self.bytes = sum(t.numel() * t.element_size() for t in self.host)

    def to_dev(self):
        hx, hl, ht, htl = [t.to(self.dev, non_blocking=True) for t in self.host]
        return ({'features': hx}, {'features': hl}, {'text': ht}, {'text': htl})


def timed_steps(trainer, batch, steps, dd, dev):
    """EXACTLY `steps` updates between two events, barrier + synchronize on both sides, max over ranks -> ms."""
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dd.barrier()
    e0.record()
    for _ in range(steps):
        loss, _ = trainer.update(*batch)
    e1.record()
    dd.barrier()
    return dd.rank_max(e0.elapsed_time(e1), dev), loss


def profile_pass(lib, trainer, batch, steps, dd):
    """A separate pass with CUDA events around every launch inside the library -> {kernel: [launches, total ms]}."""
    import torch
    dd.barrier()
    lib.nabu_profile_enable(1)
    t0 = time.perf_counter()
    for _ in range(steps):
        trainer.update(*batch)
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    lib.nabu_profile_enable(0)
    cbuf = ctypes.create_string_buffer(1 << 17)
    lib.nabu_profile_collect(cbuf, 1 << 17)
    return json.loads(cbuf.value.decode()), wall_ms


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        return p.get('hbm_gbs', 6650.0), p.get('bf16_tflops_sustained', 1400.0), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 1400.0, 'fallback (B200_PROFILING.md)'


def ncu_traffic(kernel, workload_key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu capture of this very
    command (profiles/r2_traffic.json, written by tools/ncu_traffic.py); None when there is no capture."""
    try:
        table = json.load(open(os.path.join(ROOT, 'profiles', 'r2_traffic.json')))
        e = table.get(workload_key, {}).get(kernel)
        return (e['dram_bytes_per_launch'], e.get('source')) if e else (None, None)
    except Exception:
        return None, None


def rooflines(w, wkey, prof, steps, step_ms):
    hbm_peak, tf_peak, src = peaks()
    if not prof:
        return None, None
    shares = {k: round(v[1] / steps / max(step_ms, 1e-9), 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}
    top, (cnt, tot) = max(prof.items(), key=lambda kv: kv[1][1])
    roofline = roofline_gemm = None
    if w['kind'] == 'ctc':
        work = ctc_step_work(w)
        rec = [k for k in prof if k.startswith('blstm_rec')]
        name = max(rec, key=lambda k: prof[k][1]) if rec else top
        cnt, tot = prof[name]
        key = 'rec_fwd_bytes' if 'fwd' in name else 'rec_bwd_bytes'
        per_launch = work[key] / work['rec_launches']
        avg_ms = tot / max(cnt, 1)
        ach = per_launch / (avg_ms * 1e-3) / 1e9
        traffic, tsrc = ncu_traffic(name, wkey)
        roofline = {'kernel': name, 'bound': 'hbm', 'achieved': ach, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': ach / hbm_peak,
                    'traffic': traffic, 'traffic_source': tsrc, 'algorithmic_bytes_per_launch': per_launch,
                    'peak_source': src, 'avg_launch_ms': avg_ms, 'serial_steps_per_launch': w['T'],
                    'us_per_serial_step': avg_ms * 1e3 / w['T'],
                    'note': 'latency-bound serial scan: T dependent time steps per launch (DESIGN.md section 6); timed in a '
                            'separate profiling pass, not in the pass `value` comes from'}
        us = {}
        for k in rec:
            us[k] = round(prof[k][1] / max(prof[k][0], 1) * 1e3 / w['T'], 3)
        roofline['us_per_serial_step_all'] = us
        gemm_ms = sum(v[1] for k, v in prof.items() if k.startswith('gemm_h2') or k.startswith('gemm_tc'))
        if gemm_ms > 0:
            ach = work['gemm_flops'] * steps / (gemm_ms * 1e-3) / 1e12
            roofline_gemm = {'kernel': 'gemm_h2 (fp16 hi/lo split, 3 MMAs per product)', 'bound': 'tensor', 'achieved': ach,
                             'peak': tf_peak, 'unit': 'TFLOP/s', 'frac': ach / tf_peak, 'mma_rate_tflops': 3 * ach,
                             'mma_frac': 3 * ach / tf_peak, 'traffic': None,
                             'peak_source': src + ' bf16 sustained (fp16 MMA runs at the same rate)'}
            try:        # the committed ncu capture of the x-projection launches of this command (None when there is none)
                e = json.load(open(os.path.join(ROOT, 'profiles', 'r2_traffic.json'))).get(wkey, {}).get('gemm_h2_nn')
                if e:
                    roofline_gemm['traffic'] = e['dram_bytes_per_launch']
                    roofline_gemm['ncu'] = {k: e[k] for k in ('launch', 'tensor_pipe_active_pct', 'sm_mhz_under_load', 'source')}
            except Exception:
                pass
    else:
        # LAS: the attention step (keys + values streamed once per decoder step)
        att = [k for k in prof if k.startswith('dec_attn_step') or k == 'dec_attn_fwd_persist']
        name = att[0] if att else top
        cnt, tot = prof[name]
        per_launch = las_attention_bytes(w) * (w['U'] if 'persist' in name else 1)
        avg_ms = tot / max(cnt, 1)
        ach = per_launch / (avg_ms * 1e-3) / 1e9
        traffic, tsrc = ncu_traffic(name, wkey)
        roofline = {'kernel': name, 'bound': 'hbm', 'achieved': ach, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': ach / hbm_peak,
                    'traffic': traffic, 'traffic_source': tsrc, 'algorithmic_bytes_per_launch': per_launch,
                    'peak_source': src, 'avg_launch_ms': avg_ms, 'launches_per_step': cnt / steps,
                    'note': 'keys [B,T\',A] + values [B,T\',2H] read once per decoder step; the 24.6 MB fit the 126 MB L2, so '
                            'after the first step the stream comes from L2, not HBM'}
    roofline['kernel_time_shares'] = shares
    roofline['shares_note'] = ('per-kernel CUDA-event time / step time from the profiling pass; the weight-gradient GEMMs run '
                               'on a side stream under the backward recurrences, so shares can sum to more than 1')
    return roofline, roofline_gemm


def bench_train(w, wkey, args, dd, dev, lib, with_cpu, extras=True):
    """Weak line (+ strong, allreduce at N > 1) of a train workload.  Returns the dict of the JSON line's fields."""
    import torch
    trainer = make_trainer(w, dev, dd.rank)
    hb = HostBatch(synth_batch(w, dd.rank), dev)
    batch = hb.to_dev()
    W = max(args.warmup, 3)
    for _ in range(W):
        trainer.update(*batch)
    dd.barrier()
    out = {}
    # ---- timed region 1: inputs resident in HBM, profiling off ------------------------------------------------------
    clocks = ClockSampler(dd.local_rank)
    clocks.start()
    launches0 = lib.nabu_kernel_launches()
    ms, loss = timed_steps(trainer, batch, args.steps, dd, dev)
    launches = lib.nabu_kernel_launches() - launches0
    out['clocks'] = clocks.stop()
    out['value'] = hb.frames * dd.world * args.steps / (ms * 1e-3)
    out['ms_per_step'] = ms / args.steps
    out['gpu_launches'] = int(launches)
    # ---- timed region 2: end to end from pinned host buffers --------------------------------------------------------
    dd.barrier()
    t0 = time.perf_counter()
    last = None
    for _ in range(args.steps):
        b = hb.to_dev()
        loss, _ = trainer.update(*b)
        last = float(loss)           # device -> host read of the step's result
    torch.cuda.synchronize()
    dt = dd.rank_max(time.perf_counter() - t0, dev)
    out['loss'] = last
    out['e2e'] = {'value': hb.frames * dd.world * args.steps / dt, 'unit': 'frames/s', 'h2d_bytes_per_step': hb.bytes,
                  'd2h_bytes_per_step': 4}
    # ---- separate pass: per-kernel times ---------------------------------------------------------------------------
    psteps = min(args.steps, 3)
    prof, _ = profile_pass(lib, trainer, batch, psteps, dd)
    out['roofline'], out['roofline_gemm'] = rooflines(w, wkey, prof, psteps, out['ms_per_step'])
    # ---- N > 1: the step's collective alone, and the strong-scaling split of ONE minibatch --------------------------
    if dd.world > 1:
        import torch.distributed as dist
        flat = trainer.model.store.grad
        for _ in range(3):
            dist.all_reduce(flat)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dd.barrier()
        e0.record()
        for _ in range(10):
            dist.all_reduce(flat)
        e1.record()
        dd.barrier()
        ar_ms = dd.rank_max(e0.elapsed_time(e1), dev) / 10
        nbytes = flat.numel() * 4
        out['allreduce'] = {'bytes': nbytes, 'ms': ar_ms, 'bus_gbs': 2.0 * (dd.world - 1) / dd.world * nbytes / (ar_ms * 1e-3) / 1e9,
                            'alg_gbs': nbytes / (ar_ms * 1e-3) / 1e9, 'n': dd.world,
                            'what': 'NCCL all_reduce(SUM) of the flat fp32 gradient buffer, 10 back-to-back calls, CUDA events, '
                                    'max over ranks; bus = 2(n-1)/n x bytes / time'}
        if w['B'] % dd.world == 0:
            g = synth_batch(w, 0)                                  # ONE global minibatch, the same on every rank
            shard = [a[dd.rank::dd.world] for a in g]              # rank r takes utterances r::n
            sb = HostBatch(shard, dev)
            sbatch = sb.to_dev()
            for _ in range(W):
                trainer.update(*sbatch)
            sms, _ = timed_steps(trainer, sbatch, args.steps, dd, dev)
            frames = int(g[1].sum())
            sprof, _ = profile_pass(lib, trainer, sbatch, psteps, dd)
            us = {k: round(v[1] / max(v[0], 1) * 1e3 / w['T'], 3) for k, v in sprof.items() if k.startswith('blstm_rec')}
            out['strong'] = {'value': frames * args.steps / (sms * 1e-3), 'unit': 'frames/s', 'ms_per_step': sms / args.steps,
                             'scaling': 'strong', 'global_batch': w['B'], 'per_gpu_batch': w['B'] // dd.world,
                             'us_per_serial_step': us,
                             'what': 'ONE %d-utterance minibatch split over the %d GPUs (rank r takes utterances r::n), one '
                                     'gradient all-reduce; the T serial steps per layer and direction do not shrink with the '
                                     'batch, so this curve is bounded by the per-step latency' % (w['B'], dd.world)}
    if with_cpu:
        cores = cpu_threads()
        fps, sec, desc, n = run_cpu(w, 3, 1, budget_s=25)
        out['cpu_baseline'] = {'value': fps, 'unit': 'frames/s', 'cores': cores, 'kind': 'port', 'sample': desc}
        if w['kind'] == 'ctc' and extras:
            try:
                out['ctc_loss_delta_vs_cpu'] = ctc_loss_delta(trainer, w, dev)
            except Exception as e:          # a reported extra: it must never cost the bench line
                out['ctc_loss_delta_vs_cpu'] = {'value': None, 'error': '%s: %s' % (type(e).__name__, e)}
    del trainer
    torch.cuda.empty_cache()
    return out


def bench_decode(w, wkey, args, dd, dev, lib, check=True):
    """configs[3]: utterances/s and frames/s of Decoder.__call__ (the recognizer's call), ids parity vs the oracle on the
    first utterances in the same run."""
    import torch
    from nabu_b200.neuralnetworks.decoders import decoder_factory
    from nabu_b200.neuralnetworks.models.model import Model
    model = Model(model_conf(w), 1, None, seed=7).build({'features': w['D']}, dev)
    V = w['V']
    alphabet = ' '.join('s%d' % i for i in range(V - 1))
    if w['kind'] == 'las_decode':
        # make EOS reachable with random weights so that hypotheses finish like a trained model's do
        st = model.store
        with torch.no_grad():
            st.vars['Speller/decoder/dense/kernel'].data.mul_(6.0)
            st.vars['Speller/decoder/dense/bias'].data[V - 1] += 1.5
        dconf = make_conf('[decoder]\ndecoder = beam_search_decoder\nmax_steps = %d\nbeam_width = %d\nalphabet = %s <eos>\n'
                          % (w['max_steps'], w['beam'], alphabet))
        dec = decoder_factory.factory('beam_search_decoder')(dconf, model)
    else:
        dconf = make_conf('[decoder]\ndecoder = ctc_decoder\ntext_alphabet = %s\n' % alphabet)
        dec = decoder_factory.factory('ctc_decoder')(dconf, model)
    x, lens, _, _ = synth_batch(w, dd.rank)
    hx = torch.from_numpy(x).pin_memory()
    hl = torch.from_numpy(lens).pin_memory()
    dx, dl = hx.to(dev), hl.to(dev)
    W = max(args.warmup, 3)
    for _ in range(W):
        outv = dec({'features': dx}, {'features': dl})
    launches0 = lib.nabu_kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dd.barrier()
    e0.record()
    for _ in range(args.steps):
        outv = dec({'features': dx}, {'features': dl})
    e1.record()
    dd.barrier()
    ms = dd.rank_max(e0.elapsed_time(e1), dev)
    launches = lib.nabu_kernel_launches() - launches0
    frames = int(lens.sum())
    res = {'value': frames * dd.world * args.steps / (ms * 1e-3), 'unit': 'frames/s', 'ms_per_step': ms / args.steps,
           'utterances_per_s': w['B'] * dd.world * args.steps / (ms * 1e-3), 'gpu_launches': int(launches),
           'workload': w['name']}
    # end to end: host features in, decoded ids back on the host
    dd.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        outv = dec({'features': hx.to(dev, non_blocking=True)}, {'features': hl.to(dev, non_blocking=True)})
        if w['kind'] == 'las_decode':
            seqs = list(outv.values())[0][0].cpu()
            d2h = seqs.numel() * 4
        else:
            d2h = list(outv.values())[0].values.size * 4
    torch.cuda.synchronize()
    dt = dd.rank_max(time.perf_counter() - t0, dev)
    res['e2e'] = {'value': frames * dd.world * args.steps / dt, 'unit': 'frames/s',
                  'h2d_bytes_per_step': hx.numel() * 4 + hl.numel() * 4, 'd2h_bytes_per_step': int(d2h)}
    prof, _ = profile_pass_decode(lib, dec, dx, dl, 1, dd)
    res['kernel_time_shares'] = {k: round(v[1] / max(res['ms_per_step'], 1e-9), 4)
                                 for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:8]}
    if w['kind'] == 'las_decode':
        seqs, lengths, scores, aligns = list(outv.values())[0]
        nsteps = int(seqs.shape[2])
        res['beam_steps'] = nsteps
        beam_ms = sum(v[1] for k, v in prof.items() if k.startswith('las_') or k.startswith('dec_'))
        res['ms_per_beam_step'] = beam_ms / max(nsteps, 1)
    if check and dd.rank == 0:
        try:
            res['ids_parity'] = decode_parity(w, model, x, lens, outv)
        except Exception as e:
            res['ids_parity'] = {'error': '%s: %s' % (type(e).__name__, e)}
    del model, dec
    torch.cuda.empty_cache()
    return res


def profile_pass_decode(lib, dec, dx, dl, steps, dd):
    import torch
    dd.barrier()
    lib.nabu_profile_enable(1)
    for _ in range(steps):
        dec({'features': dx}, {'features': dl})
    torch.cuda.synchronize()
    lib.nabu_profile_enable(0)
    cbuf = ctypes.create_string_buffer(1 << 17)
    lib.nabu_profile_collect(cbuf, 1 << 17)
    return json.loads(cbuf.value.decode()), None


def decode_parity(w, model, x, lens, outv, n=2):
    """token ids of the first `n` utterances against the oracle (fp32 restatement of TF's decoders) on the same weights"""
    import oracle as O
    params = model.store.to_numpy()
    if w['kind'] == 'ctc_decode':
        # the oracle's prefix beam search is a Python loop (~40 ms per frame): both sides decode the first 150 frames of
        # the CUDA model's own logits for `n` utterances
        import torch
        from nabu_b200 import engine
        dev = model.store.theta.device
        Tc = min(150, x.shape[1])
        with torch.no_grad():
            logits, _ = model({'features': torch.from_numpy(x[:n]).to(dev)}, {'features': torch.from_numpy(lens[:n]).to(dev)},
                              [], [], False)
            lg = list(logits.values())[0][:, :Tc].contiguous()
            ids, olen, _ = engine.ctc_beam_search(lg, torch.full((n,), Tc, dtype=torch.int32, device=dev), 100, True)
        ids, olen, lg = ids.cpu().numpy(), olen.cpu().numpy(), lg.cpu().numpy()
        exact = []
        for b in range(n):
            ref, _ = O.ctc_beam_search(lg[b], Tc, 100, True)
            exact.append(bool(olen[b] == len(ref) and np.array_equal(ids[b, :olen[b]], ref)))
        return {'checked': n, 'frames': Tc, 'ids_bit_exact': exact,
                'note': 'first %d frames of the CUDA logits decoded by both sides; the full-length check (T = 1500) is '
                        'tests/test_gpu_baseline_sizes.py' % Tc}
    from tests.test_gpu_speller import las_oracle_params
    layers, sp = las_oracle_params(params, w['layers'])
    enc, elens, _ = O.listener_fwd(x[:n], lens[:n], layers, 2, np.float32)
    ref = O.las_beam_search(enc.astype(np.float32), elens, sp, w['beam'], w['max_steps'], 'location_aware', w['dec_layers'],
                            1.0, 1.0, np.float32)
    seqs = list(outv.values())[0][0].cpu().numpy()
    L = min(seqs.shape[2], ref[0].shape[2])
    exact = [bool(np.array_equal(seqs[b, :, :L], ref[0][b, :, :L])) for b in range(n)]
    best = [bool(np.array_equal(seqs[b, 0, :L], ref[0][b, 0, :L])) for b in range(n)]
    return {'checked': n, 'all_beams_bit_exact': exact, 'best_hypothesis_bit_exact': best,
            'note': 'encoder on the CUDA path vs fp32 numpy oracle end to end; a near-tie inside fp32 rounding can '
                    'legitimately flip a beam (tests/test_gpu_baseline_sizes.py::test_cfg4 separates those cases)'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='dblstm_ctc', choices=sorted(WORKLOADS))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip the las / decode objects of the default line')
    args = ap.parse_args()
    wkey = args.workload
    w = dict(WORKLOADS[wkey])
    if os.environ.get('NABU_BENCH_T'):       # profiling aid only (ncu captures); never a bench value
        w['T'] = int(os.environ['NABU_BENCH_T'])
        w['name'] += ' [T overridden to %d for profiling]' % w['T']
        wkey += '@T%d' % w['T']
    dd = Dist()
    cores = os.cpu_count() or 1
    train = w['kind'] in ('ctc', 'las')
    config = {'workload': w['name'], 'per_gpu_batch': w['B'], 'global_batch': w['B'] * dd.world, 'frames_per_utt': w['T'],
              'parallelism': 'dp%d' % dd.world, 'precision_mode': 'fp32 parity (input_noise=0, dropout=1)',
              'l2': 'per-step working set (saved activations, >20 GB) is far larger than the 126 MB L2'}
    base = {'metric': 'acoustic frames/sec (train step)' if train else 'acoustic frames/sec (decode)', 'unit': 'frames/s',
            'n_gpus': dd.world, 'steps': args.steps, 'warmup': args.warmup, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config}

    if args.impl == 'reference':
        if dd.rank != 0:
            return
        if not train:
            print(json.dumps({'impl': 'reference', 'unavailable': 'the CPU arm times the train-step workloads only'}))
            return
        cores = cpu_threads()
        fps, sec, desc, n = run_cpu(w, args.steps, min(args.warmup, 1))
        out = dict(base)
        config['workload'] = w['name'] + ' -- CPU arm: ' + desc
        out.update({'impl': 'reference', 'value': fps, 'steps': n, 'ms_per_step': sec * 1e3,
                    'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': cores, 'kind': 'port', 'sample': desc},
                    'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                    'gpu_launches': 0})
        print(json.dumps(out))
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the product path has no CPU fallback)')
    torch.cuda.set_device(dd.local_rank)
    dev = torch.device('cuda', dd.local_rank)
    if dd.world > 1:
        dist.init_process_group('nccl', device_id=dev)
    from nabu_b200 import lib as L
    lib = L.load()

    out = dict(base)
    with_cpu = dd.rank == 0 and dd.world == 1 and not args.no_cpu_baseline
    if train:
        out.update(bench_train(w, wkey, args, dd, dev, lib, with_cpu))
    else:
        r = bench_decode(w, wkey, args, dd, dev, lib)
        out.update({k: r[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'e2e')})
        out['decode'] = r
        out['roofline'] = None
    # ---- the other BASELINE configs, measured in the same default run (N = 1) so that the driver records them ----------
    if wkey == 'dblstm_ctc' and dd.world == 1 and not args.no_extras:
        sub = argparse.Namespace(steps=min(args.steps, 5), warmup=3)
        try:
            wl = dict(WORKLOADS['las'])
            r = bench_train(wl, 'las', sub, dd, dev, lib, with_cpu, extras=False)
            r.update({'metric': base['metric'], 'unit': 'frames/s', 'config': {'workload': wl['name'], 'per_gpu_batch': wl['B'],
                                                                               'frames_per_utt': wl['T'], 'targets_per_utt': wl['U']}})
            out['las'] = r
        except Exception as e:
            out['las'] = {'error': '%s: %s' % (type(e).__name__, e)}
        out['decode'] = {}
        for k in ('las_decode', 'ctc_decode'):
            try:
                out['decode'][k] = bench_decode(dict(WORKLOADS[k]), k, sub, dd, dev, lib)
            except Exception as e:
                out['decode'][k] = {'error': '%s: %s' % (type(e).__name__, e)}
    if dd.rank == 0:
        print(json.dumps(out))
    if dd.world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
