"""Device-side plumbing between the Python mirror of nabu's plugin API and the C-ABI kernels.

* `ParamStore`: every trainable variable of a model lives in ONE flat fp32 buffer (theta) with
  matching flat grad / Adam m / Adam v buffers, so the data-parallel step is a single NCCL
  all-reduce over `grad` followed by a single fused clip+Adam launch (trainer.py:556-569).
  Variables are named like the reference's TF variable scopes.
* autograd Functions: one per C-ABI forward/backward pair.  Weight gradients are written by the
  kernels straight into the flat grad buffer (each variable is used once per step), so autograd
  only carries activations.
"""
import ctypes
import math

import numpy as np
import torch

from . import lib as L


# ---------------------------------------------------------------------------------------------
# parameters
# ---------------------------------------------------------------------------------------------

def glorot_uniform_(gen, shape):
    """tf.glorot_uniform_initializer (get_variable default in TF-1.8)."""
    if len(shape) == 1:
        fan_in = fan_out = shape[0]
    elif len(shape) == 2:
        fan_in, fan_out = shape
    else:
        rf = int(np.prod(shape[:-2]))
        fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2 - 1) * limit


class Variable(object):
    __slots__ = ('name', 'shape', 'offset', 'numel', 'data', 'grad', 'init')

    def __init__(self, name, shape, init):
        self.name, self.shape, self.init = name, tuple(shape), init
        self.numel = int(np.prod(shape))
        self.offset = None
        self.data = None
        self.grad = None


class ParamStore(object):
    """Flat fp32 parameter / gradient / Adam-moment buffers."""

    ALIGN = 64   # floats; keeps every variable 256-byte aligned for 128-bit accesses

    def __init__(self, seed=0):
        self.vars = {}
        self.order = []
        self.theta = self.grad = self.m = self.v = None
        self.seed = seed
        self.device = None

    def get(self, name, shape, init='glorot'):
        """tf.get_variable with AUTO_REUSE semantics."""
        if name in self.vars:
            v = self.vars[name]
            if v.shape != tuple(shape):
                raise ValueError('variable %s: shape %s != %s' % (name, v.shape, tuple(shape)))
            return v
        if self.theta is not None:
            raise RuntimeError('ParamStore already materialised; cannot add %s' % name)
        v = Variable(name, shape, init)
        self.vars[name] = v
        self.order.append(v)
        return v

    @property
    def materialised(self):
        return self.theta is not None

    def materialise(self, device):
        off = 0
        for v in self.order:
            v.offset = off
            off += (v.numel + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.size = off
        gen = torch.Generator().manual_seed(self.seed)
        host = torch.zeros(off, dtype=torch.float32)
        for v in self.order:
            if v.init == 'glorot':
                host[v.offset:v.offset + v.numel] = glorot_uniform_(gen, v.shape).reshape(-1)
            elif v.init == 'zeros':
                pass
            else:
                raise ValueError(v.init)
        self.device = torch.device(device)
        self.theta = host.to(self.device)
        self.grad = torch.zeros_like(self.theta)
        self.m = torch.zeros_like(self.theta)
        self.v = torch.zeros_like(self.theta)
        for v in self.order:
            d = self.theta[v.offset:v.offset + v.numel].view(v.shape)
            d.requires_grad_(True)
            v.data = d
            v.grad = self.grad[v.offset:v.offset + v.numel].view(v.shape)
        return self

    def load_numpy(self, arrays):
        """Overwrite variables from {name: ndarray} (shared weights for parity runs)."""
        with torch.no_grad():
            for name, arr in arrays.items():
                v = self.vars[name]
                v.data.copy_(torch.as_tensor(np.asarray(arr, np.float32)).reshape(v.shape))

    def to_numpy(self):
        return {v.name: v.data.detach().cpu().numpy().copy() for v in self.order}

    def grads_numpy(self):
        return {v.name: v.grad.detach().cpu().numpy().copy() for v in self.order}

    def num_params(self):
        return sum(v.numel for v in self.order)

    def state_dict(self):
        return {'theta': self.theta.cpu(), 'm': self.m.cpu(), 'v': self.v.cpu(),
                'names': [(v.name, v.shape, v.offset) for v in self.order]}

    # ---- TF checkpoint interop: the store's variable names ARE the reference's TF names --
    def save_tf_checkpoint(self, prefix, with_adam=False, global_step=None, extra=None, state_file=True):
        """Write `<prefix>.index` / `.data-*` as `tf.train.Saver(model.variables, sharded=True).save` would
        (components/hooks.py:32-52 -> model/network.ckpt).  `with_adam` adds the optimizer slots under TF's names
        (`<var>/Adam`, `<var>/Adam_1`), what the reference's validated.ckpt holds besides the variables."""
        from .processing import tfcheckpoint
        theta, m, v = self.theta.detach().cpu().numpy(), self.m.cpu().numpy(), self.v.cpu().numpy()
        arrays = {}
        for var in self.order:
            sl = slice(var.offset, var.offset + var.numel)
            arrays[var.name] = theta[sl].reshape(var.shape)
            if with_adam:
                arrays[var.name + '/Adam'] = m[sl].reshape(var.shape)
                arrays[var.name + '/Adam_1'] = v[sl].reshape(var.shape)
        if global_step is not None:
            arrays['global_step'] = np.array(global_step, np.int32)
        arrays.update(extra or {})                # the trainer's other global variables (learning_rate_fact, ...)
        tfcheckpoint.write_checkpoint(prefix, arrays, state_file=state_file)

    def load_tf_checkpoint(self, prefix, with_adam=False):
        """Restore every variable from a TF checkpoint (a nabu-trained `model/network.ckpt`, LoadAtBegin
        components/hooks.py:6-28).  Like Saver.restore it fails when a variable is missing or its shape differs;
        keys of the checkpoint the model does not own are ignored.  Returns the checkpoint's global_step or None."""
        from .processing import tfcheckpoint
        have = {name: shape for name, shape, _ in tfcheckpoint.list_variables(prefix)}
        want = [var.name for var in self.order]
        if with_adam:
            want += [n + s for n in want for s in ('/Adam', '/Adam_1')]
        missing = [n for n in want if n not in have]
        if missing:
            raise KeyError('not found in checkpoint %s: %s' % (prefix, ', '.join(missing)))
        for var in self.order:
            if tuple(have[var.name]) != var.shape:
                raise ValueError('%s: checkpoint shape %s, model shape %s' % (var.name, have[var.name], var.shape))
        names = want + (['global_step'] if 'global_step' in have else [])
        arrays = tfcheckpoint.read_checkpoint(prefix, names=set(names))
        with torch.no_grad():
            for var in self.order:
                sl = slice(var.offset, var.offset + var.numel)
                self.theta[sl].copy_(torch.from_numpy(arrays[var.name].astype(np.float32).reshape(-1)))
                if with_adam:
                    self.m[sl].copy_(torch.from_numpy(arrays[var.name + '/Adam'].astype(np.float32).reshape(-1)))
                    self.v[sl].copy_(torch.from_numpy(arrays[var.name + '/Adam_1'].astype(np.float32).reshape(-1)))
        return int(arrays['global_step']) if 'global_step' in arrays else None

    def load_state_dict(self, sd):
        self.theta.copy_(sd['theta'])
        self.m.copy_(sd['m'])
        self.v.copy_(sd['v'])


def clip_adam_step(store, lr, t, beta1=0.9, beta2=0.999, eps=1e-8, clip=1.0, grad_scale=1.0):
    lib = L.load()
    L.check(lib.nabu_clip_adam_step(L.ptr(store.theta), L.ptr(store.grad), L.ptr(store.m), L.ptr(store.v),
                                    store.size, lr, t, beta1, beta2, eps, clip, grad_scale, L.stream()),
            'nabu_clip_adam_step')


# ---------------------------------------------------------------------------------------------
# ops
# ---------------------------------------------------------------------------------------------

def _i32(t, device):
    return t.to(device=device, dtype=torch.int32).contiguous()


def planes_enabled():
    """operand planes travel with the activations (include/nabu_b200.h: nabu_blstm_*_planes); NABU_PLANES=0 turns it off"""
    import os
    return os.environ.get('NABU_PLANES', '1') != '0'


def _blstm_fwd_raw(x, lens, kf, bf, kb, bb, H, yT, x_planes=None, want_planes=False):
    lib = L.load()
    B, T, D = x.shape
    y = torch.empty((B, yT, 2 * H), device=x.device, dtype=torch.float32)
    gates = torch.empty((2, B, T, 4 * H), device=x.device, dtype=torch.float32)
    cells = torch.empty((2, B, T, H), device=x.device, dtype=torch.float32)
    nws = lib.nabu_blstm_workspace_bytes(B, T, D, H)
    if nws == 0:
        L.check(2, 'nabu_blstm_workspace_bytes')
    ws = L.WORKSPACE.get(nws, x.device)
    y_planes = None
    if want_planes and H % 4 == 0:
        y_planes = torch.empty(lib.nabu_blstm_planes_bytes(B, yT, H), device=x.device, dtype=torch.uint8)
    if x_planes is not None and D % 8:
        x_planes = None
    L.check(lib.nabu_blstm_fwd_planes(L.ptr(x), L.ptr(x_planes), L.ptr(lens), B, T, D, H, L.ptr(kf), L.ptr(bf), L.ptr(kb),
                                      L.ptr(bb), L.ptr(y), L.ptr(y_planes), yT, L.ptr(gates), L.ptr(cells), L.ptr(ws),
                                      ws.numel(), L.stream()), 'nabu_blstm_fwd_planes')
    return y, gates, cells, y_planes


# max |dx| of the last BLSTM backward call, for the call whose dy IS that dx (nabu_blstm_bwd_hints): the tensor itself is
# held (its memory cannot be handed to another tensor meanwhile) together with its version counter (any in-place
# modification, autograd's gradient accumulation included, bumps it; views share it).
_DXMAX = {'dx': None, 'version': None, 'buf': None}


def _dy_absmax_hint(dy):
    dx, ver, buf = _DXMAX['dx'], _DXMAX['version'], _DXMAX['buf']
    _DXMAX['dx'] = _DXMAX['version'] = _DXMAX['buf'] = None
    if dx is None or not dy.is_contiguous():
        return None
    same = dy.data_ptr() == dx.data_ptr() and dy.numel() == dx.numel() and dy.dtype == dx.dtype and dy._version == ver
    return buf if same else None


def _blstm_bwd_raw(x, lens, kf, kb, y, gates, cells, dy, need_dx, H, yT, gvars, x_planes=None, y_planes=None):
    lib = L.load()
    B, T, D = x.shape
    dkf, dbf, dkb, dbb = gvars
    hint_in = _dy_absmax_hint(dy)
    dy = dy.contiguous()
    dx = torch.empty_like(x) if need_dx else None
    hint_out = torch.empty(128, dtype=torch.int32, device=x.device) if need_dx else None
    L.check(lib.nabu_blstm_bwd_hints(L.ptr(hint_out), L.ptr(hint_in)), 'nabu_blstm_bwd_hints')
    ws = L.WORKSPACE.get(lib.nabu_blstm_workspace_bytes(B, T, D, H), x.device)
    if x_planes is not None and D % 8:
        x_planes = None
    L.check(lib.nabu_blstm_bwd_planes(L.ptr(x), L.ptr(x_planes), L.ptr(lens), B, T, D, H, L.ptr(kf), L.ptr(kb), L.ptr(y),
                                      L.ptr(y_planes), yT, L.ptr(gates), L.ptr(cells), L.ptr(dy), L.ptr(dx), L.ptr(dkf),
                                      L.ptr(dbf), L.ptr(dkb), L.ptr(dbb), L.ptr(ws), ws.numel(), L.stream()),
            'nabu_blstm_bwd_planes')
    if _OVERLAP['on']:
        # deferred weight gradients read these on the library's side stream until side_join()
        _OVERLAP['keep'].append((x, y, gates, kf, kb, gvars, x_planes, y_planes, hint_in))
    if need_dx:
        _DXMAX['dx'], _DXMAX['version'], _DXMAX['buf'] = dx, dx._version, hint_out
    return dx


class _BLSTM(torch.autograd.Function):
    """components/layer.py:8-51 via nabu_blstm_fwd_planes / nabu_blstm_bwd_planes.  Returns (y, y_planes): the second output
    is the fp16 operand-plane image of y (uint8, not differentiable; an empty tensor when planes are off) that the
    consumer of y hands back as `x_planes`."""

    @staticmethod
    def forward(ctx, x, lens, kf, bf, kb, bb, H, yT, gvars, x_planes):
        x = x.contiguous()
        y, gates, cells, y_planes = _blstm_fwd_raw(x, lens, kf, bf, kb, bb, H, yT, x_planes, planes_enabled())
        ctx.save_for_backward(x, lens, kf, kb, y, gates, cells)
        ctx.H, ctx.yT, ctx.gvars = H, yT, gvars
        ctx.x_planes, ctx.y_planes = x_planes, y_planes
        ctx.need_dx = x.requires_grad
        out_planes = y_planes if y_planes is not None else torch.empty(0, dtype=torch.uint8, device=x.device)
        ctx.mark_non_differentiable(out_planes)
        # without this autograd hands backward() a zero-filled "gradient" of the planes: a 786 MB byte fill per layer at cfg-3
        # (5 x 0.2 ms per step in profiles/r2f_launches_step.csv)
        ctx.set_materialize_grads(False)
        return y, out_planes

    @staticmethod
    def backward(ctx, dy, _dplanes):
        x, lens, kf, kb, y, gates, cells = ctx.saved_tensors
        if dy is None:
            dy = torch.zeros_like(y)
        dx = _blstm_bwd_raw(x, lens, kf, kb, y, gates, cells, dy, ctx.need_dx, ctx.H, ctx.yT, ctx.gvars, ctx.x_planes,
                            ctx.y_planes)
        return dx, None, None, None, None, None, None, None, None, None


_OVERLAP = {'on': False, 'keep': []}


def set_overlap(on):
    """Deferred weight gradients (include/nabu_b200.h: nabu_set_overlap).  The trainer turns this on; side_join() must
    run between the backward pass and the first reader of the gradients."""
    L.check(L.load().nabu_set_overlap(1 if on else 0), 'nabu_set_overlap')
    _OVERLAP['on'] = bool(on)


def side_join():
    _DXMAX['dx'] = _DXMAX['version'] = _DXMAX['buf'] = None
    if _OVERLAP['on']:
        L.check(L.load().nabu_side_join(L.stream()), 'nabu_side_join')
        _OVERLAP['keep'].clear()


REC_UNIT = 64        # the recurrence kernels tile the hidden units by 64 (csrc/blstm.cu: KC)


def _pad_gates(w, H, Hp, rows_h):
    """[(D+H) | 1, 4H] -> [(D+Hp) | 1, 4Hp]: every gate block i,j,f,o widened to Hp columns, the h rows to Hp, zeros"""
    lead = w.shape[0] - H if rows_h else None
    g = w.reshape(w.shape[:-1] + (4, H))
    g = torch.nn.functional.pad(g, (0, Hp - H)).reshape(w.shape[:-1] + (4 * Hp,))
    if rows_h:
        g = torch.cat([g, g.new_zeros((Hp - H, 4 * Hp))], 0)
        assert g.shape[0] == lead + Hp
    return g.contiguous()


def _unpad_gates(g, H, Hp, rows_h):
    if rows_h:
        g = g[:g.shape[0] - (Hp - H)]
    return g.reshape(g.shape[:-1] + (4, Hp))[..., :H].reshape(g.shape[:-1] + (4 * H,))


class _BLSTMPadded(torch.autograd.Function):
    """num_units that is not a multiple of 64 (the reference takes any): the same kernels on Hp = ceil64(H) units.  The
    extra units have zero weights and biases: every gate pre-activation is 0, so c' = c*sigmoid(1) + sigmoid(0)*tanh(0)
    stays 0 from the zero initial state and h = 0 -- they contribute nothing forward or backward, and the valid
    units' arithmetic is unchanged.  Costs a padded copy of the weights and a slice of the outputs per call."""

    @staticmethod
    def forward(ctx, x, lens, kf, bf, kb, bb, H, yT, gvars):
        Hp = rec_width(H, x.shape[0])
        x = x.contiguous()
        pkf, pbf = _pad_gates(kf.detach(), H, Hp, True), _pad_gates(bf.detach(), H, Hp, False)
        pkb, pbb = _pad_gates(kb.detach(), H, Hp, True), _pad_gates(bb.detach(), H, Hp, False)
        yp, gates, cells, _ = _blstm_fwd_raw(x, lens, pkf, pbf, pkb, pbb, Hp, yT)
        ctx.save_for_backward(x, lens, pkf, pkb, yp, gates, cells)
        ctx.H, ctx.Hp, ctx.yT, ctx.gvars = H, Hp, yT, gvars
        ctx.need_dx = x.requires_grad
        return torch.cat([yp[..., :H], yp[..., Hp:Hp + H]], -1)

    @staticmethod
    def backward(ctx, dy):
        x, lens, pkf, pkb, yp, gates, cells = ctx.saved_tensors
        H, Hp = ctx.H, ctx.Hp
        dyp = dy.new_zeros(dy.shape[:-1] + (2 * Hp,))
        dyp[..., :H] = dy[..., :H]
        dyp[..., Hp:Hp + H] = dy[..., H:]
        pgrads = (torch.empty_like(pkf), pkf.new_empty(4 * Hp), torch.empty_like(pkb), pkb.new_empty(4 * Hp))
        dx = _blstm_bwd_raw(x, lens, pkf, pkb, yp, gates, cells, dyp, ctx.need_dx, Hp, ctx.yT, pgrads)
        side_join()                                      # deferred weight gradients must have landed before the slices
        for real, pg, rows_h in zip(ctx.gvars, pgrads, (True, False, True, False)):
            real.copy_(_unpad_gates(pg, H, Hp, rows_h))
        return dx, None, None, None, None, None, None, None, None


def rec_width(H, B=None):
    """The number of hidden units the recurrence kernels run for a layer of `H` units.  The tcgen05 cluster kernels exist
    for 256, 512 and 1024 units (csrc/blstm_cl_fwdc.cu, blstm_cl_bwd8c.cu); a narrower layer -- every recipe the reference ships
    uses num_units = 128 -- is zero-padded up to the next of those (exact, see _BLSTMPadded) when the batch is large
    enough for that to win: measured on a B200 at num_units = 128, T = 800 (tools/h128_probe.py, profiles/r2_h128_probe.txt)
    the padded tcgen05 path takes 6.6 / 10.7 us per time step forward / backward against 8.9 / 12.1 on the FFMA cluster
    kernels at B = 128, but 5.7 / 7.4 against 5.6 / 5.8 at B = 16.  NABU_PAD_UNITS=0 / 1 forces it off / on.
    Otherwise: the next multiple of the 64-unit tile."""
    import os
    Hp = (H + REC_UNIT - 1) // REC_UNIT * REC_UNIT
    mode = os.environ.get('NABU_PAD_UNITS', 'auto')
    if mode != '0' and H < 1024 and H not in (256, 512):
        if H > 512:                    # 513 .. 1023 units: the FFMA kernels take 17 / 35 us per time step there
            return 1024
        if mode == '1' or (B is not None and B > 32):
            return 256 if H <= 256 else 512
    return Hp


def blstm(x, lens, vf_k, vf_b, vb_k, vb_b, H, yT=None, x_planes=None, want_planes=False):
    """x [B,T,D] -> y [B,yT,2H]; v*_ are engine.Variable.  `x_planes`: the operand planes of x when x IS the unchanged
    output of another blstm (or its pyramid_stack reshape); `want_planes`: also return y's planes (or None)."""
    yT = x.shape[1] if yT is None else yT
    gv = (vf_k.grad, vf_b.grad, vb_k.grad, vb_b.grad)
    if rec_width(H, x.shape[0]) == H:
        y, planes = _BLSTM.apply(x, lens, vf_k.data, vf_b.data, vb_k.data, vb_b.data, H, yT, gv, x_planes)
        planes = planes if planes.numel() else None
    else:
        y, planes = _BLSTMPadded.apply(x, lens, vf_k.data, vf_b.data, vb_k.data, vb_b.data, H, yT, gv), None
    return (y, planes) if want_planes else y


def pyramid_lengths(lens, numsteps):
    lib = L.load()
    out = torch.empty_like(lens)
    L.check(lib.nabu_pyramid_lengths(L.ptr(lens), lens.numel(), numsteps, L.ptr(out), L.stream()),
            'nabu_pyramid_lengths')
    return out


class _Linear(torch.autograd.Function):
    """models/ed_decoders/dnn_decoder.py:53-57 via nabu_linear_fwd / nabu_linear_bwd."""

    @staticmethod
    def forward(ctx, x, W, b, gvars):
        lib = L.load()
        x = x.contiguous()
        D, V = W.shape
        N = x.numel() // D
        y = torch.empty(x.shape[:-1] + (V,), device=x.device, dtype=torch.float32)
        L.check(lib.nabu_linear_fwd(L.ptr(x), N, D, V, L.ptr(W), L.ptr(b), L.ptr(y), None, 0, L.stream()),
                'nabu_linear_fwd')
        ctx.save_for_backward(x, W)
        ctx.gvars = gvars
        ctx.need_dx = x.requires_grad
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = L.load()
        x, W = ctx.saved_tensors
        D, V = W.shape
        N = x.numel() // D
        dW, db = ctx.gvars
        dy = dy.contiguous()
        dx = torch.empty_like(x) if ctx.need_dx else None
        ws = L.WORKSPACE.get(lib.nabu_gemm_workspace_bytes(), x.device)
        L.check(lib.nabu_linear_bwd(L.ptr(x), N, D, V, L.ptr(W), L.ptr(dy), L.ptr(dx), L.ptr(dW), L.ptr(db),
                                    L.ptr(ws), ws.numel(), L.stream()), 'nabu_linear_bwd')
        return dx, None, None, None


def linear(x, vW, vb):
    return _Linear.apply(x, vW.data, vb.data, (vW.grad, vb.grad))


class _CTC(torch.autograd.Function):
    """trainers/loss_functions.py:203-212: mean over the batch of tf.nn.ctc_loss."""

    @staticmethod
    def forward(ctx, logits, logit_len, labels, label_len):
        lib = L.load()
        logits = logits.contiguous()
        B, T, V = logits.shape
        labels = labels.contiguous()
        Lmax = labels.shape[1]
        loss = torch.empty(B, device=logits.device, dtype=torch.float32)
        grad = torch.empty_like(logits)
        ws = L.WORKSPACE.get(lib.nabu_ctc_workspace_bytes(B, T, V, Lmax), logits.device)
        L.check(lib.nabu_ctc_loss_fwd_bwd(L.ptr(logits), L.ptr(logit_len), L.ptr(labels), Lmax, L.ptr(label_len),
                                          B, T, V, 1.0 / B, L.ptr(loss), L.ptr(grad), L.ptr(ws), ws.numel(),
                                          L.stream()), 'nabu_ctc_loss_fwd_bwd')
        ctx.save_for_backward(grad)
        ctx.per_utt = loss
        return loss.mean()

    @staticmethod
    def backward(ctx, dloss):
        (grad,) = ctx.saved_tensors
        return grad * dloss, None, None, None


def ctc_loss_per_utt(logits, logit_len, labels, label_len, want_grad=False, grad_scale=1.0):
    """Raw per-utterance NLL (and optionally grad_scale * dNLL/dlogits)."""
    lib = L.load()
    logits = logits.contiguous()
    B, T, V = logits.shape
    labels = labels.contiguous()
    Lmax = labels.shape[1]
    loss = torch.empty(B, device=logits.device, dtype=torch.float32)
    grad = torch.empty_like(logits) if want_grad else None
    ws = L.WORKSPACE.get(lib.nabu_ctc_workspace_bytes(B, T, V, Lmax), logits.device)
    L.check(lib.nabu_ctc_loss_fwd_bwd(L.ptr(logits), L.ptr(logit_len), L.ptr(labels), Lmax, L.ptr(label_len), B, T,
                                      V, grad_scale, L.ptr(loss), L.ptr(grad), L.ptr(ws), ws.numel(), L.stream()),
            'nabu_ctc_loss_fwd_bwd')
    return loss, grad


class _MaskedCE(torch.autograd.Function):
    """trainers/loss_functions.py:155-165 average_cross_entropy (single output)."""

    @staticmethod
    def forward(ctx, logits, targets, logit_len, target_len):
        lib = L.load()
        logits = logits.contiguous()
        B, U, V = logits.shape
        targets = targets.contiguous()
        loss = torch.empty(B, device=logits.device, dtype=torch.float32)
        grad = torch.empty_like(logits)
        L.check(lib.nabu_masked_ce_fwd_bwd(L.ptr(logits), L.ptr(targets), targets.shape[1], L.ptr(logit_len),
                                           L.ptr(target_len), B, U, V, 1.0 / B, L.ptr(loss), L.ptr(grad),
                                           L.stream()), 'nabu_masked_ce_fwd_bwd')
        ctx.save_for_backward(grad)
        return loss.mean()

    @staticmethod
    def backward(ctx, dloss):
        (grad,) = ctx.saved_tensors
        return grad * dloss, None, None, None


def ctc_mean(logits, logit_len, labels, label_len):
    return _CTC.apply(logits, logit_len, labels, label_len)


def masked_ce_mean(logits, targets, logit_len, target_len):
    return _MaskedCE.apply(logits, targets, logit_len, target_len)


def gemm(mode, A, B, M, N, K, lda, ldb, ldc, C=None, alpha=1.0, beta=0.0, bias=None, precision=0):
    """Raw nabu_gemm call (tests)."""
    lib = L.load()
    if C is None:
        C = torch.zeros((M, ldc), device=A.device, dtype=torch.float32)
    nbytes = lib.nabu_gemm_h2_workspace_bytes(mode, M, N, K) if precision == 2 else lib.nabu_gemm_workspace_bytes()
    ws = L.WORKSPACE.get(nbytes, A.device)
    L.check(lib.nabu_gemm(mode, precision, M, N, K, alpha, L.ptr(A), lda, L.ptr(B), ldb, beta, L.ptr(C), ldc,
                          L.ptr(bias), L.ptr(ws), ws.numel(), L.stream()), 'nabu_gemm')
    return C


# ---------------------------------------------------------------------------------------------
# Speller (attention decoder) and the beam searches
# ---------------------------------------------------------------------------------------------

ATTENTION_IDS = {'vanilla': 0, 'location_aware': 1, 'windowed': 2}
PROBABILITY_FN_IDS = {'softmax': 0, 'normalized_sigmoid': 1, 'sigmoid': 2}      # components/attention.py:9-13


class SpellerVars(object):
    """The variables of Speller.create_cell as engine.Variable objects (see speller.py mirror)."""

    def __init__(self, cell_kernels, cell_biases, memory_kernel, query_kernel, attention_v, conv_kernel,
                 conv_dense_kernel, out_kernel, out_bias):
        self.cell_kernels, self.cell_biases = cell_kernels, cell_biases
        self.memory_kernel, self.query_kernel, self.attention_v = memory_kernel, query_kernel, attention_v
        self.conv_kernel, self.conv_dense_kernel = conv_kernel, conv_dense_kernel
        self.out_kernel, self.out_bias = out_kernel, out_bias

    def all(self):
        out = list(self.cell_kernels) + list(self.cell_biases) + [self.memory_kernel, self.query_kernel,
                                                                  self.attention_v]
        if self.conv_kernel is not None:
            out += [self.conv_kernel, self.conv_dense_kernel]
        return out + [self.out_kernel, self.out_bias]

    def pack(self, grad=False):
        p = L.SpellerParams()
        get = (lambda v: L.ptr(v.grad)) if grad else (lambda v: L.ptr(v.data.detach()))
        for i, (k, b) in enumerate(zip(self.cell_kernels, self.cell_biases)):
            p.cell_kernel[i] = get(k)
            p.cell_bias[i] = get(b)
        p.memory_kernel, p.query_kernel, p.attention_v = get(self.memory_kernel), get(self.query_kernel), get(
            self.attention_v)
        if self.conv_kernel is not None:
            p.conv_kernel, p.conv_dense_kernel = get(self.conv_kernel), get(self.conv_dense_kernel)
        p.out_kernel, p.out_bias = get(self.out_kernel), get(self.out_bias)
        return p


def speller_desc(B, Tm, E, V, H, num_layers, attention, numfilt, filtersize, U, dropout_keep=1.0, sample_prob=0.0,
                 seed=0):
    d = L.SpellerDesc()
    d.B, d.Tm, d.E, d.V, d.H, d.num_layers, d.A = B, Tm, E, V, H, num_layers, H
    base, _, pf = attention.partition('+')
    d.attention = ATTENTION_IDS[base]
    d.probability_fn = PROBABILITY_FN_IDS[pf or 'softmax']
    d.numfilt, d.filtersize, d.U = numfilt, filtersize, U
    d.dropout_keep, d.sample_prob, d.seed = float(dropout_keep), float(sample_prob), int(seed) & 0xFFFFFFFF
    return d


class _Speller(torch.autograd.Function):
    """rnn_decoder.py:40-82 (teacher forcing) via nabu_speller_fwd / nabu_speller_bwd."""

    @staticmethod
    def forward(ctx, memory, mem_len, targets, target_len, desc, svars, *param_tensors):
        lib = L.load()
        memory = memory.contiguous()
        targets = targets.contiguous()
        logits = torch.empty((desc.B, desc.U, desc.V), device=memory.device, dtype=torch.float32)
        nsaved = lib.nabu_speller_saved_bytes(ctypes.byref(desc))
        if nsaved == 0:
            L.check(2, 'nabu_speller_saved_bytes')
        saved = torch.empty(nsaved, dtype=torch.uint8, device=memory.device)
        p = svars.pack()
        L.check(lib.nabu_speller_fwd(ctypes.byref(desc), ctypes.byref(p), L.ptr(memory), L.ptr(mem_len),
                                     L.ptr(targets), targets.shape[1], L.ptr(target_len), L.ptr(logits),
                                     L.ptr(saved), None, 0, L.stream()), 'nabu_speller_fwd')
        ctx.save_for_backward(memory, mem_len, targets, target_len, saved)
        ctx.desc, ctx.svars = desc, svars
        ctx.need_dmem = memory.requires_grad
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        lib = L.load()
        memory, mem_len, targets, target_len, saved = ctx.saved_tensors
        desc, svars = ctx.desc, ctx.svars
        dlogits = dlogits.contiguous()
        dmem = torch.empty_like(memory) if ctx.need_dmem else None
        ws = L.WORKSPACE.get(lib.nabu_speller_workspace_bytes(ctypes.byref(desc)), memory.device)
        p, g = svars.pack(), svars.pack(grad=True)
        L.check(lib.nabu_speller_bwd(ctypes.byref(desc), ctypes.byref(p), L.ptr(memory), L.ptr(mem_len),
                                     L.ptr(targets), targets.shape[1], L.ptr(target_len), L.ptr(dlogits),
                                     L.ptr(saved), L.ptr(dmem), ctypes.byref(g), L.ptr(ws), ws.numel(), L.stream()),
                'nabu_speller_bwd')
        return (dmem, None, None, None, None, None) + (None,) * len(svars.all())


def speller(memory, mem_len, targets, target_len, svars, V, H, num_layers, attention, numfilt, filtersize,
            dropout_keep=1.0, sample_prob=0.0, seed=0):
    """dropout_keep < 1: DropoutWrapper(output_keep_prob) on every LSTM layer; sample_prob > 0: scheduled sampling;
    both drawn from the counter generator keyed by `seed` (use a new seed every training step)."""
    B, Tm, E = memory.shape
    U = targets.shape[1]
    desc = speller_desc(B, Tm, E, V, H, num_layers, attention, numfilt, filtersize, U, dropout_keep, sample_prob, seed)
    return _Speller.apply(memory, mem_len, targets, target_len, desc, svars, *[v.data for v in svars.all()])


def las_beam_search(memory, mem_len, svars, V, H, num_layers, attention, numfilt, filtersize, beam_width,
                    max_steps, length_penalty, temperature):
    """components/beam_search_decoder.py under dynamic_decode: returns (sequences [B,W,L], lengths [B,W],
    scores [B,W], alignments [B,W,L,Tm]) with L = the number of steps the loop ran."""
    lib = L.load()
    memory = memory.contiguous()
    B, Tm, E = memory.shape
    W = beam_width
    desc = speller_desc(B, Tm, E, V, H, num_layers, attention, numfilt, filtersize, 1)
    dev = memory.device
    seqs = torch.zeros((B, W, max_steps), dtype=torch.int32, device=dev)
    lens = torch.zeros((B, W), dtype=torch.int32, device=dev)
    scores = torch.zeros((B, W), dtype=torch.float32, device=dev)
    aligns = torch.zeros((B, W, max_steps, Tm), dtype=torch.float32, device=dev)
    nws = lib.nabu_las_beam_workspace_bytes(ctypes.byref(desc), W, max_steps)
    if nws == 0:
        L.check(2, 'nabu_las_beam_workspace_bytes')
    ws = L.WORKSPACE.get(nws, dev)
    n = ctypes.c_int(0)
    p = svars.pack()
    L.check(lib.nabu_las_beam_search(ctypes.byref(desc), ctypes.byref(p), L.ptr(memory), L.ptr(mem_len), W,
                                     max_steps, length_penalty, temperature, L.ptr(seqs), L.ptr(lens),
                                     L.ptr(scores), L.ptr(aligns), ctypes.byref(n), L.ptr(ws), ws.numel(),
                                     L.stream()), 'nabu_las_beam_search')
    n = n.value
    return seqs[:, :, :n].contiguous(), lens, scores, aligns[:, :, :n].contiguous()


def ctc_beam_search(logits, logit_len, beam_width=100, merge_repeated=True):
    """tf.nn.ctc_beam_search_decoder(top_paths=1): returns (ids [B,T] int32, lengths [B], neg log prob [B])."""
    lib = L.load()
    logits = logits.contiguous()
    B, T, V = logits.shape
    out = torch.zeros((B, T), dtype=torch.int32, device=logits.device)
    out_len = torch.zeros(B, dtype=torch.int32, device=logits.device)
    nlp = torch.zeros(B, dtype=torch.float32, device=logits.device)
    nws = lib.nabu_ctc_beam_workspace_bytes(B, T, V, beam_width)
    if nws == 0:
        L.check(2, 'nabu_ctc_beam_workspace_bytes')
    ws = L.WORKSPACE.get(nws, logits.device)
    L.check(lib.nabu_ctc_beam_search(L.ptr(logits), L.ptr(logit_len), B, T, V, beam_width, int(merge_repeated),
                                     L.ptr(out), L.ptr(out_len), L.ptr(nlp), L.ptr(ws), ws.numel(), L.stream()),
            'nabu_ctc_beam_search')
    return out, out_len, nlp
