"""ctypes binding of libnabu_b200.so (include/nabu_b200.h).

There is deliberately no fallback: if the shared library is missing, or an entry point returns a
non-zero status, this module raises -- the product path never computes on the CPU.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libnabu_b200.so')

c_int, c_float, c_size_t, c_void_p = ctypes.c_int, ctypes.c_float, ctypes.c_size_t, ctypes.c_void_p
P = c_void_p


class NabuError(RuntimeError):
    pass


class SpellerDesc(ctypes.Structure):
    _fields_ = [(n, c_int) for n in ('B', 'Tm', 'E', 'V', 'H', 'num_layers', 'A', 'attention', 'numfilt',
                                    'filtersize', 'U', 'probability_fn')] + \
               [('dropout_keep', c_float), ('sample_prob', c_float), ('seed', ctypes.c_uint)]


class SpellerParams(ctypes.Structure):
    _fields_ = [('cell_kernel', P * 4), ('cell_bias', P * 4), ('memory_kernel', P), ('query_kernel', P),
                ('attention_v', P), ('conv_kernel', P), ('conv_dense_kernel', P), ('out_kernel', P),
                ('out_bias', P)]


# name -> (restype, argtypes); must list every symbol include/nabu_b200.h declares
SIGNATURES = {
    'nabu_last_error': (ctypes.c_char_p, []),
    'nabu_version': (c_int, []),
    'nabu_kernel_launches': (ctypes.c_ulonglong, []),
    'nabu_profile_enable': (c_int, [c_int]),
    'nabu_profile_collect': (c_int, [ctypes.c_char_p, c_size_t]),
    'nabu_set_overlap': (c_int, [c_int]),
    'nabu_side_join': (c_int, [P]),
    'nabu_gemm_workspace_bytes': (c_size_t, []),
    'nabu_gemm_h2_workspace_bytes': (c_size_t, [c_int, c_int, c_int, c_int]),
    'nabu_gemm': (c_int, [c_int, c_int, c_int, c_int, c_int, c_float, P, c_int, P, c_int, c_float, P, c_int, P, P,
                          c_size_t, P]),
    'nabu_blstm_workspace_bytes': (c_size_t, [c_int] * 4),
    'nabu_blstm_fwd': (c_int, [P, P, c_int, c_int, c_int, c_int, P, P, P, P, P, c_int, P, P, P, c_size_t, P]),
    'nabu_blstm_bwd': (c_int, [P, P, c_int, c_int, c_int, c_int, P, P, P, c_int, P, P, P, P, P, P, P, P, P, c_size_t,
                               P]),
    'nabu_blstm_planes_bytes': (c_size_t, [c_int] * 3),
    'nabu_blstm_fwd_planes': (c_int, [P, P, P, c_int, c_int, c_int, c_int, P, P, P, P, P, P, c_int, P, P, P, c_size_t, P]),
    'nabu_blstm_bwd_planes': (c_int, [P, P, P, c_int, c_int, c_int, c_int, P, P, P, P, c_int, P, P, P, P, P, P, P, P, P,
                                      c_size_t, P]),
    'nabu_blstm_bwd_hints': (c_int, [P, P]),
    'nabu_pyramid_lengths': (c_int, [P, c_int, c_int, P, P]),
    'nabu_linear_fwd': (c_int, [P, c_int, c_int, c_int, P, P, P, P, c_size_t, P]),
    'nabu_linear_bwd': (c_int, [P, c_int, c_int, c_int, P, P, P, P, P, P, c_size_t, P]),
    'nabu_ctc_workspace_bytes': (c_size_t, [c_int] * 4),
    'nabu_ctc_loss_fwd_bwd': (c_int, [P, P, P, c_int, P, c_int, c_int, c_int, c_float, P, P, P, c_size_t, P]),
    'nabu_masked_ce_fwd_bwd': (c_int, [P, P, c_int, P, P, c_int, c_int, c_int, c_float, P, P, P]),
    'nabu_clip_adam_step': (c_int, [P, P, P, P, c_size_t, c_float, c_int, c_float, c_float, c_float, c_float, c_float,
                                    P]),
    'nabu_speller_workspace_bytes': (c_size_t, [ctypes.POINTER(SpellerDesc)]),
    'nabu_speller_saved_bytes': (c_size_t, [ctypes.POINTER(SpellerDesc)]),
    'nabu_speller_fwd': (c_int, [ctypes.POINTER(SpellerDesc), ctypes.POINTER(SpellerParams), P, P, P, c_int, P, P, P,
                                 P, c_size_t, P]),
    'nabu_speller_bwd': (c_int, [ctypes.POINTER(SpellerDesc), ctypes.POINTER(SpellerParams), P, P, P, c_int, P, P, P,
                                 P, ctypes.POINTER(SpellerParams), P, c_size_t, P]),
    'nabu_las_beam_workspace_bytes': (c_size_t, [ctypes.POINTER(SpellerDesc), c_int, c_int]),
    'nabu_las_beam_search': (c_int, [ctypes.POINTER(SpellerDesc), ctypes.POINTER(SpellerParams), P, P, c_int, c_int,
                                     c_float, c_float, P, P, P, P, ctypes.POINTER(c_int), P, c_size_t, P]),
    'nabu_ctc_beam_workspace_bytes': (c_size_t, [c_int] * 4),
    'nabu_ctc_beam_search': (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P, P, P, P, c_size_t, P]),
    'nabu_crc32c': (ctypes.c_uint, [ctypes.c_char_p, c_size_t, ctypes.c_uint]),
    'nabu_attn_workspace_bytes': (c_size_t, [ctypes.POINTER(SpellerDesc), c_int]),
    'nabu_attn_keys': (c_int, [ctypes.POINTER(SpellerDesc), ctypes.POINTER(SpellerParams), P, P, P, P, P]),
    'nabu_attn_step_fwd': (c_int, [ctypes.POINTER(SpellerDesc), ctypes.POINTER(SpellerParams), P, c_int, c_int, P, P, P, P,
                                   P, P, P, P, P, P, c_size_t, P]),
    'nabu_attn_step_bwd': (c_int, [ctypes.POINTER(SpellerDesc), ctypes.POINTER(SpellerParams), P, c_int, P, P, P, P, P, P,
                                   P, P, P, P, P, P, P, P, ctypes.POINTER(SpellerParams), P, c_size_t, P]),
    'nabu_ctc_prefix_beam': (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P, P, P, P, c_size_t, P]),
    'nabu_comm_unique_id': (c_int, [ctypes.c_char_p]),
    'nabu_comm_init': (c_int, [ctypes.c_char_p, c_int, c_int]),
    'nabu_comm_world': (c_int, []),
    'nabu_comm_destroy': (c_int, []),
    'nabu_allreduce_grads': (c_int, [P, c_size_t, P]),
}

_lib = None


def load():
    """Load the C-ABI library (once).  Raises NabuError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NabuError('%s not found: run `python -m nabu_b200.build` (no CPU fallback exists)' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        msg = load().nabu_last_error().decode('utf-8', 'replace')
        raise NabuError('%s failed (status %d): %s' % (what, status, msg))


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise NabuError('nabu_b200 kernels need CUDA tensors (got %s); there is no CPU path' % t.device)
    if not t.is_contiguous():
        raise NabuError('tensor must be contiguous')
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


class Workspace(object):
    """Grow-only per-device scratch buffer handed to the C-ABI calls (caller-owned memory)."""

    def __init__(self):
        self._buf = {}

    def get(self, nbytes, device):
        key = (device.index if device.index is not None else torch.cuda.current_device())
        buf = self._buf.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
            self._buf[key] = buf
        return buf


WORKSPACE = Workspace()
