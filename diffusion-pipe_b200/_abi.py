"""ctypes signatures for the C-ABI entry points beyond the GEMM (kept in one place so that the CPU-side
test can check every symbol of include/dpipe.h is both exported and declared)."""
import ctypes

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_int64 = ctypes.c_int64
c_float = ctypes.c_float

SIGNATURES = {}


def declare(l):
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(l, name, None)
        if fn is None:
            continue
        fn.restype = restype
        fn.argtypes = argtypes


class AttnArgs(ctypes.Structure):
    _fields_ = [
        ('q', c_void_p), ('k', c_void_p), ('v', c_void_p),
        ('o', c_void_p), ('ldo', c_int64),
        ('lse', c_void_p),
        ('batch', c_int), ('heads', c_int), ('seq_q', c_int), ('seq_k', c_int),
        ('scale', c_float),
    ]


SIGNATURES['dpipe_attn_fwd'] = (c_int, [ctypes.POINTER(AttnArgs), c_void_p])


class AttnBwdArgs(ctypes.Structure):
    _fields_ = [
        ('q', c_void_p), ('k', c_void_p), ('v', c_void_p),
        ('o', c_void_p), ('ldo', c_int64),
        ('d_o', c_void_p), ('lddo', c_int64),
        ('lse', c_void_p), ('delta', c_void_p),
        ('dq', c_void_p), ('dk', c_void_p), ('dv', c_void_p),
        ('batch', c_int), ('heads', c_int), ('seq_q', c_int), ('seq_k', c_int),
        ('scale', c_float),
    ]


SIGNATURES['dpipe_attn_bwd'] = (c_int, [ctypes.POINTER(AttnBwdArgs), c_void_p])
