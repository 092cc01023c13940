"""CPU: the C-ABI library loads, exports every symbol declared in include/dpipe.h, and the ctypes layer declares
each of them.  No compute call is made."""
import ctypes

from diffusion_pipe_b200 import _abi, _lib


def test_library_loads_and_reports_version():
    lib = _lib.lib()
    assert lib.dpipe_abi_version() >= 1
    assert isinstance(lib.dpipe_last_error(), bytes)


def test_every_declared_symbol_is_exported():
    lib = _lib.lib()
    syms = _lib.exported_symbols()
    assert len(syms) >= 19
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f'declared in include/dpipe.h but not exported: {missing}'


def test_ctypes_signatures_cover_the_header():
    known = set(_abi.SIGNATURES) | {'dpipe_last_error', 'dpipe_abi_version', 'dpipe_check_device', 'dpipe_gemm_bf16'}
    assert set(_lib.exported_symbols()) <= known


def test_struct_sizes_match_c_layout():
    # a drifted struct would silently corrupt arguments: pin the sizes the C compiler produces (LP64)
    assert ctypes.sizeof(_lib.QkvEpilogue) == 11 * 8 + 4 * 4 + 4 + 4   # 11 pointers, 4 ints, float, tail padding
    assert ctypes.sizeof(_lib.GemmArgs) % 8 == 0
    assert ctypes.sizeof(_abi.AttnArgs) == 6 * 8 + 4 * 4 + 4 + 4
    assert ctypes.sizeof(_abi.AttnBwdArgs) == 12 * 8 + 4 * 4 + 4 + 4


def test_bad_arguments_fail_loudly_without_a_gpu():
    lib = _lib.lib()
    rc = lib.dpipe_gemm_bf16(None, None)
    assert rc < 0 and b'null' in lib.dpipe_last_error()
    assert lib.dpipe_sched_train(0, 1, 0, None, 0) < 0
