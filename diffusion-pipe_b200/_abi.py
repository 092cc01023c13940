"""ctypes structures and signatures for the C-ABI entry points beyond the GEMM (kept in one place so that the
CPU-side test can check every symbol of include/dpipe.h is both exported and declared)."""
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


class QkBwdArgs(ctypes.Structure):
    _fields_ = [
        ('dq', c_void_p), ('dk', c_void_p), ('dv', c_void_p),
        ('qhat', c_void_p), ('khat', c_void_p),
        ('q_rstd', c_void_p), ('k_rstd', c_void_p),
        ('q_norm_w', c_void_p), ('k_norm_w', c_void_p),
        ('rope_cos', c_void_p), ('rope_sin', c_void_p),
        ('dqkv', c_void_p), ('ld', c_int64),
        ('dbias', c_void_p), ('dw', c_void_p),
        ('batch', c_int), ('heads', c_int), ('seq_total', c_int), ('seq_offset', c_int), ('rows_per_batch', c_int),
    ]


class WanNormProj(ctypes.Structure):
    _fields_ = [('src', c_void_p), ('ld', c_int64), ('weight', c_void_p), ('dst', c_void_p), ('xhat', c_void_p),
                ('rstd', c_void_p), ('rope', c_int)]


class WanNormFwdArgs(ctypes.Structure):
    _fields_ = [('proj', WanNormProj * 3), ('nproj', c_int), ('cos', c_void_p), ('sin', c_void_p),
                ('batch', c_int), ('seq', c_int), ('heads', c_int), ('eps', c_float)]


class WanNormBwdProj(ctypes.Structure):
    _fields_ = [('dy', c_void_p), ('xhat', c_void_p), ('rstd', c_void_p), ('weight', c_void_p), ('dx', c_void_p),
                ('ld', c_int64), ('dw_partials', c_void_p), ('rope', c_int)]


class WanNormBwdArgs(ctypes.Structure):
    _fields_ = [('proj', WanNormBwdProj * 3), ('nproj', c_int), ('cos', c_void_p), ('sin', c_void_p),
                ('batch', c_int), ('seq', c_int), ('heads', c_int)]


SIGNATURES['dpipe_wan_norm_rope_fwd'] = (c_int, [ctypes.POINTER(WanNormFwdArgs), c_void_p])
SIGNATURES['dpipe_wan_norm_rope_bwd'] = (c_int, [ctypes.POINTER(WanNormBwdArgs), c_void_p])
SIGNATURES['dpipe_wan_norm_rows'] = (c_int, [])
SIGNATURES['dpipe_attn_fwd'] = (c_int, [ctypes.POINTER(AttnArgs), c_void_p])
SIGNATURES['dpipe_attn_bwd'] = (c_int, [ctypes.POINTER(AttnBwdArgs), c_void_p])
SIGNATURES['dpipe_row_chunk'] = (c_int, [])
SIGNATURES['dpipe_ln_modulate_fwd'] = (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                                               c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p])
SIGNATURES['dpipe_ln_modulate_bwd'] = (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                               c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int, c_int,
                                               c_int, c_void_p])
SIGNATURES['dpipe_ln_modulate_fwd_ex'] = (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                                                  c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_int, c_void_p])
SIGNATURES['dpipe_ln_modulate_bwd_ex'] = (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                                  c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int, c_int,
                                                  c_int, c_int, c_void_p])
SIGNATURES['dpipe_gate_bwd'] = (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                                        c_void_p, c_int, c_int, c_int, c_void_p])
SIGNATURES['dpipe_colreduce_finish'] = (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int64, c_void_p,
                                                c_int64, c_void_p, c_void_p, c_void_p])
SIGNATURES['dpipe_colsum_chunks'] = (c_int, [c_int])
SIGNATURES['dpipe_colsum'] = (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p])
SIGNATURES['dpipe_qknorm_rope_bwd'] = (c_int, [ctypes.POINTER(QkBwdArgs), c_void_p])
SIGNATURES['dpipe_sched_num_pipe_buffers'] = (c_int, [c_int, c_int, c_int])
SIGNATURES['dpipe_sched_train'] = (c_int, [c_int, c_int, c_int, c_void_p, c_int])
SIGNATURES['dpipe_sched_infer'] = (c_int, [c_int, c_int, c_int, c_void_p, c_int])
SIGNATURES['dpipe_sched_zb'] = (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int])
SIGNATURES['dpipe_sched_zb_makespan'] = (ctypes.c_longlong, [c_int, c_int, c_int, c_int, c_int, c_int])
SIGNATURES['dpipe_sched_zb_ex'] = (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int])
SIGNATURES['dpipe_sched_zb_makespan_ex'] = (ctypes.c_longlong, [c_int, c_int, c_int, c_int, c_int, c_int, c_void_p])
SIGNATURES['dpipe_partition_balanced'] = (c_int, [c_void_p, c_int, c_int, c_void_p])
SIGNATURES['dpipe_ipc_alloc'] = (c_int, [c_int64, ctypes.POINTER(c_void_p), c_void_p])
SIGNATURES['dpipe_ipc_open'] = (c_int, [c_void_p, ctypes.POINTER(c_void_p)])
SIGNATURES['dpipe_ipc_close'] = (c_int, [c_void_p])
SIGNATURES['dpipe_ipc_free'] = (c_int, [c_void_p])
SIGNATURES['dpipe_peer_copy'] = (c_int, [c_void_p, c_int, c_void_p, c_int, c_int64, c_void_p])
SIGNATURES['dpipe_flag_write'] = (c_int, [c_void_p, ctypes.c_uint64, c_void_p])
SIGNATURES['dpipe_flag_wait_geq'] = (c_int, [c_void_p, ctypes.c_uint64, ctypes.c_double, c_void_p])
SIGNATURES['dpipe_mod_fwd'] = (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p])
SIGNATURES['dpipe_mod_bwd_chunks'] = (c_int, [c_int])
SIGNATURES['dpipe_mod_bwd'] = (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                        c_int, c_int, c_int, c_void_p])
SIGNATURES['dpipe_mse_loss'] =(c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p])
SIGNATURES['dpipe_fp8_to_bf16'] = (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p])
SIGNATURES['dpipe_fp8_code_table'] = (c_int, [c_int, c_void_p])
SIGNATURES['dpipe_grad_sumsq_blocks'] = (c_int, [])
SIGNATURES['dpipe_grad_sumsq'] = (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int64, c_void_p, c_void_p])
SIGNATURES['dpipe_grad_scale'] = (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p])
SIGNATURES['dpipe_noise_pack'] = (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int64, c_int, c_int,
                                          c_int, c_void_p])
SIGNATURES['dpipe_exec_create'] = (c_int, [c_int, c_int, ctypes.c_double, ctypes.POINTER(c_void_p)])
SIGNATURES['dpipe_exec_destroy'] = (c_int, [c_void_p])
SIGNATURES['dpipe_exec_copy_stream'] = (c_void_p, [c_void_p])
SIGNATURES['dpipe_exec_bind'] = (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64, c_int64, c_int])
SIGNATURES['dpipe_exec_set_layout'] = (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p])
SIGNATURES['dpipe_exec_forget_layouts'] = (c_int, [c_void_p])
SIGNATURES['dpipe_exec_load_plan'] = (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int])
SIGNATURES['dpipe_exec_stage_send'] = (c_int, [c_void_p, c_int, c_int, c_int, c_void_p])
SIGNATURES['dpipe_exec_recv_base'] = (c_int, [c_void_p, c_int, c_int, ctypes.POINTER(c_void_p)])
SIGNATURES['dpipe_exec_next'] = (c_int, [c_void_p, c_void_p, c_void_p])
