"""ctypes binding of libpqlb200.so (the C ABI declared in include/pqlb200.h).

There is no fallback: if the shared library is missing or a call fails, we raise.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PQLB_LIB") or os.path.join(_HERE, "libpqlb200.so")   # PQLB_LIB: kernel-variant experiments

MAX_GROUPS = 4
(EPI_STORE, EPI_BIAS, EPI_BIAS_ELU, EPI_BIAS_ELU_HEAD, EPI_BIAS_TANH, EPI_BIAS_TANH_NOISE,
 EPI_BIAS_SOFTMAX, EPI_MUL_ELUGRAD, EPI_MUL_TANHGRAD) = range(9)
K_MAJOR, MN_MAJOR = 0, 1

_f = C.c_void_p          # float* (device pointer)
_i64 = C.c_int64
_int = C.c_int
_flt = C.c_float
_st = C.c_void_p         # cudaStream_t


class GemmGroup(C.Structure):
    _fields_ = [("a", _f), ("lda", _i64), ("b", _f), ("ldb", _i64),
                ("a2", _f), ("lda2", _i64), ("b2", _f), ("ldb2", _i64),
                ("bias", _f), ("aux", _f), ("ldaux", _i64),
                ("head_w", _f), ("head_b", _f), ("q", _f),
                ("out", _f), ("ldo", _i64), ("out2", _f), ("ldo2", _i64),
                ("split_stride", _i64)]


MAX_COLSUM = 8


class ColsumDesc(C.Structure):
    _fields_ = [("n", _int), ("rows", _i64), ("dz", _f * MAX_COLSUM), ("ld", _i64 * MAX_COLSUM),
                ("n_cols", _int * MAX_COLSUM), ("part", _f * MAX_COLSUM)]


class GemmDesc(C.Structure):
    _fields_ = [("M", _int), ("N", _int), ("K", _int), ("K2", _int),
                ("a_major", _int), ("b_major", _int), ("epilogue", _int), ("tile_n", _int),
                ("splits", _int), ("n_groups", _int), ("col_lo", _int), ("col_hi", _int), ("cluster", _int),
                ("noise_bound", _flt), ("noise_std", _flt), ("g", GemmGroup * MAX_GROUPS)]


class MlpGroup(C.Structure):
    _fields_ = [("x", _f), ("ldx", _i64), ("w1", _f), ("ldw1", _i64), ("w2", _f), ("w3", _f),
                ("b1", _f), ("b2", _f), ("b3", _f), ("head_w", _f), ("head_b", _f), ("q", _f),
                ("h1", _f), ("h2", _f), ("h3", _f),
                ("act_w", _f), ("act_b", _f), ("act_noise", _f), ("act_out", _f), ("act_out2", _f),
                ("act_ldo", _i64), ("act_ldo2", _i64), ("act_ldnoise", _i64),
                ("noise_std", _flt), ("noise_bound", _flt), ("act_n", _int)]


MAX_FWD_GROUPS = 5


class MlpHGroup(C.Structure):
    _fields_ = [("x", _f), ("ldx", _i64), ("w1h", _f), ("w1l", _f), ("ldw1", _i64),
                ("w2h", _f), ("w2l", _f), ("w3h", _f), ("w3l", _f),
                ("b1", _f), ("b2", _f), ("b3", _f), ("head_w", _f), ("head_b", _f), ("q", _f),
                ("h1", _f), ("h2", _f), ("h3", _f),
                ("act_wh", _f), ("act_wl", _f), ("act_b", _f), ("act_noise", _f), ("act_out", _f), ("act_out2", _f),
                ("act_ldo", _i64), ("act_ldo2", _i64), ("act_ldnoise", _i64),
                ("noise_std", _flt), ("noise_bound", _flt), ("act_n", _int), ("terms", _int), ("k_in", _int),
                ("sm_wh", _f), ("sm_wl", _f), ("sm_b", _f), ("sm_out", _f), ("sm_ldp", _i64), ("sm_n", _int),
                ("publish", _int), ("wait", _int)]


class MlpHDesc(C.Structure):
    _fields_ = [("M", _int), ("k_in", _int), ("n_groups", _int), ("tile_sync", _f), ("g", MlpHGroup * MAX_FWD_GROUPS)]


class DpDesc(C.Structure):
    _fields_ = [("grad_peers", _f * 8), ("red_peers", _f * 8), ("ctl_peers", _f * 8), ("local", _f),
                ("rank", _int), ("world", _int), ("grid", _int), ("grad_mc", _f), ("red_mc", _f)]


class MlpBwdGroup(C.Structure):
    _fields_ = [("dz3", _f), ("w3", _f), ("w2", _f), ("h2", _f), ("h1", _f), ("dz2", _f), ("dz1", _f),
                ("bias_part2", _f), ("bias_part1", _f)]


class MlpBwdDesc(C.Structure):
    _fields_ = [("M", _int), ("n_groups", _int), ("g", MlpBwdGroup * MAX_GROUPS)]


MAX_WGRAD = 8


class WgradProblem(C.Structure):
    _fields_ = [("dz", _f), ("lddz", _i64), ("h", _f), ("ldh", _i64), ("part", _f), ("ldo", _i64),
                ("split_stride", _i64), ("M", _int), ("N", _int), ("tile_n", _int), ("splits", _int)]


class WgradDesc(C.Structure):
    _fields_ = [("K", _int), ("n_problems", _int), ("p", WgradProblem * MAX_WGRAD)]


class MlpDesc(C.Structure):
    _fields_ = [("M", _int), ("k_in", _int), ("n_groups", _int), ("g", MlpGroup * MAX_GROUPS)]


_PROTOS = {
    "pqlb_version": (_int, []),
    "pqlb_error_string": (C.c_char_p, [_int]),
    "pqlb_launch_count": (C.c_uint64, []),
    "pqlb_init": (_int, []),
    "pqlb_mma_peak": (_int, [_int, _int, _int, _st]),
    "pqlb_obs_pad": (_int, [_int]),
    "pqlb_record_ld": (_int, [_int, _int]),
    "pqlb_record_stride_mode": (None, [_int]),
    "pqlb_x_ld": (_int, [_int, _int]),
    "pqlb_ring_insert": (_int, [_f, _i64, _int, _int, _f, _f, _f, _f, _f, _i64, _i64, _st]),
    "pqlb_ring_insert_force_ldg": (None, [_int]),
    "pqlb_obsring_insert": (_int, [_f, _i64, _int, _f, _i64, _i64, _st]),
    "pqlb_nstep_push": (_int, [_f, _int, _int, _int, _int, _f, _f, _f, _f, _f, _int, _i64,
                               C.POINTER(_flt), _f, _f, _f, _f, _f, _st]),
    "pqlb_sample_gather": (_int, [_f, _i64, _int, _int, _f, _i64, _f, _f, _f, _f, _f, _st]),
    "pqlb_sample_critic_batch": (_int, [_f, _i64, _int, _int, _f, _i64, _f, _f, _flt, _f, _f, _int,
                                        _f, _f, _f, _f, _st]),
    "pqlb_sample_obs_batch": (_int, [_f, _i64, _int, _f, _i64, _f, _f, _flt, _f, _int, _int, _f, _st]),
    "pqlb_sample_critic_batch_rng": (_int, [_f, _i64, _int, _int, _f, _i64, _f, _f, _flt, _f, _f, _int,
                                            _f, _f, _f, _f, _f, _f, _i64, _f, _f, _st]),
    "pqlb_sample_obs_batch_rng": (_int, [_f, _i64, _int, _f, _i64, _f, _f, _flt, _f, _int, _int, _f, _f, _f, _f, _st]),
    "pqlb_store_i64": (_int, [_f, _i64, _st]),
    "pqlb_add_i64": (_int, [_f, _f, _i64, _st]),
    "pqlb_rms_workspace_bytes": (_i64, [_i64, _int]),
    "pqlb_rms_update": (_int, [_f, _i64, _int, _i64, _f, _f, _f, _f, _i64, _st]),
    "pqlb_rms_moments": (_int, [_f, _i64, _int, _i64, _f, _f, _i64, _st]),
    "pqlb_rms_apply": (_int, [_f, _i64, _int, _f, _f, _f, _st]),
    "pqlb_actor_inputs": (_int, [_f, _i64, _int, _i64, _f, _f, _flt, _int, _int, _f, _int, _f, _int, _f, _flt,
                                 _i64, _i64, _st]),
    "pqlb_env_post": (_int, [_f, _f, _f, _flt, _int, _f, _f, _f, _f, _int, _f, _f, _f, _st]),
    "pqlb_gemm_tf32": (_int, [C.POINTER(GemmDesc), _st]),
    "pqlb_mlp_forward": (_int, [C.POINTER(MlpDesc), _st]),
    "pqlb_mlp_forward_cluster": (None, [_int]),
    "pqlb_mlp_forward_h": (_int, [C.POINTER(MlpHDesc), _st]),
    "pqlb_mlp_forward_h_mode": (None, [_int]),
    "pqlb_mlp_forward_h_debug": (None, [_f]),
    "pqlb_split_f16": (_int, [_f, _f, _f, _i64, _st]),
    "pqlb_f16_weight_scale": (_flt, []),
    "pqlb_mlp_backward": (_int, [C.POINTER(MlpBwdDesc), _st]),
    "pqlb_wgrad_multi": (_int, [C.POINTER(WgradDesc), _st]),
    "pqlb_round_tf32": (_int, [_f, _f, _i64, _st]),
    "pqlb_doubleq_td_loss": (_int, [_f, _f, _f, _f, _f, _f, _flt, _i64, _f, _f, _f, _f, _f, _f, _f,
                                    _f, _f, _f, _st]),
    "pqlb_doubleq_td_loss_b3": (_int, [_f, _f, _f, _f, _f, _f, _flt, _i64, _f, _f, _f, _f, _f, _f, _f,
                                       _f, _f, _f, _f, _f, _st]),
    "pqlb_dpg_loss": (_int, [_f, _f, _i64, _f, _f, _f, _f, _f, _f, _f, _st]),
    "pqlb_c51_td_loss": (_int, [_f, _f, _f, _f, _int, _f, _f, _f, _flt, _flt, _flt, _int, _i64,
                                _f, _f, _f, _int, _f, _st]),
    "pqlb_colsum_partial": (_int, [_f, _i64, _i64, _int, _f, _st]),
    "pqlb_colsum_partial_multi": (_int, [C.POINTER(ColsumDesc), _st]),
    "pqlb_c51_dpg_loss": (_int, [_f, _f, _int, _f, _int, _i64, _f, _f, _int, _f, _f, _st]),
    "pqlb_pack_x": (_int, [_f, _i64, _int, _f, _i64, _int, _f, _int, _i64, _st]),
    "pqlb_grad_reduce": (_int, [_f, _int, _f, _f, _f, _st]),
    "pqlb_grad_sumsq": (_int, [_f, _int, _f, _f, _st]),
    "pqlb_adamw_polyak": (_int, [_f, _f, _f, _f, _f, _f, _f, _f, _f, _i64, _f, _int, _flt, _flt, _flt, _flt,
                                 _flt, _flt, _flt, _i64, _f, _flt, _f, _st]),
    "pqlb_grad_reduce_finish": (_int, [_f, _int, _f, _f, _f, _f, _int, _flt, _f, _f, _f, _int,
                                       _flt, _flt, _flt, _flt, _flt, _flt, _f, _st]),
    "pqlb_adamw_polyak_pre": (_int, [_f, _f, _f, _f, _f, _f, _f, _f, _f, _i64, _f, _int, _flt, _flt, _f, _f, _f, _st]),
    "pqlb_adamw_polyak_dp": (_int, [_f, _f, _f, _f, _f, _f, _f, _f, _i64, C.POINTER(DpDesc), _flt, _f, _f, _f, _st]),
    "pqlb_dp_spin_limit": (_int, [C.c_double]),
    "pqlb_grad_exchange_dp": (_int, [_i64, C.POINTER(DpDesc), _st]),
    "pqlb_sum_partials": (_int, [_f, _int, _flt, _f, _f, _f, _int, _st]),
}

_lib = None


def launch_count():
    """CUDA kernels launched by libpqlb200.so in this process (bench.py reports the delta)."""
    return int(load().pqlb_launch_count())


def exported_symbols():
    return sorted(_PROTOS)


def load():
    """Load the shared library once; raise loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "pql_b200 has no CPU or PyTorch fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if os.environ.get("PQLB_REC_STRIDE"):        # record-stride experiment: "128" = 128-byte aligned instead of a power of two
            lib.pqlb_record_stride_mode(1 if os.environ["PQLB_REC_STRIDE"] == "128" else 0)
        if os.environ.get("PQLB_FWD_H_MODE"):       # kernel-schedule experiments: 1 = CTA per tile, 2 = persistent
            lib.pqlb_mlp_forward_h_mode(int(os.environ["PQLB_FWD_H_MODE"]))
        _lib = lib
    return _lib


class PqlbError(RuntimeError):
    pass


def check(rc, what):
    if rc != 0:
        msg = load().pqlb_error_string(rc).decode()
        if rc in (-1, -2, -3):
            raise ValueError(f"{what}: {msg} (code {rc})")
        raise PqlbError(f"{what}: {msg} (code {rc})")


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def call(name, *args):
    """Invoke one C-ABI entry point on torch's current stream and raise on failure."""
    fn = getattr(load(), name)
    rc = fn(*args, stream())
    check(rc, name)


def require_cuda(t, name="tensor"):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must live on a CUDA device: pql_b200 has no CPU path")
    return t
