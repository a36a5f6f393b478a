"""ctypes binding of ``include/svgpfa_b200.h`` (the only way the package reaches the GPU).

There is deliberately no fallback: if ``libsvgpfa_b200.so`` is missing or fails to load,
importing :func:`lib` raises.
"""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libsvgpfa_b200.so")
PROBES_LIB_PATH = os.path.join(PKG, "libsvgpfa_b200_probes.so")      # measurement probes / test hooks, not product

ABI_VERSION = 8
MAX_M = 64
EMBED_TN = 128
SHARED_HDR = 8
TERM1_SLOTS = 4096
FIN_SLOTS = 1024
SHARED_STATUS = 5            # shared[5..7] = (status, trial, latent) of a failed Cholesky, as doubles
KERNEL_EXPQUAD, KERNEL_PERIODIC = 0, 1
GRAD_POSTERIOR, GRAD_EMBEDDING, GRAD_KERNEL, GRAD_INDLOCS = 1, 2, 4, 8
GRAD_ALL = 15
REUSE_KZZ, REUSE_SPIKE, REBUILD_PANELS, REUSE_VQ = 16, 32, 64, 128
SPIKE_DIRECT, SPIKE_PANEL = 1, 2
PM_P = 16
INFO_NOT_PD = 1


class LatentDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("ktype", "M", "moff", "mmoff", "poff", "thoff", "P", "nth")]


class Dims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("R", "N", "K", "Q", "KM", "MM", "PP", "TH", "Mmax", "n_ntiles")] + [
        ("S", C.c_int64), ("reg", C.c_double), ("desc_host", C.POINTER(LatentDesc)), ("r0", C.c_int32), ("rn", C.c_int32),
        ("spike_chunks", C.c_int32), ("quad_warps", C.c_int32), ("spike_method", C.c_int32), ("pm_B", C.c_int32),
        ("pm_lo", C.c_double), ("pm_w", C.c_double)]


BUFFER_FIELDS = (
    "desc", "kscale", "theta", "Z", "m", "cholvec", "C", "d", "tq", "wq", "spike_t", "seg_off", "spike_cnt",
    "L", "Li", "X", "c", "alpha", "logdetL", "kl_rk", "A_q", "abar_q", "abar_spk", "dz_acc", "dth_part",
    "mu_q", "var_q", "v_q", "mubar_part", "varbar_part", "term1_part", "fin_part", "pm_tau", "pm_mun", "pm_mt", "mu_s", "gsum",
    "shared", "gZ", "gm", "gcholvec", "info")


class Buffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in BUFFER_FIELDS]


HOST_IO_FIELDS = ("theta_host", "Z_host", "m_host", "cholvec_host", "C_host", "d_host", "tq_host", "wq_host",
                  "spike_t_host", "seg_off_host", "spike_cnt_host",
                  "shared_host", "gZ_host", "gm_host", "gcholvec_host", "info_host")


class HostIO(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in HOST_IO_FIELDS] + [("copy_static", C.c_int32), ("n_blocks", C.c_int32)]


# every symbol include/svgpfa_b200.h declares: name -> (restype, argtypes)
_P = C.POINTER
SYMBOLS = {
    "svgpfa_abi_version": (C.c_int, []),
    "svgpfa_last_error": (C.c_char_p, []),
    "svgpfa_kzz_chol_fwd": (C.c_int, [_P(Dims), _P(Buffers), C.c_void_p]),
    "svgpfa_indpoints_fwd": (C.c_int, [_P(Dims), _P(Buffers), C.c_void_p]),
    "svgpfa_quad_latent_fwd": (C.c_int, [_P(Dims), _P(Buffers), C.c_void_p]),
    "svgpfa_quad_latent_fwd_cached": (C.c_int, [_P(Dims), _P(Buffers), C.c_void_p]),
    "svgpfa_quad_embed_fwd_bwd": (C.c_int, [_P(Dims), _P(Buffers), C.c_uint32, C.c_void_p]),
    "svgpfa_quad_latent_bwd": (C.c_int, [_P(Dims), _P(Buffers), C.c_uint32, C.c_void_p]),
    "svgpfa_spike_fwd_bwd": (C.c_int, [_P(Dims), _P(Buffers), C.c_uint32, C.c_void_p]),
    "svgpfa_panel_moments": (C.c_int, [_P(Dims), _P(Buffers), C.c_void_p]),
    "svgpfa_panel_neuron_sums": (C.c_int, [_P(Dims), _P(Buffers), C.c_void_p]),
    "svgpfa_spike_panel_fwd_bwd": (C.c_int, [_P(Dims), _P(Buffers), C.c_uint32, C.c_void_p]),
    "svgpfa_indpoints_bwd": (C.c_int, [_P(Dims), _P(Buffers), C.c_uint32, C.c_void_p]),
    "svgpfa_finalize": (C.c_int, [_P(Dims), _P(Buffers), C.c_uint32, C.c_void_p]),
    "svgpfa_elbo_grad": (C.c_int, [_P(Dims), _P(Buffers), C.c_uint32, C.c_void_p]),
    "svgpfa_spike_latent_means": (C.c_int, [_P(Dims), _P(Buffers), C.c_void_p]),
    "svgpfa_cached_ell_fwd_bwd": (C.c_int, [_P(Dims), _P(Buffers), C.c_uint32, C.c_void_p]),
    "svgpfa_embed_predict": (C.c_int, [_P(Dims), _P(Buffers), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "svgpfa_build_segments_host": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "svgpfa_elbo_grad_host": (C.c_int, [_P(Dims), _P(Buffers), _P(HostIO), C.c_uint32, C.c_void_p]),
    "svgpfa_set_stage_events": (C.c_int, [C.c_void_p]),
    "svgpfa_release_thread_resources": (C.c_int, []),
    "svgpfa_lbfgs_ws_doubles": (C.c_uint64, []),
    "svgpfa_lbfgs_multidot": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_uint64, C.c_void_p, C.c_void_p,
                                        C.c_void_p]),
    "svgpfa_lbfgs_combine": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_uint64,
                                       C.c_void_p, C.c_void_p, C.c_void_p]),
    "svgpfa_lbfgs_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "svgpfa_lbfgs_update": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_uint64,
                                      C.c_void_p]),
    "svgpfa_lbfgs_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_uint64, C.c_void_p]),
}
# include/svgpfa_b200_probes.h (separate library)
PROBE_SYMBOLS = {
    "svgpfa_probes_last_error": (C.c_char_p, []),
    "svgpfa_peak_probe": (C.c_int, [C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
    "svgpfa_exp_neg_eval": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "svgpfa_exp2m_eval": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
}
STAGES = ("kzz_chol", "indpoints_fwd", "quad_latent_fwd", "quad_embed", "quad_latent_bwd", "spike_fwd_bwd",
          "indpoints_bwd", "finalize")

_lib = None


def lib():
    """The loaded library; raises (never falls back) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA library is not built "
                "(run `python -m svgpfa_b200.build`); there is no CPU fallback")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.svgpfa_abi_version() != ABI_VERSION:
            raise RuntimeError("libsvgpfa_b200.so ABI version mismatch; rebuild")
        _lib = handle
    return _lib


_probes = None


def probes():
    """The probes / test-hooks library (bench.py roofline leg, tools/, tests/ only)."""
    global _probes
    if _probes is None:
        if not os.path.exists(PROBES_LIB_PATH):
            raise RuntimeError(f"{PROBES_LIB_PATH} not found (run `python -m svgpfa_b200.build`)")
        handle = C.CDLL(PROBES_LIB_PATH)
        for name, (res, args) in PROBE_SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _probes = handle
    return _probes


def check_probe(rc: int, what: str = ""):
    if rc != 0:
        raise RuntimeError(f"probe call failed ({rc}) {what}: {probes().svgpfa_probes_last_error().decode()}")


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().svgpfa_last_error().decode()
        raise RuntimeError(f"svgpfa_b200 C-ABI call failed ({rc}) {what}: {msg}")
