"""ctypes binding of libpcnbr.so (include/pcnbr.h).  No torch types cross this boundary: only raw
device pointers, sizes and the current CUDA stream handle.

The product path has NO CPU fallback: if the library is missing, or a tensor is not a CUDA tensor,
the call raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_long, c_longlong, c_size_t, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libpcnbr.so")
ABI_VERSION = 1

_P, _I, _L, _F, _Z, _D = c_void_p, c_int, c_long, c_float, c_size_t, c_double

# name -> (restype, argtypes): mirrors include/pcnbr.h one to one
PROTOTYPES = {
    "pcnbr_abi_version": (_I, []),
    "pcnbr_error_string": (c_char_p, [_I]),
    "pcnbr_fps_ws_bytes": (_Z, [_I, _I]),
    "pcnbr_fps_f32": (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _Z, _P]),
    "pcnbr_fps_len_f32": (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _P, _Z, _P]),
    "pcnbr_ball_query_len_f32": (_I, [_P, _P, _I, _I, _I, _F, _I, _P, _P, _P, _P, _Z, _P]),
    "pcnbr_knn_direct_len_f32": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _Z, _P]),
    "pcnbr_ball_query_multi_len_f32": (_I, [_P, _P, _I, _I, _I, _P, _P, _I, _P, _P, _P, _Z, _P]),
    "pcnbr_knn_expand_len_f32": (_I, [_P, _I, _I, _I, _L, _L, _I, _P, _P, _P, _Z, _P]),
    "pcnbr_ball_query_f32": (_I, [_P, _P, _I, _I, _I, _F, _I, _P, _P]),
    "pcnbr_knn_direct_f32": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "pcnbr_grid_ws_bytes": (_Z, [_I, _I]),
    "pcnbr_ball_query_grid_f32": (_I, [_P, _P, _I, _I, _I, _F, _I, _P, _P, _Z, _P]),
    "pcnbr_knn_direct_grid_f32": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _Z, _P]),
    "pcnbr_ball_query_multi_ws_bytes": (_Z, [_I, _I, _I]),
    "pcnbr_ball_query_multi_f32": (_I, [_P, _P, _I, _I, _I, _P, _P, _I, _P, _P, _Z, _P]),
    "pcnbr_knn_expand_ws_bytes": (_Z, [_I, _I, _I, _I]),
    "pcnbr_knn_expand_f32": (_I, [_P, _I, _I, _I, _L, _L, _I, _P, _P, _Z, _P]),
    "pcnbr_knn_tc_debug_f32": (_I, [_P, _I, _I, _I, _L, _L, _I, _P, _P, _Z, _P, _P, _P]),
    "pcnbr_group_f32": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P, _I, _P]),
    "pcnbr_gather_rows_f32": (_I, [_P, _P, _I, _I, _L, _I, _P, _P]),
    "pcnbr_gather_rows_bwd_f32": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "pcnbr_square_distance_f32": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "pcnbr_csr_ws_bytes": (_Z, [_I, _I, _I]),
    "pcnbr_csr_build": (_I, [_P, _I, _I, _I, _P, _P, _P, _Z, _P]),
    "pcnbr_csr_rows_ws_bytes": (_Z, [_I, _I, _I, _I]),
    "pcnbr_csr_build_rows": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _Z, _P]),
    "pcnbr_group_bwd_f32": (_I, [_P, _I, _P, _P, _I, _I, _I, _I, _P, _P]),
    "pcnbr_maxpool_f32": (_I, [_P, _L, _I, _I, _L, _L, _L, _P, _P, _P]),
    "pcnbr_maxpool_bwd_f32": (_I, [_P, _P, _L, _I, _I, _L, _L, _L, _P, _P]),
    "pcnbr_interp_f32": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P]),
    "pcnbr_interp_bwd_f32": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "pcnbr_edge_feature_f32": (_I, [_P, _P, _I, _I, _I, _I, _P, _P]),
    "pcnbr_edge_feature_bwd_f32": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "pcnbr_edgeconv_fwd_blocks": (_I, [_I]),
    "pcnbr_edgeconv_fwd_f32": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "pcnbr_edgeconv_bwd_f32": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "pcnbr_bn_supported": (_I, [_L, _I]),
    "pcnbr_bn_blocks": (_I, [_L, _I]),
    "pcnbr_bn_stats_f32": (_I, [_P, _L, _I, _P, _P]),
    "pcnbr_bn_finalize_f32": (_I, [_P, _I, _P, _D, _I, _P, _P, _F, _F, _P, _P, _P, _P]),
    "pcnbr_bn_act_fwd_f32": (_I, [_P, _L, _P, _L, _L, _I, _P, _F, _P, _P, _F, _P, _P]),
    "pcnbr_bn_act_bwd_reduce_f32": (_I, [_P, _P, _L, _P, _L, _L, _I, _P, _F, _P, _P, _P, _F, _P]),
    "pcnbr_bn_bwd_finalize_f32": (_I, [_P, _I, _P, _D, _I, _I, _P, _P, _P, _P]),
    "pcnbr_bn_act_bwd_apply_f32": (_I, [_P, _P, _L, _I, _P, _P, _F, _P, _P, _F, _P, _P]),
    "pcnbr_pool_bn_act_fwd_f32": (_I, [_P, _L, _I, _I, _P, _F, _P, _P, _P, _P]),
    "pcnbr_pool_bn_bwd_apply_f32": (_I, [_P, _P, _P, _L, _I, _I, _P, _P, _P]),
    "pcnbr_gemm3x_splits": (_I, [_I, _I, _I]),
    "pcnbr_gemm3x_ws_bytes": (_Z, [_I, _I, _I, _I]),
    "pcnbr_gemm3x_f32": (_I, [_P, _L, _I, _P, _L, _I, _I, _I, _I, _P, _P, _I, _P, _Z, _P]),
    "pcnbr_gemm3x_ex_f32": (_I, [_P, _L, _I, _P, _L, _I, _P, _L, _I, _I, _I, _I, _P, _P, _L, _I, _P, _Z, _P]),
    "pcnbr_amax_slots": (_I, []),
    "pcnbr_absmax_f32": (_I, [_P, _L, _L, _L, _P, _P]),
    "pcnbr_gemm2h_preferred": (_I, [_I, _I, _I]),
    "pcnbr_split_f16": (_I, [_P, _I, _I, _L, _I, _P, _P, _L, _L, _P]),
    "pcnbr_gemm2h_ex_f32": (_I, [_P, _L, _I, _P, _L, _I, _P, _L, _I, _I, _I, _I, _P, _P, _L, _I, _P, _Z, _P, _P, _P, _P, _L, _L, _P]),
    "pcnbr_gemm2h_ex2_f32": (_I, [_P, _L, _I, _P, _L, _I, _P, _L, _I, _I, _I, _I, _P, _P, _L, _I, _P, _Z, _P, _P, _P, _P, _L, _L,
                                  _P, _L, _L, _P, _L, _L, _P, _L, _L, _P, _L, _L, _P]),
    "pcnbr_gemm2h_trace": (_I, [_P]),
    "pcnbr_confusion_f32": (_I, [_P, _P, _P, _I, _I, _I, _P, _P]),
    "pcnbr_confusion_ex_f32": (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _P]),
    "pcnbr_masked_ce_blocks": (_I, [_I]),
    "pcnbr_masked_ce_f32": (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _P, _P]),
    "pcnbr_block_batch": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P]),
    "pcnbr_window_merge_f32": (_I, [_P, _P, _I, c_longlong, _I, _I, _I, _P, _P, _P, _P]),
    "pcnbr_prof_enable": (None, [_I]),
    "pcnbr_prof_collect": (_I, [_P, _Z]),
}

# CUDA kernels launched per C-ABI call (csr_build = count + scan + fill + sort; knn_expand = sumsq + select)
KERNELS_PER_CALL = {"pcnbr_ball_query_multi_f32": 2, "pcnbr_masked_ce_f32": 2, "pcnbr_ball_query_grid_f32": 5, "pcnbr_knn_direct_grid_f32": 5, "pcnbr_csr_build": 4, "pcnbr_csr_build_rows": 3, "pcnbr_knn_expand_f32": 5, "pcnbr_knn_expand_len_f32": 5, "pcnbr_ball_query_len_f32": 5, "pcnbr_knn_direct_len_f32": 5, "pcnbr_ball_query_multi_len_f32": 2, "pcnbr_knn_tc_debug_f32": 5}

_lib = None
launches = 0          # number of libpcnbr CUDA kernels launched by this process (bench.py reports it)
timing = None         # None, or {name: [(start_event, end_event), ...]} while bench.py profiles a region


class PcnbrError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """dlopen libpcnbr.so and bind every prototype; raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PcnbrError(
            f"{LIB_PATH} is missing: the CUDA library has not been built "
            "(run `python __graft_entry__.py build`). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if a declared symbol is not exported
        fn.restype, fn.argtypes = res, args
    if lib.pcnbr_abi_version() != ABI_VERSION:
        raise PcnbrError(f"libpcnbr ABI {lib.pcnbr_abi_version()} != host layer ABI {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def call(name: str, *args, tag: str = "") -> None:
    """Invoke a compute entry point and turn a non-zero return code into an exception."""
    global launches
    lib = load()
    if timing is not None:
        import torch
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()                      # on the current stream = the stream handed to the kernel
        rc = getattr(lib, name)(*args)
        ev1.record()
        timing.setdefault(name + tag, []).append((ev0, ev1))
    else:
        rc = getattr(lib, name)(*args)
    launches += KERNELS_PER_CALL.get(name, 1)
    if rc != 0:
        raise PcnbrError(f"{name} failed: {lib.pcnbr_error_string(rc).decode()} (code {rc})")


def size(name: str, *args) -> int:
    return int(getattr(load(), name)(*args))


def start_timing() -> None:
    """Bracket every C-ABI call with CUDA events on the launching stream (bench.py roofline)."""
    global timing
    timing = {}


def stop_timing() -> dict:
    """-> {entry point: (calls, total ms)}; synchronises the device."""
    global timing
    import torch
    torch.cuda.synchronize()
    out = {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in (timing or {}).items()}
    timing = None
    return out


def prof_enable(on: bool = True) -> None:
    """Per-KERNEL timing inside libpcnbr (csrc/prof.cu): CUDA events on the launch stream around every kernel,
    plus the launch's algorithmic bytes / flops.  Keep it off while capturing a CUDA graph."""
    load().pcnbr_prof_enable(1 if on else 0)


def prof_collect() -> dict:
    """-> {kernel: {"calls", "ms", "bytes", "flops"}} summed over the launches since the last collect; synchronises."""
    cap = 1 << 22
    buf = ctypes.create_string_buffer(cap)
    load().pcnbr_prof_collect(ctypes.cast(buf, c_void_p), cap)
    out: dict = {}
    for line in buf.value.decode().splitlines():
        name, ms, nbytes, flops = line.split("\t")
        d = out.setdefault(name, {"calls": 0, "ms": 0.0, "bytes": 0.0, "flops": 0.0})
        d["calls"] += 1
        d["ms"] += float(ms)
        d["bytes"] += float(nbytes)
        d["flops"] += float(flops)
    return out
