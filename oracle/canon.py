"""ctypes front-end of oracle/canon.c  -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package never does (it fails loudly without its CUDA library).

All functions take/return CPU torch tensors (fp32 / int32) and follow the canonical tie rule
(lowest index wins).  Reference lines are cited in canon.c next to each function.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liborc_canon.so")
_SRC = os.path.join(_HERE, "canon.c")
_lib = None


def build(force: bool = False) -> str:
    """gcc -O2 -fopenmp -ffp-contract=off canon.c -> oracle/_build/liborc_canon.so (host threads over independent
    clouds / query rows only: every result is computed by one thread in the reference's order)"""
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        # -march=x86-64-v2 keeps the .so runnable on any host the snapshot travels to; fmaf()
        # then resolves to glibc's exact software/hardware-dispatched fma.
        cmd = ["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
               "-fvisibility=hidden", "-o", _SO, _SRC, "-lm"]
        subprocess.run(cmd, check=True)
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _fp(t):
    assert t.dtype == torch.float32 and t.is_contiguous() and t.device.type == "cpu"
    return ctypes.c_void_p(t.data_ptr())


def _ip(t):
    assert t.dtype == torch.int32 and t.is_contiguous() and t.device.type == "cpu"
    return ctypes.c_void_p(t.data_ptr())


def _f32(t):
    return t.detach().to("cpu", torch.float32).contiguous()


def fps(xyz: torch.Tensor, C: int, start: torch.Tensor):
    """-> (idx (B,C) int32, coords (B,C,3))   [models/utils/common.py:6-34]"""
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    start = start.detach().to("cpu", torch.int32).contiguous()
    idx = torch.empty(B, C, dtype=torch.int32)
    out = torch.empty(B, C, 3, dtype=torch.float32)
    _load().orc_fps(_fp(xyz), B, N, C, _ip(start), _ip(idx), _fp(out))
    return idx, out


def ball_query(q: torch.Tensor, p: torch.Tensor, r: float, K: int) -> torch.Tensor:
    """-> idx (B,M,K) int32   [common.py:54-61]"""
    q, p = _f32(q), _f32(p)
    B, M, _ = q.shape
    N = p.shape[1]
    r2 = torch.tensor(r ** 2, dtype=torch.float32).item()
    idx = torch.empty(B, M, K, dtype=torch.int32)
    _load().orc_ball_query(_fp(q), _fp(p), B, M, N, ctypes.c_float(r2), K, _ip(idx))
    return idx


def knn_direct(q: torch.Tensor, p: torch.Tensor, K: int):
    """-> (idx (B,M,K) int32, d2 (B,M,K))   [common.py:110-114]"""
    q, p = _f32(q), _f32(p)
    B, M, _ = q.shape
    N = p.shape[1]
    idx = torch.empty(B, M, K, dtype=torch.int32)
    d2 = torch.empty(B, M, K, dtype=torch.float32)
    _load().orc_knn_direct(_fp(q), _fp(p), B, M, N, K, _ip(idx), _fp(d2))
    return idx, d2


def sumsq(x: torch.Tensor) -> torch.Tensor:
    """x (B,F,N) -> (B,N): torch.sum(x**2, dim=1) in ATen's summation order  [dgcnn.py:17]"""
    x = _f32(x)
    B, F, N = x.shape
    xx = torch.empty(B, N, dtype=torch.float32)
    _load().orc_sumsq(_fp(x), B, F, N, _fp(xx))
    return xx


def knn_expand(x: torch.Tensor, K: int):
    """x (B,F,N) -> (idx (B,N,K) int32, pd (B,N,K))   [models/dgcnn/dgcnn.py:7-21]"""
    x = _f32(x)
    B, F, N = x.shape
    idx = torch.empty(B, N, K, dtype=torch.int32)
    pd = torch.empty(B, N, K, dtype=torch.float32)
    _load().orc_knn_expand(_fp(x), B, F, N, K, _ip(idx), _fp(pd))
    return idx, pd


def knn_expand_rows(x: torch.Tensor, K: int, rows: torch.Tensor) -> torch.Tensor:
    """x (F,N) one cloud, rows (R,) query indices -> idx (R,K) int32: knn_expand restricted to those rows
    [models/dgcnn/dgcnn.py:7-21]"""
    x = _f32(x)
    F, N = x.shape
    rows = rows.detach().to("cpu", torch.int32).contiguous()
    idx = torch.empty(rows.numel(), K, dtype=torch.int32)
    _load().orc_knn_expand_rows(_fp(x), F, N, K, _ip(rows), rows.numel(), _ip(idx))
    return idx


def group(q, p, feat, idx, r: float, normalize: bool) -> torch.Tensor:
    """-> (B,M,K,3+D)   [common.py:62-71]"""
    q, p, feat = _f32(q), _f32(p), _f32(feat)
    idx = idx.detach().to("cpu", torch.int32).contiguous()
    B, M, K = idx.shape
    N, D = p.shape[1], feat.shape[2]
    out = torch.empty(B, M, K, 3 + D, dtype=torch.float32)
    rdiv = torch.tensor(r, dtype=torch.float32).item() if normalize else 0.0
    _load().orc_group(_fp(p), _fp(feat), _fp(q), _ip(idx), B, N, M, K, D, ctypes.c_float(rdiv), _fp(out))
    return out


def interp(feat, idx, d2) -> torch.Tensor:
    """feat (B,M,D), idx/d2 (B,N,k) -> (B,N,D)   [common.py:115-122]"""
    feat, d2 = _f32(feat), _f32(d2)
    idx = idx.detach().to("cpu", torch.int32).contiguous()
    B, N, K = idx.shape
    M, D = feat.shape[1], feat.shape[2]
    out = torch.empty(B, N, D, dtype=torch.float32)
    _load().orc_interp(_fp(feat), _ip(idx), _fp(d2), B, N, M, D, K, _fp(out))
    return out


def edge_feature(x, idx) -> torch.Tensor:
    """x (B,F,N), idx (B,N,k) -> (B,2F,N,k)   [dgcnn.py:41-55]"""
    x = _f32(x)
    idx = idx.detach().to("cpu", torch.int32).contiguous()
    B, F, N = x.shape
    K = idx.shape[2]
    out = torch.empty(B, 2 * F, N, K, dtype=torch.float32)
    _load().orc_edge_feature(_fp(x), _ip(idx), B, F, N, K, _fp(out))
    return out
