"""CPU: the C-ABI library loads and exports every symbol include/pcnbr.h declares; the host layer
mirrors the reference interface (names, signatures, state_dict keys, error behaviour) and refuses
to run without CUDA tensors (no CPU fallback)."""
import ctypes
import inspect
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pcnbr.h")).read()
    return sorted(set(re.findall(r"PCNBR_API [^;(]*?\b(pcnbr_\w+)\(", text)))


def test_header_symbols_are_exported_and_bound(pkg):
    syms = _declared_symbols()
    assert len(syms) >= 18
    lib = ctypes.CDLL(pkg._lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/pcnbr.h but not exported"
    assert sorted(pkg._lib.PROTOTYPES) == syms          # the ctypes table covers the header 1:1
    assert pkg._lib.load().pcnbr_abi_version() == 1
    assert b"workspace" in pkg._lib.load().pcnbr_error_string(-3)


def test_workspace_size_helpers(pkg):
    assert pkg._lib.size("pcnbr_fps_ws_bytes", 4, 4096) == 0
    assert pkg._lib.size("pcnbr_fps_ws_bytes", 4, 24000) == 4 * 24000 * 4
    assert pkg._lib.size("pcnbr_csr_ws_bytes", 2, 100, 10) == 4 * (2 * 11 + 200)
    assert pkg._lib.size("pcnbr_knn_expand_ws_bytes", 2, 64, 100, 20) == 800


def test_split_k_plan_of_the_gemms(pkg):
    """pcnbr_gemm3x_splits (host logic, no GPU needed): no empty split for any shape -- the kernels reject a plan with one --,
    1 where the output tiles fill the chip or K is short, a split owns at least 4 K blocks, the big weight gradients are
    bounded by two units per SM, the 26 us weight gradients of the deep PointNet++ levels get every K block they can use;
    the fp16-split dispatch rule takes the tensor-bound layers of DGCNN's head and none of PointNet++'s."""
    import random
    rng = random.Random(0)
    f = lambda M, N, K: pkg._lib.size("pcnbr_gemm3x_splits", M, N, K)
    for _ in range(2000):
        M, N, K = rng.randint(1, 2048), rng.randint(1, 2048), rng.randint(1, 1 << 20)
        s = f(M, N, K)
        kb = (K + 31) // 32
        per = (kb + s - 1) // s
        assert s >= 1 and (kb + per - 1) // per == s, (M, N, K, s)
        assert s == 1 or per >= 4, (M, N, K, s)
    assert f(65536, 1024, 384) == 1 and f(1 << 20, 32, 32) == 1 and f(256, 256, 1024) == 1
    assert f(256, 512, 2048) == 16 and f(128, 256, 8192) == 64
    assert f(1024, 384, 65536) == 9 and f(128, 128, 131072) == 147
    h2 = lambda M, N, K: pkg._lib.size("pcnbr_gemm2h_preferred", M, N, K)
    assert h2(65536, 1024, 384) and h2(65536, 512, 1408) and h2(65536, 256, 512) and h2(1024, 384, 65536)
    assert not h2(131072, 128, 128) and not h2(1 << 20, 64, 32) and not h2(16384, 512, 256)


def test_no_cpu_fallback(pkg):
    x = torch.rand(1, 64, 3)
    with pytest.raises(RuntimeError, match="CUDA"):
        pkg.common.sample(x, 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        pkg.common.group(x[:, :4], x, x, 0.1, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        pkg.dgcnn.knn(torch.rand(1, 3, 64), 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        pkg.common.reduce(torch.rand(1, 2, 3, 4), "max")
    with pytest.raises(ValueError):
        pkg.common.reduce(torch.rand(1, 2, 3, 4), "sum")          # reference common.py:91


def test_product_never_imports_oracle():
    pkg_dir = os.path.join(ROOT, "3d-semantic-segmentation-benchmark_b200")
    for fn in os.listdir(pkg_dir):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg_dir, fn)).read()
            assert "oracle" not in src.replace("# oracle", ""), fn


def test_signatures_mirror_reference(pkg):
    c, d = pkg.common, pkg.dgcnn

    def leading(fn, names):
        """the reference's parameters come first, in order; anything behind them is an optional extension (default None:
        start_idx for seeded tests, lengths for the length-aware evaluation of zero-padded batches)"""
        ps = list(inspect.signature(fn).parameters.values())
        assert [q.name for q in ps[:len(names)]] == names
        assert all(q.default is None for q in ps[len(names):]), [q.name for q in ps[len(names):]]

    leading(c.group, ["centroid_coords", "coords", "features", "r", "K", "normalize"])
    leading(c.interpolate, ["points", "coords_1", "coords_2", "k"])
    leading(d.knn, ["x", "k"])
    assert list(inspect.signature(c.reduce).parameters) == ["x", "type"]
    assert list(inspect.signature(c.sample).parameters)[:2] == ["coords", "C"]
    assert list(inspect.signature(c.SetAbstraction.__init__).parameters)[1:] == [
        "C", "radius", "in_channels", "mlps", "K", "pooling_type", "grouping_norm"]
    assert list(inspect.signature(c.InvResMLP.__init__).parameters)[1:] == [
        "radius", "in_channels", "mlp_size", "K", "pooling_type"]
    assert list(inspect.signature(d.get_graph_feature).parameters) == ["x", "k", "idx", "dim9"]
    assert list(inspect.signature(d.EdgeConv.__init__).parameters)[1:] == ["in_channels", "out_channels", "k"]
    assert isinstance(d.get_loss(), torch.nn.CrossEntropyLoss) and d.get_loss().ignore_index == -1
    assert isinstance(d.get_model(13, use_color=False, k=8), d.DGCNN)
    with pytest.raises(ValueError, match="6-channel"):
        d.DGCNNWithColor(13)(torch.rand(1, 3, 32))                 # reference dgcnn.py:221-222


def test_state_dict_keys_interchange_with_reference(pkg, golden):
    """Checkpoints written by the reference (train.py:88) must load: same keys, same shapes."""
    keys = golden("state_keys")
    mine = {
        "PointNetpp": pkg.PointNetpp(14), "PointNeXt": pkg.PointNeXt(14, "s"),
        "DGCNN": pkg.DGCNN(13), "DGCNNWithColor": pkg.DGCNNWithColor(13),
    }
    for name, model in mine.items():
        got = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
        assert got == [(k, tuple(s)) for k, s in keys[name]], name
