"""GPU: the fused EdgeConv (algebraic split) against the oracle's literal EdgeConv
(models/dgcnn/dgcnn.py:60-77) with the same parameters: outputs, input/parameter gradients and BatchNorm
running statistics within 1e-4 relative; and against the exact (unfused) CUDA path."""
import pytest
import torch

from oracle import ref_ops as O

pytestmark = pytest.mark.gpu


def _close(a, b, tol=1e-4):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    scale = b.abs().max().item() + 1e-30
    err = (a - b).abs().max().item()
    assert err <= tol * scale, f"max abs err {err:.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("F,Oc,N,k,B", [(3, 64, 1024, 20, 2), (64, 64, 1024, 20, 2), (64, 128, 700, 20, 3), (8, 32, 300, 7, 2)])
@pytest.mark.parametrize("train", [True, False])
def test_fused_edgeconv_matches_reference_layer(pkg, dev, F, Oc, N, k, B, train):
    g = torch.Generator().manual_seed(F * 1000 + Oc + N)
    x = torch.randn(B, F, N, generator=g)
    if F == 3:
        x = x * 0.3 + torch.tensor([15.0, 4.0, 1.0]).view(1, 3, 1)
    torch.manual_seed(7)
    ref = O.EdgeConv(F, Oc, k)
    with torch.no_grad():                                   # non-trivial affine incl. NEGATIVE gammas (min-pool branch)
        ref.conv[1].weight.copy_(torch.randn(Oc, generator=g))
        ref.conv[1].bias.copy_(torch.randn(Oc, generator=g))
        ref.conv[1].running_mean.copy_(torch.randn(Oc, generator=g) * 0.1)
        ref.conv[1].running_var.copy_(torch.rand(Oc, generator=g) + 0.5)
    net = pkg.dgcnn.EdgeConv(F, Oc, k)
    net.load_state_dict(ref.state_dict())
    net = net.to(dev)
    assert net.fused
    ref.train(train), net.train(train)
    w = torch.randn(B, Oc, N, generator=g)
    xr = x.clone().requires_grad_(True)
    out_r = ref(xr)
    (out_r * w).sum().backward()
    xg = x.to(dev).requires_grad_(True)
    out_g = net(xg)
    (out_g * w.to(dev)).sum().backward()
    _close(out_g, out_r)
    _close(xg.grad, xr.grad, 2e-4)
    for (n, pg), (_, pr) in zip(net.named_parameters(), ref.named_parameters()):
        _close(pg.grad, pr.grad, 2e-4)
    _close(net.conv[1].running_mean, ref.conv[1].running_mean)
    _close(net.conv[1].running_var, ref.conv[1].running_var)
    assert int(net.conv[1].num_batches_tracked) == int(ref.conv[1].num_batches_tracked)


def test_fused_equals_exact_cuda_path(pkg, dev):
    torch.manual_seed(3)
    a = pkg.dgcnn.EdgeConv(64, 64, 20).to(dev)
    b = pkg.dgcnn.EdgeConv(64, 64, 20).to(dev)
    b.load_state_dict(a.state_dict())
    b.fused = False
    x = torch.randn(2, 64, 2048, device=dev)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya, yb = a(xa), b(xb)
    _close(ya, yb)
    ya.square().sum().backward()
    yb.square().sum().backward()
    _close(xa.grad, xb.grad, 2e-4)
    _close(a.conv[0].weight.grad, b.conv[0].weight.grad, 2e-4)


# --------------------------------------------------------------------------- 1x1 convolutions on the 3xTF32 tcgen05 GEMM

@pytest.mark.parametrize("R,Cin,Cout,bias", [(8192, 1408, 512, False), (4096, 384, 1024, True), (5000, 132, 68, True), (65536, 512, 256, False),
                                             (16384, 12, 32, True), (32768, 32, 64, False), (8192, 64, 32, True), (1024, 260, 256, True)])
def test_linear_rows_tensor_core_matches_fp64(pkg, dev, R, Cin, Cout, bias):
    """ops.linear_rows (the Conv1d/Conv2d kernel-1 layers of common.py:125-178 / dgcnn.py:95-126 on point-major rows):
    output, input gradient, weight gradient (split-K, deterministic) and bias gradient against float64, well inside the
    fp32 parity bar of 1e-4 relative: the 3xTF32 split itself is good to ~1e-6 |a||b|; the tensor core's fp32
    accumulation (truncating, one step per K=8 instruction) brings it to ~1.5e-5 of the largest output."""
    g = torch.Generator().manual_seed(R + Cin)
    x = (torch.randn(R, Cin, generator=g) * 0.7 + 0.3)
    w = torch.randn(Cout, Cin, generator=g) / Cin ** 0.5
    b = torch.randn(Cout, generator=g) if bias else None
    gy = torch.randn(R, Cout, generator=g)
    xd, wd = x.to(dev).requires_grad_(True), w.to(dev).requires_grad_(True)
    bd = b.to(dev).requires_grad_(True) if bias else None
    launches0 = pkg._lib.launches
    y = pkg.ops.linear_rows(xd, wd, bd)
    y.backward(gy.to(dev))
    assert pkg._lib.launches > launches0, "linear_rows did not run on libpcnbr"
    x64, w64 = x.double().requires_grad_(True), w.double().requires_grad_(True)
    b64 = b.double().requires_grad_(True) if bias else None
    y64 = torch.nn.functional.linear(x64, w64, b64)
    y64.backward(gy.double())
    _close(y, y64, 3e-5)
    _close(xd.grad, x64.grad, 3e-5)
    _close(wd.grad, w64.grad, 3e-5)
    if bias:
        _close(bd.grad, b64.grad, 1e-5)
    # deterministic: the split-K weight gradient is summed in a fixed order
    xd2, wd2 = x.to(dev).requires_grad_(True), w.to(dev).requires_grad_(True)
    pkg.ops.linear_rows(xd2, wd2, bd.detach() if bias else None).backward(gy.to(dev))
    assert torch.equal(wd2.grad, wd.grad) and torch.equal(xd2.grad, xd.grad)


@pytest.mark.parametrize("a_mn", [False, True])
@pytest.mark.parametrize("b_mn", [False, True])
@pytest.mark.parametrize("M,N,K", [(300, 72, 100), (128, 256, 64), (1000, 40, 2048), (64, 12, 40000)])
def test_gemm3x_all_operand_layouts(pkg, dev, a_mn, b_mn, M, N, K):
    """pcnbr_gemm3x_f32 with K-major and MN-major (transposed-in-memory) operands, ragged sizes, split-K: all against
    float64.  The operands are plain fp32 matrices; the hi/lo split happens inside the kernel."""
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, generator=g) + 0.25
    Bm = torch.randn(N, K, generator=g) - 0.1
    pad = lambda t: torch.nn.functional.pad(t, (0, (-t.shape[1]) % 4))            # 16-byte row pitch, logical width kept
    Am = pad(A.t().contiguous() if a_mn else A).to(dev)
    Bmm = pad(Bm.t().contiguous() if b_mn else Bm).to(dev)
    out = pkg.ops._gemm3x(Am, a_mn, Bmm, b_mn, M, N, K)
    ref = A.double() @ Bm.double().t()
    _close(out, ref, 3e-5)
    assert torch.equal(out, pkg.ops._gemm3x(Am, a_mn, Bmm, b_mn, M, N, K))


@pytest.mark.parametrize("a_mn", [False, True])
@pytest.mark.parametrize("b_mn", [False, True])
@pytest.mark.parametrize("M,N,K,scale_a,scale_b", [(4096, 512, 1408, 1.0, 1.0), (4000, 500, 1400, 3e-7, 2e3), (2048, 1024, 384, 40.0, 1e-3),
                                                   (512, 1408, 16384, 1e-4, 1.0), (640, 1000, 9000, 1.0, 1.0),
                                                   (4096, 384, 2048, 1.0, 1.0), (1024, 384, 16384, 5e-3, 1.0), (2000, 372, 2100, 1.0, 30.0),
                                                   (640, 576, 4096, 1.0, 1.0)])
def test_gemm_fp16_split_all_operand_layouts(pkg, dev, a_mn, b_mn, M, N, K, scale_a, scale_b):
    """csrc/gemm_h2.cu (two-term fp16 split on kind::f16, per-tensor power-of-two scales from pcnbr_absmax_f32): K-major and
    MN-major operands (transposed by the in-kernel converters), ragged sizes, split-K, operands far outside fp16's range,
    wide dynamic range inside one operand -- all against float64 at the 3xTF32 kernel's bar.  N = 384 / 372 / 576 take the
    192-column tile (MN-major or pre-split B; a raw K-major B keeps 256), full and ragged, with and without split-K."""
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = (torch.randn(M, K, generator=g) + 0.25) * scale_a
    Bm = (torch.randn(N, K, generator=g) - 0.1) * scale_b
    A[::7] *= 1e-4                                      # rows 10^4 below the tensor's maximum: lo falls into fp16 subnormals
    Bm[:, ::5] *= 1e-3
    pad = lambda t: torch.nn.functional.pad(t, (0, (-t.shape[1]) % 4))
    Am = pad(A.t().contiguous() if a_mn else A).to(dev)
    Bmm = pad(Bm.t().contiguous() if b_mn else Bm).to(dev)
    pkg._lib.prof_enable(True)
    pkg._lib.prof_collect()
    out = pkg.ops._gemm3x(Am, a_mn, Bmm, b_mn, M, N, K, force_h2=True)
    ran = pkg._lib.prof_collect()
    pkg._lib.prof_enable(False)
    assert any(k.startswith("gemm2h_kernel") for k in ran) and "absmax_kernel" in ran, sorted(ran)
    ref = A.double() @ Bm.double().t()
    _close(out, ref, 3e-5)
    # element-wise error against sum_k |a||b| (scale-free: rows / columns 10^3-10^4 below the tensor maximum count fully):
    # the split contributes ~3 * 2^-24; what remains is the tensor core's own fp32 accumulation (measured 1.6e-6 at K = 1408,
    # the same hardware path as the 3xTF32 kernel) -- fp32-grade, 60x inside the 1e-4 parity bar
    bound = A.double().abs() @ Bm.double().abs().t()
    rel = ((out.cpu().double() - ref).abs() / (bound + 1e-300)).max().item()
    print(f"gemm2h M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn}: max |err| / sum|a||b| = {rel:.3e}")
    assert rel <= 4e-6, f"max error / sum|a||b| = {rel:.3e}"
    assert torch.equal(out, pkg.ops._gemm3x(Am, a_mn, Bmm, b_mn, M, N, K, force_h2=True))
    # B pre-split once (pcnbr_split_f16, the weight path of the layers): the same products, bit for bit
    amax_b = pkg.ops._absmax(Bmm)
    bs = pkg.ops._presplit(Bmm, b_mn, amax_b)
    assert bs.shape[:2] == (2, N)
    out2 = pkg.ops._gemm3x(Am, a_mn, Bmm, b_mn, M, N, K, amax_b=amax_b, b_split=bs, force_h2=True)
    assert torch.equal(out2, out)


@pytest.mark.parametrize("R,Cin,Cout", [(8192, 384, 1024), (4100, 1024, 512), (16384, 512, 256), (3000, 200, 136), (40000, 64, 128)])
def test_gemm_fp16_split_planes_link_the_three_gemms_of_a_layer(pkg, dev, R, Cin, Cout):
    """pcnbr_gemm2h_ex2_f32: the forward GEMM (A = x) and the input-gradient GEMM (A = gy) also write the [hi | lo] fp16 planes
    of the A tiles they convert -- bit-identical to pcnbr_split_f16 -- and the weight gradient reads BOTH operands from those
    planes (MN-major, no in-kernel conversion): bit-identical to the weight gradient that converts gy and x itself.  Ragged
    row counts (TMA clipping / zero fill), a 192-column tile, split-K."""
    ops = pkg.ops
    g = torch.Generator().manual_seed(R + Cin)
    x = ((torch.randn(R, Cin, generator=g) + 0.3) * 7.0).to(dev)
    w = (torch.randn(Cout, Cin, generator=g) / Cin ** 0.5).to(dev)
    gy = (torch.randn(R, Cout, generator=g) * 1e-4).to(dev)
    gy[::5] *= 1e-3
    ax, aw, ag = ops._absmax(x), ops._absmax(w), ops._absmax(gy)
    xp, gp = ops._new_planes(x, ax), ops._new_planes(gy, ag)
    xp.fill_(float("nan")); gp.fill_(float("nan"))
    y = ops._gemm3x(x, False, w, False, R, Cout, Cin, amax_a=ax, amax_b=aw, b_split=ops._wsplit(w, False, aw), a_planes_out=xp, force_h2=True)
    dx = ops._gemm3x(gy, False, w, True, R, Cin, Cout, amax_a=ag, amax_b=aw, b_split=ops._wsplit(w, True, aw), a_planes_out=gp, force_h2=True)
    # the outputs do not depend on the extra stores
    assert torch.equal(y, ops._gemm3x(x, False, w, False, R, Cout, Cin, amax_a=ax, amax_b=aw, b_split=ops._wsplit(w, False, aw), force_h2=True))
    assert torch.equal(dx, ops._gemm3x(gy, False, w, True, R, Cin, Cout, amax_a=ag, amax_b=aw, b_split=ops._wsplit(w, True, aw), force_h2=True))
    assert torch.equal(xp[:, :, :Cin], ops._presplit(x, False, ax)[:, :, :Cin])
    assert torch.equal(gp[:, :, :Cout], ops._presplit(gy, False, ag)[:, :, :Cout])
    dw_ref = ops._gemm3x(gy, True, x, True, Cout, Cin, R, amax_a=ag, amax_b=ax, force_h2=True)
    pkg._lib.prof_enable(True)
    pkg._lib.prof_collect()
    dw = ops._gemm3x(gy, True, x, True, Cout, Cin, R, amax_a=ag, amax_b=ax, a_mns=gp, b_mns=xp, force_h2=True)
    ran = pkg._lib.prof_collect()
    pkg._lib.prof_enable(False)
    assert any(k.startswith("gemm2h_kernel") for k in ran), sorted(ran)
    _close(dw, gy.double().cpu().t() @ x.double().cpu(), 3e-5)
    assert torch.equal(dw, dw_ref)


def test_gemm_fp16_split_planes_of_a_concatenated_input(pkg, dev):
    """A = [x1 | x2] (never materialised): the forward GEMM writes the planes of both halves -- or of the second one only,
    when the first half's exist already -- scaled with the ONE power of two it uses for the concatenation; a GEMM that does
    not take the fp16-split kernel leaves the buffers unmarked and the weight gradient then converts its operands itself."""
    ops = pkg.ops
    R, K1, K2, C = 4096, 384, 1024, 512
    g = torch.Generator().manual_seed(11)
    x1, x2 = torch.randn(R, K1, generator=g).to(dev), (torch.randn(R, K2, generator=g) * 3).to(dev)
    w = (torch.randn(C, K1 + K2, generator=g) / 40).to(dev)
    a1, a2, aw = ops._absmax(x1), ops._absmax(x2), ops._absmax(w)
    for both in (True, False):
        xp1, xp2 = ops._new_planes(x1, a1), ops._new_planes(x2, a2)
        xp1.fill_(float("nan")); xp2.fill_(float("nan"))
        ops._gemm3x(x1, False, w, False, R, C, K1 + K2, None, A2=x2, K1=K1, amax_a=a1, amax_a2=a2, amax_b=aw, b_split=ops._wsplit(w, False, aw),
                    a_planes_out=xp1 if both else None, a2_planes_out=xp2, force_h2=True)
        a12 = xp2._pcnbr_amax
        assert torch.equal(a12, torch.maximum(a1, a2)) and ops._planes_ok(xp1) == both
        assert torch.equal(xp2[:, :, :K2], ops._presplit(x2, False, a12)[:, :, :K2])
        if both:
            assert torch.equal(xp1[:, :, :K1], ops._presplit(x1, False, a12)[:, :, :K1])
        else:
            assert bool(torch.isnan(xp1.float()).all())
    # below the fp16-split threshold the 3xTF32 kernel runs: the planes stay unmarked and unused
    xs, ws = x1[:512].contiguous(), w[:, :K1].contiguous()
    xp = ops._new_planes(xs, ops._absmax(xs))
    ops._gemm3x(xs, False, ws, False, 512, C, K1, amax_a=ops._absmax(xs), amax_b=ops._absmax(ws), a_planes_out=xp)
    assert not ops._planes_ok(xp)


def test_fp16_split_layers_share_planes_and_match_the_converting_path(pkg, dev):
    """conv5 / conv6 of DGCNNWithColor as the model issues them (linear_bn_act_rows on the skip concatenation, then
    linear_bn_act_cat_rows on [skip | conv5 output]): with the operand planes (the skip tensor's split is written once, by
    conv5's forward GEMM, and serves both layers' weight gradients) every gradient equals that of the path that converts
    both operands inside the weight-gradient kernels (PCNBR_GEMM_NO_PLANES)."""
    ops = pkg.ops
    R, K1, C5, C6 = 32768, 384, 1024, 512          # every GEMM of both layers above the fp16-split threshold
    g = torch.Generator().manual_seed(11)
    skip0 = torch.randn(R, K1, generator=g).to(dev)
    w5, w6 = (torch.randn(C5, K1, generator=g) / 20).to(dev), (torch.randn(C6, K1 + C5, generator=g) / 40).to(dev)
    gout = torch.randn(R, C6, generator=g).to(dev)
    def run(no_planes):
        old = ops._GEMM_NO_PLANES
        ops._GEMM_NO_PLANES = no_planes
        try:
            torch.manual_seed(3)
            bn5, bn6 = torch.nn.BatchNorm1d(C5).to(dev), torch.nn.BatchNorm1d(C6).to(dev)
            skip = skip0.clone().requires_grad_(True)
            a5, a6 = w5.clone().requires_grad_(True), w6.clone().requires_grad_(True)
            skip_r = skip * 1.0                       # a non-leaf like the model's concatenation
            r5 = ops.linear_bn_act_rows(skip_r, a5, None, bn5, 0.2)
            shared = getattr(skip_r, "_pcnbr_planes", None) is not None
            r6 = ops.linear_bn_act_cat_rows(skip_r, r5, a6, None, bn6, 0.2)
            r6.backward(gout)
            return [skip.grad, a5.grad, a6.grad, bn5.weight.grad, bn6.weight.grad, r6.detach()], shared
        finally:
            ops._GEMM_NO_PLANES = old
    got, shared = run(False)
    ref, shared_ref = run(True)
    assert shared and not shared_ref
    for a, b in zip(got, ref):
        # conv6 scales [skip | r5] with ONE power of two in its forward GEMM, so r5's planes carry that scale while the
        # converting weight-gradient kernel scales r5 by its own maximum: the same fp16 split except where `lo` falls into
        # fp16 subnormals -- differences at 2^-40 of the tensor maximum, a last-bit flip here and there
        _close(a, b, 2e-6)
    assert torch.equal(got[1], ref[1]) and torch.equal(got[5], ref[5])        # conv5 (own scale) and the forward: bit for bit


def test_gemm_fp16_split_concatenated_input_and_degenerate_operands(pkg, dev):
    """A = [A1 | A2] read from two matrices with different magnitudes (one shared scale); an all-zero operand; a bias."""
    g = torch.Generator().manual_seed(5)
    M, K1, K2, N = 4096, 384, 1024, 512
    A1, A2 = torch.randn(M, K1, generator=g) * 30.0, torch.randn(M, K2, generator=g) * 0.02
    W = torch.randn(N, K1 + K2, generator=g) / 40.0
    b = torch.randn(N, generator=g)
    out = pkg.ops._gemm3x(A1.to(dev), False, W.to(dev), False, M, N, K1 + K2, b.to(dev), A2=A2.to(dev), K1=K1, force_h2=True)
    ref = torch.cat((A1, A2), 1).double() @ W.double().t() + b.double()
    _close(out, ref, 3e-5)
    z = pkg.ops._gemm3x(torch.zeros(M, K1 + K2, device=dev), False, W.to(dev), False, M, N, K1 + K2, force_h2=True)
    assert float(z.abs().max()) == 0.0


@pytest.mark.parametrize("R,Cin,Cout", [(4096, 9, 32), (512, 768, 256), (32, 512, 256), (65536, 256, 13), (100, 7, 5)])
def test_linear_rows_odd_widths_and_few_rows_run_on_libpcnbr(pkg, dev, R, Cin, Cout):
    """Channel counts that are not multiples of 4 (9-channel stem, 13-class head) are zero-padded to the TMA pitch and
    tiny row counts (the deepest PointNet++ levels at small batches) take the same tensor-core kernel: no shape of the
    models falls back to the library SGEMM (ops.fallbacks() stays empty)."""
    g = torch.Generator().manual_seed(R + Cin)
    x, w, b = torch.randn(R, Cin, generator=g), torch.randn(Cout, Cin, generator=g), torch.randn(Cout, generator=g)
    gy = torch.randn(R, Cout, generator=g)
    pkg.ops.reset_fallbacks()
    xd, wd, bd = (t.to(dev).requires_grad_(True) for t in (x, w, b))
    launches0 = pkg._lib.launches
    y = pkg.ops.linear_rows(xd, wd, bd)
    y.backward(gy.to(dev))
    assert pkg._lib.launches > launches0 and pkg.ops.fallbacks() == {}
    x64, w64, b64 = (t.double().requires_grad_(True) for t in (x, w, b))
    y64 = torch.nn.functional.linear(x64, w64, b64)
    y64.backward(gy.double())
    _close(y, y64, 3e-5)
    _close(xd.grad, x64.grad, 3e-5)
    _close(wd.grad, w64.grad, 3e-5)
    _close(bd.grad, b64.grad, 1e-5)


# --------------------------------------------------------------------------- fused BatchNorm + (Leaky)ReLU over rows

@pytest.mark.parametrize("R,C,slope", [(65536, 64, 0.0), (4099, 32, 0.0), (8192, 1024, 0.2), (1000, 2048, 0.2), (300, 8, 0.0),
                                       (16384, 256, 0.2)])
@pytest.mark.parametrize("train", [True, False])
def test_batchnorm_act_rows_matches_fp64(pkg, dev, R, C, slope, train):
    """ops.batchnorm_act_rows = nn.BatchNorm1d -> ReLU / LeakyReLU of common.py:146,175 and dgcnn.py:67-70 on point-major
    rows: output, input / gamma / beta gradients, running statistics and num_batches_tracked against the float64 module,
    with a large per-channel offset (S3DIS coordinates are tens of metres from the origin)."""
    g = torch.Generator().manual_seed(R + C)
    x = torch.randn(R, C, generator=g) * (0.2 + torch.rand(C, generator=g)) + 20.0 * torch.randn(C, generator=g)
    gy = torch.randn(R, C, generator=g)
    bn = torch.nn.BatchNorm1d(C)
    with torch.no_grad():
        bn.weight.copy_(torch.randn(C, generator=g))
        bn.bias.copy_(torch.randn(C, generator=g))
        bn.running_mean.copy_(x.mean(0) + 0.1 * torch.randn(C, generator=g))
        bn.running_var.copy_(x.var(0) * (0.5 + torch.rand(C, generator=g)))
    import copy
    bn64 = copy.deepcopy(bn).double()
    bnd = copy.deepcopy(bn).to(dev)
    bn64.train(train); bnd.train(train)
    # the activation's derivative jumps at 0: keep the upstream gradient off the pre-activations that fp32 and fp64
    # could put on different sides of it
    with torch.no_grad():
        pre = copy.deepcopy(bn64)(x.double())
        gy = gy * (pre.abs() > 1e-3).float()
    xd = x.to(dev).requires_grad_(True)
    launches0 = pkg._lib.launches
    y = pkg.ops.batchnorm_act_rows(xd, bnd, slope)
    y.backward(gy.to(dev))
    assert pkg._lib.launches >= launches0 + 5, "batchnorm_act_rows did not run on libpcnbr"
    x64 = x.double().requires_grad_(True)
    y64 = torch.nn.functional.leaky_relu(bn64(x64), slope)
    y64.backward(gy.double())
    _close(y, y64, 2e-5)
    _close(xd.grad, x64.grad, 1e-4)
    _close(bnd.weight.grad, bn64.weight.grad, 1e-4)
    _close(bnd.bias.grad, bn64.bias.grad, 1e-4)
    _close(bnd.running_mean, bn64.running_mean, 1e-6)
    _close(bnd.running_var, bn64.running_var, 1e-5)
    assert int(bnd.num_batches_tracked) == int(bn64.num_batches_tracked)
    # deterministic
    xd2 = x.to(dev).requires_grad_(True)
    bnd.zero_grad()
    pkg.ops.batchnorm_act_rows(xd2, bnd, slope).backward(gy.to(dev))
    assert torch.equal(xd2.grad, xd.grad)


@pytest.mark.parametrize("D", [30.0, 100.0])
def test_batchnorm_statistics_survive_an_outlier_first_row(pkg, dev, D):
    """the statistics are shifted sums about row 0.  With row 0 D sigma away from everything else, sums
    ACCUMULATED about it lose ~rows-per-block * 2^-24 * D^2 of the variance (60 % at D = 100); every block now accumulates
    about its own first row and moves its sums to the common pivot once -- what remains is a few 2^-24 * D^2."""
    R, C = 200000, 64
    g = torch.Generator().manual_seed(int(D))
    x = 5.0 + 0.01 * torch.randn(R, C, generator=g)                  # near-constant channels: mean 5, sigma 0.01
    x[0] += D * 0.01                                                  # the outlier pivot
    x[R // 2, :8] -= D * 0.01                                         # and an outlier that is some block's first row or not
    bn = torch.nn.BatchNorm1d(C).to(dev)
    y = pkg.ops.batchnorm_act_rows(x.to(dev), bn, 1.0).cpu().double()          # slope 1: the normalised values themselves
    x64 = x.double()
    ref = (x64 - x64.mean(0)) / torch.sqrt(x64.var(0, unbiased=False) + bn.eps)
    err = (y - ref).abs().max().item() / ref.abs().max().item()
    assert err <= 6e-7 * D * D + 1e-5, f"relative error {err:.2e}"
    rv = bn.running_var.cpu().double()
    want = 0.9 + 0.1 * x64.var(0, unbiased=True)
    assert ((rv - want).abs() / want).max().item() <= 1e-5


def test_batchnorm_act_rows_unsupported_width_uses_library(pkg, dev):
    x = torch.randn(512, 13, device=dev)
    bn = torch.nn.BatchNorm1d(13).to(dev)
    y = pkg.ops.batchnorm_act_rows(x, bn, 0.2)
    ref = torch.nn.functional.leaky_relu(torch.nn.functional.batch_norm(x.double(), None, None, bn.weight.double(), bn.bias.double(), True), 0.2)
    _close(y, ref, 1e-5)


@pytest.mark.parametrize("R,Cin,Cout,slope,bias", [(16384, 12, 32, 0.0, True), (8192, 384, 1024, 0.2, False), (2048, 768, 256, 0.0, True)])
@pytest.mark.parametrize("train", [True, False])
def test_linear_bn_act_rows_matches_fp64(pkg, dev, R, Cin, Cout, slope, bias, train):
    """ops.linear_bn_act_rows = one Conv(kernel 1) -> BatchNorm -> (Leaky)ReLU block (common.py:141-147,170-176,
    dgcnn.py:95-126) as a single autograd node, against the float64 modules: output, input / weight / BatchNorm
    gradients.  The conv bias gradient is exactly 0 in training mode (BatchNorm removes per-channel shifts; the fp64
    value is ~1e-12 of the weight gradients) and gamma*rstd*sum(g') in eval mode."""
    import copy
    g = torch.Generator().manual_seed(R + Cin + Cout)
    x = torch.randn(R, Cin, generator=g) * 0.5 + 0.2
    lin = torch.nn.Linear(Cin, Cout, bias=bias)
    bn = torch.nn.BatchNorm1d(Cout)
    with torch.no_grad():
        bn.weight.copy_(torch.randn(Cout, generator=g)); bn.bias.copy_(torch.randn(Cout, generator=g) * 0.5)
        bn.running_mean.copy_(torch.randn(Cout, generator=g) * 0.1); bn.running_var.copy_(torch.rand(Cout, generator=g) * 0.2 + 0.05)
    lin64, bn64 = copy.deepcopy(lin).double(), copy.deepcopy(bn).double()
    lind, bnd = copy.deepcopy(lin).to(dev), copy.deepcopy(bn).to(dev)
    bn64.train(train); bnd.train(train)
    with torch.no_grad():
        pre = copy.deepcopy(bn64)(lin64(x.double()))
    gy = torch.randn(R, Cout, generator=g) * (pre.abs() > 1e-3).float()
    xd = x.to(dev).requires_grad_(True)
    y = pkg.ops.linear_bn_act_rows(xd, lind.weight, lind.bias, bnd, slope)
    y.backward(gy.to(dev))
    x64 = x.double().requires_grad_(True)
    y64 = torch.nn.functional.leaky_relu(bn64(lin64(x64)), slope)
    y64.backward(gy.double())
    _close(y, y64, 3e-5)
    _close(xd.grad, x64.grad, 1e-4)
    _close(lind.weight.grad, lin64.weight.grad, 1e-4)
    _close(bnd.weight.grad, bn64.weight.grad, 1e-4)
    _close(bnd.bias.grad, bn64.bias.grad, 1e-4)
    _close(bnd.running_mean, bn64.running_mean, 1e-5)
    _close(bnd.running_var, bn64.running_var, 1e-4)
    if bias:
        if train:
            assert float(lind.bias.grad.abs().max()) == 0.0
            assert float(lin64.bias.grad.abs().max()) < 1e-9 * float(lin64.weight.grad.abs().max())
        else:
            _close(lind.bias.grad, lin64.bias.grad, 1e-4)


@pytest.mark.parametrize("G,K,Cin,Cout,slope", [(2048, 32, 32, 64, 0.0), (512, 32, 132, 128, 0.0), (300, 16, 64, 32, 0.2)])
@pytest.mark.parametrize("train", [True, False])
def test_linear_bn_act_maxpool_rows_matches_fp64(pkg, dev, G, K, Cin, Cout, slope, train):
    """ops.linear_bn_act_maxpool_rows = last set-abstraction MLP layer + reduce(.., 'max') (common.py:141-147, 85-86,
    211-214) as one node that pools BEFORE BatchNorm/activation (monotone per channel; negative gammas take the min):
    pooled output, input / weight / BatchNorm gradients and running statistics against the float64 modules."""
    import copy
    g = torch.Generator().manual_seed(G + K + Cin)
    x = torch.randn(1, G, K, Cin, generator=g) * 0.5 + 0.2
    lin = torch.nn.Linear(Cin, Cout)
    bn = torch.nn.BatchNorm1d(Cout)
    with torch.no_grad():
        bn.weight.copy_(torch.randn(Cout, generator=g)); bn.bias.copy_(torch.randn(Cout, generator=g) * 0.5)
        bn.running_mean.copy_(torch.randn(Cout, generator=g) * 0.1); bn.running_var.copy_(torch.rand(Cout, generator=g) * 0.2 + 0.05)
    lin64, bn64 = copy.deepcopy(lin).double(), copy.deepcopy(bn).double()
    lind, bnd = copy.deepcopy(lin).to(dev), copy.deepcopy(bn).to(dev)
    bn64.train(train); bnd.train(train)
    with torch.no_grad():                                   # keep the upstream gradient off near-ties and the kink at 0
        act = torch.nn.functional.leaky_relu(copy.deepcopy(bn64)(lin64(x.double()).view(-1, Cout)), slope).view(G, K, Cout)
        top2 = act.topk(2, dim=1).values
        safe = ((top2[:, 0] - top2[:, 1]) > 1e-4) & (top2[:, 0].abs() > 1e-3)
    gy = torch.randn(G, Cout, generator=g) * safe.float()
    xd = x.to(dev).requires_grad_(True)
    y = pkg.ops.linear_bn_act_maxpool_rows(xd, lind.weight, lind.bias, bnd, slope)
    assert y.shape == (1, G, Cout)
    y.backward(gy.to(dev).view(1, G, Cout))
    x64 = x.double().requires_grad_(True)
    y64 = torch.nn.functional.leaky_relu(bn64(lin64(x64).view(-1, Cout)), slope).view(G, K, Cout).max(dim=1).values
    y64.backward(gy.double())
    _close(y, y64.view(1, G, Cout), 3e-5)
    _close(xd.grad, x64.grad, 1e-4)
    _close(lind.weight.grad, lin64.weight.grad, 1e-4)
    _close(bnd.weight.grad, bn64.weight.grad, 1e-4)
    _close(bnd.bias.grad, bn64.bias.grad, 1e-4)
    _close(bnd.running_mean, bn64.running_mean, 1e-5)
    _close(bnd.running_var, bn64.running_var, 1e-4)


@pytest.mark.parametrize("R,K1,K2,Cout", [(8192, 384, 1024, 512), (4096, 64, 36, 128)])
def test_linear_bn_act_cat_rows_equals_materialised_concatenation(pkg, dev, R, K1, K2, Cout):
    """ops.linear_bn_act_cat_rows (dgcnn.py:147: cat((x1..x4, x5)) -> conv6 -> bn6 -> LeakyReLU) never builds the
    concatenated matrix; it must give what the same block gives on torch.cat of the two inputs."""
    import copy
    g = torch.Generator().manual_seed(R + K1)
    x1 = (torch.randn(R, K1, generator=g) * 0.5).to(dev)
    x2 = (torch.randn(R, K2, generator=g) * 0.5 + 0.1).to(dev)
    w = (torch.randn(Cout, K1 + K2, generator=g) / (K1 + K2) ** 0.5).to(dev)
    gy = torch.randn(R, Cout, generator=g).to(dev)
    bn_a = torch.nn.BatchNorm1d(Cout).to(dev)
    bn_b = copy.deepcopy(bn_a)
    a1, a2, wa = x1.clone().requires_grad_(True), x2.clone().requires_grad_(True), w.clone().requires_grad_(True)
    b1, b2, wb = x1.clone().requires_grad_(True), x2.clone().requires_grad_(True), w.clone().requires_grad_(True)
    ya = pkg.ops.linear_bn_act_cat_rows(a1, a2, wa, None, bn_a, 0.2)
    yb = pkg.ops.linear_bn_act_rows(torch.cat((b1, b2), dim=1), wb, None, bn_b, 0.2)
    ya.backward(gy); yb.backward(gy)
    _close(ya, yb, 1e-6)
    _close(a1.grad, b1.grad, 1e-5)
    _close(a2.grad, b2.grad, 1e-5)
    _close(wa.grad, wb.grad, 1e-5)
    _close(bn_a.weight.grad, bn_b.weight.grad, 1e-5)
    _close(bn_a.running_var, bn_b.running_var, 1e-6)


def test_fused_dropout_mask_and_gradient(pkg, dev):
    """ops.linear_bn_act_rows(..., dropout_p): the nn.Dropout behind conv6 / conv7 (dgcnn.py:117,122) folded into the
    BatchNorm+LeakyReLU kernels.  The mask is not stored: the backward recomputes it from the seed -- so the output must be
    the undropped output times mask/(1-p), the keep rate ~1-p, and all gradients those of multiplying by that same mask."""
    import copy
    R, Cin, Cout, p = 16384, 64, 128, 0.5
    g = torch.Generator().manual_seed(99)
    x = torch.randn(R, Cin, generator=g).to(dev)
    w = (torch.randn(Cout, Cin, generator=g) / 8.0).to(dev)
    gy = torch.randn(R, Cout, generator=g).to(dev)
    bn0 = torch.nn.BatchNorm1d(Cout).to(dev)
    with torch.no_grad():
        bn0.weight.copy_(torch.rand(Cout, generator=g).to(dev) + 0.5)
    bn_a, bn_b = copy.deepcopy(bn0), copy.deepcopy(bn0)
    xa, wa = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    torch.manual_seed(1234)
    ya = pkg.ops.linear_bn_act_rows(xa, wa, None, bn_a, 0.2, dropout_p=p)
    ya.backward(gy)
    xb, wb = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    yb = pkg.ops.linear_bn_act_rows(xb, wb, None, bn_b, 0.2)                 # no dropout: LeakyReLU output is never exactly 0
    mask = (ya != 0).float()
    keep = mask.mean().item()
    assert abs(keep - (1 - p)) < 0.01, keep
    _close(ya, yb.detach() * mask / (1 - p), 1e-6)
    (yb * mask / (1 - p)).backward(gy)
    _close(xa.grad, xb.grad, 1e-5)
    _close(wa.grad, wb.grad, 1e-5)
    _close(bn_a.weight.grad, bn_b.weight.grad, 1e-5)
    _close(bn_a.bias.grad, bn_b.bias.grad, 1e-5)
    # a different draw gives a different mask; the same generator state gives the same one
    torch.manual_seed(1234)
    y2 = pkg.ops.linear_bn_act_rows(x, w, None, copy.deepcopy(bn0), 0.2, dropout_p=p)
    y3 = pkg.ops.linear_bn_act_rows(x, w, None, copy.deepcopy(bn0), 0.2, dropout_p=p)
    assert torch.equal(y2 != 0, mask.bool()) and not torch.equal(y3 != 0, mask.bool())
    # rows and columns are both mixed: per-channel and per-row keep rates are all near 1-p
    assert (mask.mean(0) - (1 - p)).abs().max().item() < 0.03 and (mask.mean(1) - (1 - p)).abs().max().item() < 0.25
