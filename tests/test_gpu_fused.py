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
