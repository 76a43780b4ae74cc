"""GPU: the drop-in, exercised against the UNMODIFIED host code of the reference (oracle/_ref = verbatim copy made by
oracle/make_ref.py; it travels to the GPU box like a built .so).

sys.path = [dropin/, repo root, oracle/_ref]: the reference's own `models/PointNetpp/PointNetpp.py`,
`models/PointNeXt/PointNeXt.py`, `Training/training.py`, `Training/train_model.py`, `Training/metrics.py` are imported
as they are; only `models/utils/common.py`, `models/dgcnn/dgcnn.py` and `data_processing/block_datasets.py` resolve to the
shims (dropin/README.md), i.e. to libpcnbr.  What /root/reference/train.py:53-88 does is then replayed: build the model,
the block dataloaders over on-disk block files, Adam, `train_epoch`, `evaluate`, save / load a `state_dict`."""
import os
import subprocess
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
DROPIN = os.path.join(ROOT, "dropin")

pytestmark = pytest.mark.gpu

_NS = ("models", "Training", "data_processing")


@pytest.fixture()
def reference_tree():
    if not os.path.isdir(os.path.join(REF, "models", "utils")):
        pytest.fail("oracle/_ref is missing: run `python __graft_entry__.py build` where /root/reference exists")
    saved_path = list(sys.path)
    saved_mods = {k: v for k, v in sys.modules.items() if k.split(".")[0] in _NS}
    for k in saved_mods:
        del sys.modules[k]
    stubs = []
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):      # imported at the top of Training/training.py, used only
        if name not in sys.modules:                                    # by plot_confusion_matrix (never called, training.py:176)
            try:
                __import__(name)
            except ImportError:
                sys.modules[name] = types.ModuleType(name)
                stubs.append(name)
    if "matplotlib" in stubs:
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
        sys.modules["matplotlib.pyplot"].Figure = object
    sys.path[:0] = [DROPIN, ROOT, REF]
    yield REF
    sys.path[:] = saved_path
    for k in [k for k in sys.modules if k.split(".")[0] in _NS]:
        del sys.modules[k]
    sys.modules.update(saved_mods)
    for name in stubs:
        sys.modules.pop(name, None)


class _Logger:
    def __init__(self):
        self.scalars = []

    def add_scalar(self, tag, value, step):
        self.scalars.append((tag, float(value), int(step)))


def _write_blocks(root, areas, n_blocks, seed):
    """Synthetic block files in the on-disk format of data_processing/preprocess_dataset.py:134:
    area_<a>/room<rr>_block<bbb>.pt = (points (n,9) f32, labels (n,14) u8)."""
    g = torch.Generator().manual_seed(seed)
    for a in areas:
        os.makedirs(os.path.join(root, f"area_{a}"), exist_ok=True)
        for b in range(n_blocks):
            n = int(torch.randint(300, 6000, (1,), generator=g))
            ox, oy = torch.randint(0, 20, (2,), generator=g).tolist()
            xyz = torch.rand(n, 3, generator=g) * torch.tensor([1.0, 1.0, 3.0]) + torch.tensor([float(ox), float(oy), 0.0])
            rgb = torch.randint(0, 256, (n, 3), generator=g).float()
            ctr = torch.tensor([ox + 0.5, oy + 0.5, float((xyz[:, 2].min() + xyz[:, 2].max()) / 2)])
            pts = torch.cat((xyz, rgb, xyz - ctr), dim=1)
            lab = torch.nn.functional.one_hot(torch.randint(0, 14, (n,), generator=g), 14).to(torch.uint8)
            torch.save((pts, lab), os.path.join(root, f"area_{a}", f"room{b // 3:02d}_block{b % 3:03d}.pt"))


@pytest.mark.parametrize("which", ["PointNet++", "PointNeXt"])
def test_unmodified_reference_train_loop_runs_on_the_shims(reference_tree, pkg, dev, tmp_path, which):
    # ---- imports exactly as /root/reference/train.py:4-10
    from data_processing.block_datasets import create_block_dataloaders
    from models.PointNeXt.PointNeXt import PointNeXt
    from models.PointNetpp.PointNetpp import PointNetpp
    from Training.train_model import masked_onehot_cross_entropy
    from Training.training import evaluate, train_epoch
    import models.utils.common as shim_common
    import Training.training as ref_training

    assert os.path.realpath(sys.modules[PointNetpp.__module__].__file__).startswith(os.path.realpath(REF))
    assert os.path.realpath(sys.modules[PointNeXt.__module__].__file__).startswith(os.path.realpath(REF))
    assert os.path.realpath(ref_training.__file__).startswith(os.path.realpath(REF))
    assert os.path.realpath(shim_common.__file__).startswith(os.path.realpath(DROPIN))
    assert shim_common.SetAbstraction is pkg.common.SetAbstraction

    _write_blocks(str(tmp_path), areas=(1, 2, 3, 4, 5, 6), n_blocks=3, seed=1)     # 15 train blocks -> 2 batches of 8 / 7
    train_loader, test_loader = create_block_dataloaders(
        data_dir=str(tmp_path), test_areas={6}, train_batch_size=8, test_batch_size=2, num_workers=2,
        train_sampling=4096, test_sampling=None, train_shuffle=True, test_shuffle=False)          # train.py:64-74
    torch.manual_seed(0)
    model = (PointNetpp(part_classes=14) if which == "PointNet++" else PointNeXt(part_classes=14)).to("cuda")   # train.py:55-58
    optimizer = torch.optim.Adam(model.parameters(), lr=0.001)                                     # train.py:79
    logger = _Logger()
    pkg.ops.reset_fallbacks()
    launches0 = pkg._lib.launches
    before = {k: v.detach().clone() for k, v in model.state_dict().items()}
    loss, steps = train_epoch(model, train_loader, masked_onehot_cross_entropy, optimizer, "cuda", logger, 1, 0)   # training.py:29-78
    assert steps == len(train_loader) == 2 and loss == loss and 0.0 < loss < 20.0
    assert pkg._lib.launches - launches0 > 100, "the train steps did not run on libpcnbr"
    assert pkg.ops.fallbacks() == {}, f"library fallbacks on the drop-in path: {pkg.ops.fallbacks()}"
    assert {t for t, _, _ in logger.scalars} == {"Train/Loss", "Train/Accuracy", "Train/Mean_IoU"}
    after = model.state_dict()
    assert any(not torch.equal(before[k], after[k]) for k in before if before[k].dtype.is_floating_point), "Adam did not move the weights"
    # ---- evaluation over zero-padded variable-N batches (training.py:80-133)
    val_loss, acc, miou, ious, matrix = evaluate(model, test_loader, masked_onehot_cross_entropy, "cuda")
    assert 0.0 <= acc <= 1.0 and 0.0 <= miou <= 1.0 and matrix.shape == (14, 14) and int(matrix.sum()) > 0
    assert pkg.ops.fallbacks() == {}, f"library fallbacks in evaluation: {pkg.ops.fallbacks()}"
    # ---- the package's own validation loop (one device-side confusion matrix, no per-class host syncs) returns the same tuple
    torch.manual_seed(11)
    r_loss, r_acc, r_miou, r_ious, r_matrix = evaluate(model, test_loader, masked_onehot_cross_entropy, "cuda")
    torch.manual_seed(11)                                     # same FPS start draws (common.py:22)
    o_loss, o_acc, o_miou, o_ious, o_matrix = pkg.train.evaluate(model, test_loader, masked_onehot_cross_entropy, "cuda")
    assert torch.equal(o_matrix, r_matrix) and abs(o_acc - r_acc) < 1e-6
    assert abs(o_miou - r_miou) < 1e-6 and torch.allclose(o_ious, r_ious, atol=1e-6) and abs(o_loss - float(r_loss)) < 1e-4
    # ---- train.py:88 saves model.state_dict(); the keys are the reference's (checkpoints interchange)
    path = tmp_path / "model.pt"
    torch.save(model.state_dict(), path)
    golden_keys = torch.load(os.path.join(ROOT, "tests", "golden", "state_keys.pt"), weights_only=True)
    want = golden_keys["PointNetpp" if which == "PointNet++" else "PointNeXt"]
    got = [(k, tuple(v.shape)) for k, v in torch.load(path, weights_only=True).items()]
    assert got == [(k, tuple(s)) for k, s in want]


_PURE_REFERENCE = r'''
import sys, torch
sys.path.insert(0, sys.argv[1])                      # oracle/_ref ONLY: the pure reference, CPU
from models.PointNetpp.PointNetpp import PointNetpp
_randint = torch.randint
torch.randint = lambda *a, **k: torch.zeros(a[2] if len(a) > 2 else k["size"], dtype=k.get("dtype", torch.int64))   # FPS start 0 (common.py:22)
torch.manual_seed(123)
net = PointNetpp(13).eval()
g = torch.Generator().manual_seed(5)
with torch.no_grad():
    for m in net.modules():                          # non-trivial running statistics, as a trained checkpoint has
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
xyz = torch.rand(2, 1024, 3, generator=g) * 0.05 + torch.tensor([0.5, 0.25, 0.0])
rgb = torch.rand(2, 1024, 3, generator=g) * 255
x = torch.cat([xyz, rgb, xyz - xyz.mean(dim=1, keepdim=True)], dim=-1)
with torch.no_grad():
    logits = net(x)
torch.save({"state": net.state_dict(), "x": x, "logits": logits}, sys.argv[2])
'''


def test_reference_written_checkpoint_loads_and_reproduces_reference_logits(reference_tree, pkg, dev, tmp_path):
    """A checkpoint written by the PURE reference (its own common.py, CPU, in a subprocess) is loaded into the reference's
    PointNetpp class running on the shims; eval-mode logits on a cloud without selection ties agree to 1e-4."""
    out = tmp_path / "ref_ckpt.pt"
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", PYTHONPATH="")
    subprocess.run([sys.executable, "-c", _PURE_REFERENCE, REF, str(out)], check=True, env=env, timeout=600)
    blob = torch.load(out, weights_only=True)
    from models.PointNetpp.PointNetpp import PointNetpp
    net = PointNetpp(13)
    missing = net.load_state_dict(blob["state"], strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    net = net.to(dev).eval()
    for sa in (net.sa1, net.sa2, net.sa3, net.sa4):
        sa.fps_start = torch.zeros(2, dtype=torch.int32, device=dev)
    pkg.ops.reset_fallbacks()
    with torch.no_grad():
        logits = net(blob["x"].to(dev)).cpu()
    assert pkg.ops.fallbacks() == {}
    ref = blob["logits"]
    err = (logits - ref).abs().max().item()
    assert err <= 1e-4 * ref.abs().max().item() + 1e-5, f"max abs err {err:.3e} vs scale {ref.abs().max().item():.3e}"


def test_length_aware_evaluate_equals_cloud_by_cloud(pkg, dev, tmp_path):
    """pkg.train.evaluate(length_aware=True) over the zero-padded variable-N batches of the evaluation loader
    (data_processing/block_datasets.py:19-25) == every block evaluated alone, unpadded."""
    _write_blocks(str(tmp_path), areas=(1, 2, 3, 4, 5, 6), n_blocks=5, seed=3)
    _, test_loader = pkg.block_datasets.create_block_dataloaders(
        data_dir=str(tmp_path), test_areas={6}, train_batch_size=2, test_batch_size=3, num_workers=0,
        train_sampling=4096, test_sampling=None, train_shuffle=False, test_shuffle=False)
    torch.manual_seed(2)
    model = pkg.PointNetpp(14).to(dev).eval()
    batches = [(p.clone(), l.clone(), n.clone()) for p, l, n in test_loader]
    assert len({int(n) for _, _, ns in batches for n in ns}) > 1, "the fixture needs clouds of different lengths"
    loss, acc, miou, ious, matrix = pkg.train.evaluate(_PerBatchStart(model, dev), batches, None, "cuda", length_aware=True)
    want = torch.zeros(14, 14, dtype=torch.int64)
    with torch.no_grad():
        for pts, lab, lens in batches:
            for b, n in enumerate(lens.tolist()):
                for sa in (model.sa1, model.sa2, model.sa3, model.sa4):
                    sa.fps_start = torch.zeros(1, dtype=torch.int32, device=dev)
                pred = model(pts[b:b + 1, :n].contiguous().to(dev))[0].argmax(-1).cpu()
                want.index_put_((lab[b, :n].argmax(-1).long().cpu(), pred), torch.ones(n, dtype=torch.int64), accumulate=True)
    assert int(matrix.sum()) == int(want.sum())
    assert int((matrix - want).abs().sum()) <= 2 * max(1, int(want.sum()) // 2000), "argmax may flip on an exact near-tie only"
    assert 0.0 <= acc <= 1.0 and loss == loss


class _PerBatchStart(torch.nn.Module):
    """sets the (B,) FPS start of every level to zeros for whatever batch size arrives"""

    def __init__(self, model, dev):
        super().__init__()
        self.model, self.dev = model, dev

    def forward(self, points, lengths=None):
        for sa in (self.model.sa1, self.model.sa2, self.model.sa3, self.model.sa4):
            sa.fps_start = torch.zeros(points.shape[0], dtype=torch.int32, device=self.dev)
        return self.model(points, lengths=lengths)
