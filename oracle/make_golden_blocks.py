"""Generate tests/golden/blocks.pt and tests/golden/scene_windows.pt by running the UNMODIFIED reference
(data_processing/block_datasets.py, models/dgcnn/utils.py imported from /root/reference).
TEST INFRASTRUCTURE ONLY; run in the build container:

    python oracle/make_golden_blocks.py
"""
from __future__ import annotations

import os
import sys
import tempfile

import torch

REF = os.environ.get("PCNBR_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF)
sys.path.insert(1, os.path.join(REF, "models", "dgcnn"))
sys.path.insert(2, ROOT)

from data_processing import block_datasets as RB         # noqa: E402  (reference)
import utils as RU                                        # noqa: E402  (reference models/dgcnn/utils.py)
from oracle import ref_ops as O                           # noqa: E402

SEED = 20261018
SAMPLING = 64


def write_blocks(data_dir, g):
    """Synthetic block files in the reference's layout (area_<a>/room<rr>_block<bbb>.pt = (points (n,9) f32, labels (n,14)
    u8)); some blocks below, some above the sampling size."""
    layout = {1: [(1, 0), (1, 1), (2, 0)], 2: [(1, 3)], 3: [(4, 0), (4, 10)], 4: [(2, 2)], 5: [(7, 1), (7, 2)],
              6: [(1, 0), (1, 1), (3, 5)]}
    sizes = iter([150, 40, 64, 65, 200, 17, 90, 33, 128, 77, 20, 141])
    for area, blocks in layout.items():
        os.makedirs(os.path.join(data_dir, f"area_{area}"))
        for room, block in blocks:
            n = next(sizes)
            pts = torch.rand(n, 9, generator=g) * 3.0
            lab = torch.nn.functional.one_hot(torch.randint(0, 14, (n,), generator=g), 14).to(torch.uint8)
            torch.save((pts, lab), os.path.join(data_dir, f"area_{area}", f"room{room:02d}_block{block:03d}.pt"))


def records(ds):
    return [torch.load(os.path.join(ds.data_dir, f"area_{a}", f"room{r:02d}_block{b:03d}.pt")) for a, r, b in ds.blocks.tolist()]


class TinyModel(torch.nn.Module):
    """Deterministic stand-in for a trained DGCNN: per-point linear scores that also depend on the window's mean, so a
    wrong window cut changes the result.  (B,F,n) -> (logits (B,n,C), None, None)."""
    num_classes = 13

    def __init__(self, g):
        super().__init__()
        self.w = torch.nn.Parameter(torch.randn(13, 6, generator=g))

    def forward(self, x):
        x = x - x.mean(dim=2, keepdim=True)
        return torch.einsum("cf,bfn->bnc", self.w, x), None, None


def main():
    g = torch.Generator().manual_seed(SEED)
    out = {}
    with tempfile.TemporaryDirectory() as d:
        write_blocks(d, g)
        train_loader, test_loader = RB.create_block_dataloaders(d, {6}, train_batch_size=4, test_batch_size=2, num_workers=0,
                                                                train_sampling=SAMPLING, test_sampling=None)
        out["train_blocks"], out["test_blocks"] = records(train_loader.dataset), records(test_loader.dataset)
        out["train_index"], out["test_index"] = train_loader.dataset.blocks.to(torch.int32), test_loader.dataset.blocks.to(torch.int32)
        torch.manual_seed(SEED)
        out["train_batches"] = [tuple(t.clone() for t in b) for b in train_loader] + [tuple(t.clone() for t in b) for b in train_loader]
        out["test_batches"] = [tuple(t.clone() for t in b) for b in test_loader]
    out["train_batches"] = [(p, l, n.to(torch.int64)) for p, l, n in out["train_batches"]]
    out["test_batches"] = [(p, l, n.to(torch.int64)) for p, l, n in out["test_batches"]]
    # the restatement reproduces the reference's collate on the test split (no randomness there)
    for i, (p, l, n) in enumerate(out["test_batches"]):
        op, ol, on = O.collate_blocks(out["test_blocks"][2 * i:2 * i + 2])
        assert torch.equal(p, op) and torch.equal(l, ol) and torch.equal(n, on)
    out.update(seed=SEED, sampling=SAMPLING, train_batch_size=4, test_batch_size=2)
    torch.save(out, os.path.join(ROOT, "tests", "golden", "blocks.pt"))
    print("wrote tests/golden/blocks.pt:", len(out["train_batches"]), "train batches (2 epochs),", len(out["test_batches"]), "test batches")

    # sliding-window inference
    model = TinyModel(g)
    cases = []
    for n, window, overlap in [(300, 512, 64), (1000, 256, 64), (777, 200, 50), (4096 + 100, 4096, 512), (512, 128, 0)]:
        pts = torch.rand(n, 6, generator=g)
        pred, conf = RU.predict_single_scene(model, pts, device="cpu", batch_size=window, overlap=overlap)
        mean, opred, oconf = O.predict_single_scene(model, pts, window, overlap)
        assert torch.equal(pred, opred) and torch.equal(conf, oconf)
        cases.append(dict(points=pts, window=window, overlap=overlap, pred=pred, conf=conf, mean_logits=mean))
    torch.save(dict(w=model.w.detach().clone(), cases=cases), os.path.join(ROOT, "tests", "golden", "scene_windows.pt"))
    print("wrote tests/golden/scene_windows.pt:", len(cases), "cases")


if __name__ == "__main__":
    main()
