"""Golden vectors for the LENGTH-AWARE forms, from the UNMODIFIED reference (/root/reference).

TEST INFRASTRUCTURE ONLY.  Run in the build container (the reference does not travel to the GPU box):

    CUDA_VISIBLE_DEVICES="" python oracle/make_golden_lengths.py

The reference has no length-aware path: its evaluation loop feeds zero-padded batches and the padding takes part in
FPS / grouping (data_processing/block_datasets.py:19-25, Training/training.py:112).  The length-aware contract of this
repo is "the real rows get what the reference computes when the cloud is passed ALONE, unpadded", so every fixture here
is the reference run once per cloud on `batch[b:b+1, :lengths[b]]`, stored next to the zero-padded batch (built exactly
as collate_blocks builds it).  Inputs are filled clouds (no ties among the selected keys: raw topk == canonical order,
asserted before saving).
"""
from __future__ import annotations

import os
import sys

os.environ.setdefault("CUDA_VISIBLE_DEVICES", "")   # reference dgcnn.py:39 picks 'cuda' if available
import torch

REF = os.environ.get("PCNBR_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

from models.utils import common as RC                      # noqa: E402  (reference)
from models.dgcnn import dgcnn as RD                       # noqa: E402  (reference)
from models.PointNetpp.PointNetpp import PointNetpp as RefPointNetpp   # noqa: E402

from oracle import ref_ops as O                            # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def padded(clouds, width):
    """collate_blocks (block_datasets.py:16-25): zeros behind every cloud up to the longest one."""
    B, N = len(clouds), max(c.shape[0] for c in clouds)
    out = torch.zeros(B, N, width)
    for b, c in enumerate(clouds):
        out[b, :c.shape[0]] = c
    return out


def main():
    g = torch.Generator().manual_seed(20261019)
    fx = {}
    lengths = [2048, 1500, 1100]
    clouds = [torch.rand(n, 3, generator=g) * 0.2 + torch.tensor([4.0, 9.0, 0.0]) for n in lengths]
    xyz = padded(clouds, 3)

    # ---- sample(): per cloud, alone; the start draw replayed from the same seed (common.py:22)
    picks, starts = [], []
    for b, c in enumerate(clouds):
        torch.manual_seed(100 + b)
        picks.append(RC.sample(c[None], 256)[0])
        torch.manual_seed(100 + b)
        starts.append(torch.randint(0, c.shape[0], (1,), dtype=torch.int))
    cen = torch.stack(picks)
    fx["sample"] = dict(start=torch.cat(starts), C=256, coords=cen)

    # ---- group() (ball query + gather) per cloud against its own picks
    feats = [torch.randn(n, 6, generator=g) for n in lengths]
    rows = []
    for b, (c, f) in enumerate(zip(clouds, feats)):
        ref = RC.group(cen[b:b + 1], c[None], f[None], 0.1, 32, True)
        assert torch.equal(ref, O.group(cen[b:b + 1], c[None], f[None], 0.1, 32, True, tie="canon")), "ties in the group fixture"
        rows.append(ref[0])
    fx["group"] = dict(features=padded(feats, 6), r=0.1, K=32, out=torch.stack(rows))

    # ---- interpolate(): the cloud's own points are the queries (padding query rows have no reference value)
    coarse = torch.randn(3, 256, 16, generator=g)
    up = []
    for b, c in enumerate(clouds):
        ref = RC.interpolate(coarse[b:b + 1], c[None], cen[b:b + 1])
        assert torch.equal(ref, O.interpolate(coarse[b:b + 1], c[None], cen[b:b + 1], tie="canon"))
        up.append(ref[0])
    fx["interpolate"] = dict(points=coarse, out=padded(up, 16))

    # ---- knn() in feature space, F = 3 and F = 64, per cloud (the |x|^2 summation order depends on the cloud's length)
    for F in (3, 64):
        xs = [torch.randn(F, n, generator=g) for n in lengths]
        idx = []
        for x in xs:
            ref = RD.knn(x[None], 20)
            assert torch.equal(ref, O.knn(x[None], 20, "canon")), "ties in the knn fixture"
            idx.append(ref[0].int())
        xp = torch.zeros(3, F, max(lengths))
        for b, x in enumerate(xs):
            xp[b, :, :x.shape[1]] = x
        fx[f"knn_F{F}"] = dict(x=xp, k=20, idx=padded([i.float() for i in idx], 20).int())

    # ---- PointNet++ SSG in eval mode: logits per cloud, alone (FPS start draws replayed)
    rgb = [torch.randint(0, 256, (n, 3), generator=g).float() for n in lengths]
    x9 = [torch.cat([c, col, c - c.mean(dim=0, keepdim=True)], dim=-1) for c, col in zip(clouds, rgb)]
    torch.manual_seed(41)
    net = RefPointNetpp(13).eval()
    torch.manual_seed(41)
    net_o = O.PointNetpp(13, tie="canon").eval()
    logits, draws = [], []
    with torch.no_grad():
        for b, x in enumerate(x9):
            n = x.shape[0]
            torch.manual_seed(500 + b)
            lo = net(x[None])
            torch.manual_seed(500 + b)
            st = [torch.randint(0, n_src, (1,), dtype=torch.int) for n_src in (n, 1024, 256, 64)]
            for sa_o, s in zip((net_o.sa1, net_o.sa2, net_o.sa3, net_o.sa4), st):
                sa_o.fps_start = s
            assert torch.equal(lo, net_o(x[None])), "restated PointNet++ differs from the reference"
            logits.append(lo[0])
            draws.append(torch.cat(st))
    fx["pointnetpp_eval"] = dict(x=padded(x9, 9), seed=41, fps_starts=torch.stack(draws).t().contiguous(),   # (level, cloud)
                                 logits=padded(logits, 13))

    fx["lengths"] = torch.tensor(lengths, dtype=torch.int64)
    fx["xyz"] = xyz
    path = os.path.join(OUT, "lengths.pt")
    torch.save(fx, path)
    print(f"wrote {path}: {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
