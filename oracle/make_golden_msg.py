"""Generate tests/golden/msg.pt: multi-scale grouping composed from the UNMODIFIED reference's own functions
(models/utils/common.py: sample, group, MiniPointNet, reduce -- the reference ships no MSG class).
TEST INFRASTRUCTURE ONLY; run in the build container:

    CUDA_VISIBLE_DEVICES="" python oracle/make_golden_msg.py

Inputs are a filled cloud (every ball holds more points than its K, distinct distances), so the reference's torch.topk
has no ties among the selected keys: the script asserts raw == canonical indices before saving."""
from __future__ import annotations

import os
import sys

os.environ.setdefault("CUDA_VISIBLE_DEVICES", "")
import torch

REF = os.environ.get("PCNBR_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

from models.utils import common as RC                      # noqa: E402  (reference)
from oracle import ref_ops as O                            # noqa: E402

RADII, KS, MLPS, C, SEED = [0.08, 0.12, 0.2], [8, 16, 32], [[8, 16], [16, 16], [16, 32]], 40, 41


def main():
    g = torch.Generator().manual_seed(20261019)
    B, N, D = 2, 2048, 6
    xyz = torch.rand(B, N, 3, generator=g) * 0.3
    feat = torch.randn(B, N, D, generator=g)
    # centroids: the reference's sample() with a seeded start draw (common.py:22)
    torch.manual_seed(SEED)
    cen = RC.sample(xyz, C)
    torch.manual_seed(SEED)
    start = torch.randint(0, N, (B,), dtype=torch.int)
    # per scale: group -> MiniPointNet -> reduce, exactly as SetAbstraction.forward does (common.py:205-214)
    torch.manual_seed(SEED + 1)
    nets = [RC.MiniPointNet(3 + D, m) for m in MLPS]
    outs, tables = [], []
    for r, K, net in zip(RADII, KS, nets):
        grouped = RC.group(cen, xyz, feat, r, K, False)
        raw = O.ball_query_indices(cen, xyz, r, K, tie="raw")
        assert torch.equal(raw, O.ball_query_indices(cen, xyz, r, K, tie="canon")), "ties among the selected keys"
        assert torch.equal(grouped, O.group(cen, xyz, feat, r, K, False, idx=raw))
        tables.append(raw.to(torch.int32))
        x = net(grouped.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
        outs.append(RC.reduce(x, "max"))
    out = torch.cat(outs, dim=-1).detach()
    # the oracle composition (same seeds -> same parameters) reproduces it bit for bit
    torch.manual_seed(SEED + 1)
    om = O.SetAbstractionMSG(C, RADII, 3 + D, MLPS, KS)
    om.fps_start = start
    oc, oo = om(xyz, feat)
    assert torch.equal(oc, cen) and torch.equal(oo.detach(), out)
    torch.save(dict(coords=xyz, features=feat, start=start, seed=SEED + 1, C=C, radii=RADII, Ks=KS, mlps=MLPS, cin=3 + D,
                    centroids=cen, tables=tables, out=out), os.path.join(ROOT, "tests", "golden", "msg.pt"))
    print("wrote tests/golden/msg.pt", tuple(out.shape), [tuple(t.shape) for t in tables])


if __name__ == "__main__":
    main()
