"""Synthetic S3DIS-shaped blocks in the layout the reference's block dataloader yields
(data_processing/block_datasets.py:5-29): points (B,N,9) f32, one-hot labels (B,N,classes) u8,
lengths (B,).  Channels: 0-2 absolute xyz [m] (block origin + 1 m x 1 m x 3 m), 3-5 raw rgb 0..255,
6-8 xyz minus the block centre (data_processing/preprocess_dataset.py:73-90)."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def s3dis_blocks(B: int, N: int = 4096, seed: int = 0, classes: int = 13):
    g = torch.Generator().manual_seed(seed)
    origin = torch.randint(0, 20, (B, 1, 2), generator=g).float()
    xy = origin + torch.rand(B, N, 2, generator=g)
    z = 3.0 * torch.rand(B, N, 1, generator=g)
    xyz = torch.cat([xy, z], dim=-1)
    rgb = torch.randint(0, 256, (B, N, 3), generator=g).float()
    zc = (z.amin(dim=1, keepdim=True) + z.amax(dim=1, keepdim=True)) / 2
    centre = torch.cat([origin + 0.5, zc], dim=-1)
    pts = torch.cat([xyz, rgb, xyz - centre], dim=-1)
    lab = F.one_hot(torch.randint(0, classes, (B, N), generator=g), classes).to(torch.uint8)
    return pts, lab, torch.full((B,), N, dtype=torch.int64)
