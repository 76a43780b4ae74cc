"""Shim: the reference's `data_processing/block_datasets.py` module path, served by the HBM-resident block loader
(pcnbr_b200.block_datasets; /root/reference/data_processing/block_datasets.py:5-183).  Optional: remove this file to keep
the reference's own file-per-step DataLoader in front of the B200 ops."""
from pcnbr_b200.block_datasets import BlockS3DISDataset, collate_blocks, create_block_dataloaders   # noqa: F401
