"""Host-side mirror of the reference's data_processing/block_datasets.py on top of libpcnbr.

The reference loads one `.pt` file per block per step in DataLoader worker processes, picks `sampling` random rows on
the host, zero-pads the batch in `collate_blocks` and copies it to the GPU (block_datasets.py:5-29,117-128;
Training/training.py moves every batch with `.to(device)`).  At the point rates of the CUDA path two workers cannot feed
one GPU, let alone eight.  Here a split is read ONCE, packed back to back into HBM ((T,9) f32 + (T,14) u8: all of S3DIS is
about 14 GB of a 180 GB device) and every batch is one gather launch (`csrc/blocks.cu`) that writes the
`(points (B,N,9) f32, labels (B,N,14) u8, lengths (B,))` tuple of `collate_blocks` directly on the device.

Same names as the reference: `BlockS3DISDataset(data_dir, included_areas, sampling)` (file layout
`area_<a>/room<rr>_block<bbb>.pt`, each a `(points (n,9) f32, labels (n,14) u8)` tuple, same index order, same
exceptions), `create_block_dataloaders(...)` -> `(train_loader, test_loader)`.  The loaders iterate like the reference's
DataLoaders and consume torch's host generator in the same order as a `num_workers=0` DataLoader does (RandomSampler's
seed draw + randperm, then one `randperm(n)` / `randint(n, (S,))` per block), so with the same seed the batches are
bit-identical to the reference's; `device_sampling=True` draws the rows on the device instead (same distribution, no
host work at all).  `lengths` is int64 (the reference's uint64 cannot be indexed or compared on CUDA).
CUDA only: there is no CPU fallback."""
from __future__ import annotations

import os

import torch
from torch.utils.data import BatchSampler, RandomSampler, SequentialSampler

from . import _lib
from .ops import _stream

__all__ = ["PackedBlocks", "BlockS3DISDataset", "BlockLoader", "create_block_dataloaders", "collate_blocks"]

POINT_CHANNELS = 9


def draw_rows_host(counts, block_ids, sampling: int) -> torch.Tensor:
    """The reference's per-block draw on torch's HOST generator (block_datasets.py:119-125), block after block:
    randperm(n)[:S] when the block has more than S points, randint(n, (S,)) otherwise.  -> (B,S) int32 (CPU, pinned
    when a GPU is present)."""
    sel = torch.empty(len(block_ids), sampling, dtype=torch.int32)
    if torch.cuda.is_available():
        sel = sel.pin_memory()
    for b, blk in enumerate(block_ids):
        n = counts[blk]
        if n > sampling:
            sel[b] = torch.randperm(n)[:sampling]
        else:
            sel[b] = torch.randint(n, (sampling,))
    return sel


def plan_steps(num_blocks: int, batch_size: int, world_size: int = 1, drop_last: bool | None = None) -> int:
    """Batches EVERY rank runs per epoch.  One rank: ceil(blocks / batch) (the reference's DataLoader, drop_last=False).
    Several ranks: the short last batch is dropped (a fixed batch shape: the captured train step has static inputs) and the
    epoch is truncated to a multiple of world_size batches, so all ranks issue the same number of gradient all-reduces."""
    if drop_last is None:
        drop_last = world_size > 1
    total = num_blocks // batch_size if drop_last else (num_blocks + batch_size - 1) // batch_size
    return total // world_size


def loader_plan(num_blocks: int, counts, batch_size: int, shuffle: bool, sampling: int | None, host_draws: bool = True,
                rank: int = 0, world_size: int = 1, drop_last: bool | None = None):
    """Host side of one epoch: yields (block ids, sel (B,S) int32 or None) per batch, consuming torch's host generator
    exactly like `iter(DataLoader(dataset, batch_size, shuffle, num_workers=0))` over the reference's dataset does:
    the iterator's base-seed draw (torch/utils/data/dataloader.py, _BaseDataLoaderIter.__init__), RandomSampler's seed
    draw at the first batch, then one row draw per block in batch order.

    world_size > 1 (one process per GPU, same seed on every rank): rank r takes batches r, r + world_size, ... of that
    same epoch plan -- the batch dimension is sharded, no rank sees another rank's blocks, and every rank makes all the
    draws so the generators stay in lockstep (the union over the ranks is the single-process epoch, minus the tail that
    plan_steps() cuts so that every rank runs the same number of equally shaped steps -- a rank with one batch more would
    wait for ever in its gradient all-reduce)."""
    torch.empty((), dtype=torch.int64).random_()
    sampler = RandomSampler(range(num_blocks)) if shuffle else SequentialSampler(range(num_blocks))
    usable = plan_steps(num_blocks, batch_size, world_size, drop_last) * world_size
    for i, ids in enumerate(BatchSampler(sampler, batch_size, drop_last=False)):
        sel = draw_rows_host(counts, ids, sampling) if (sampling is not None and host_draws) else None     # all ranks draw: lockstep
        if i < usable and i % world_size == rank:
            yield ids, sel


class PackedBlocks:
    """Every block of a split packed back to back in HBM.

    points (T,9) f32, labels (T,L) u8, block_start (nblocks+1) int64 on the device; `lengths_host` the same offsets as
    a Python list (batch shapes are decided without a device round trip)."""

    def __init__(self, blocks, device="cuda"):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("pcnbr: PackedBlocks lives in GPU memory (this build has no CPU fallback)")
        if len(blocks) == 0:
            raise ValueError("PackedBlocks needs at least one block")
        pts, labs, counts = [], [], []
        for p, l in blocks:
            if p.dim() != 2 or p.shape[1] != POINT_CHANNELS or l.dim() != 2 or l.shape[0] != p.shape[0]:
                raise ValueError(f"block records must be (n,{POINT_CHANNELS}) points and (n,L) labels, got {tuple(p.shape)} / {tuple(l.shape)}")
            pts.append(p.to(torch.float32))
            labs.append(l.to(torch.uint8))
            counts.append(int(p.shape[0]))
        self.num_label_channels = int(labs[0].shape[1])
        if any(l.shape[1] != self.num_label_channels for l in labs):
            raise ValueError("all blocks must have the same number of label channels")
        starts = [0]
        for c in counts:
            starts.append(starts[-1] + c)
        self.counts_host = counts
        self.points = torch.cat(pts).contiguous().to(device)
        self.labels = torch.cat(labs).contiguous().to(device)
        self.block_start = torch.tensor(starts, dtype=torch.int64, device=device)
        self.device = device

    def __len__(self):
        return len(self.counts_host)

    def nbytes(self) -> int:
        return self.points.numel() * 4 + self.labels.numel()

    # ------------------------------------------------------------------ row draws (block_datasets.py:119-125)
    def draw_host(self, block_ids, sampling: int) -> torch.Tensor:
        return draw_rows_host(self.counts_host, block_ids, sampling)

    def draw_device(self, block_ids, sampling: int, generator: torch.Generator | None = None) -> torch.Tensor:
        """The same distribution drawn on the device (no host work, no copy): blocks with more than S points take the S
        smallest of n iid uniform keys (a uniform S-subset in uniform order = randperm(n)[:S]); the others take
        floor(u * n).  -> (B,S) int32 (CUDA)."""
        n = torch.tensor([self.counts_host[b] for b in block_ids], dtype=torch.int64).to(self.device, non_blocking=True)
        B, nmax = len(block_ids), max(self.counts_host[b] for b in block_ids)
        with_repl = (torch.rand(B, sampling, device=self.device, generator=generator, dtype=torch.float64) * n[:, None]).long()
        with_repl = torch.minimum(with_repl, n[:, None] - 1)
        if nmax <= sampling:
            return with_repl.to(torch.int32)
        keys = torch.rand(B, nmax, device=self.device, generator=generator)
        keys = torch.where(torch.arange(nmax, device=self.device)[None, :] < n[:, None], keys, torch.full_like(keys, 2.0))
        without = torch.topk(keys, sampling, dim=1, largest=False, sorted=True).indices
        return torch.where((n > sampling)[:, None], without, with_repl).to(torch.int32)

    # ------------------------------------------------------------------ the batch
    def batch(self, block_ids, sampling: int | None = None, sel: torch.Tensor | None = None):
        """-> (points (B,N,9) f32, labels (B,N,L) u8, lengths (B,) int64), all on the device: what the reference's
        DataLoader yields for these blocks (`__getitem__` per block + `collate_blocks`).

        sampling given: N = sampling, rows `sel` (B,N) int32 (drawn with draw_host when not passed).
        sampling None : every row of every block in order, zero padded to the longest block of the batch."""
        block_ids = [int(b) for b in block_ids]
        if any(b < 0 or b >= len(self) for b in block_ids):
            raise IndexError(f"block index out of range [0, {len(self)})")
        B = len(block_ids)
        if sampling is not None:
            if sel is None:
                sel = self.draw_host(block_ids, sampling)
            sel = sel.to(device=self.device, dtype=torch.int32, non_blocking=True).contiguous()
            if tuple(sel.shape) != (B, sampling):
                raise ValueError(f"sel must be ({B},{sampling}), got {tuple(sel.shape)}")
            N = sampling
        else:
            N = max(self.counts_host[b] for b in block_ids)
        ids = torch.tensor(block_ids, dtype=torch.int32).to(self.device, non_blocking=True)
        L = self.num_label_channels
        pts = torch.empty(B, N, POINT_CHANNELS, dtype=torch.float32, device=self.device)
        lab = torch.empty(B, N, L, dtype=torch.uint8, device=self.device)
        lens = torch.empty(B, dtype=torch.int64, device=self.device)
        _lib.call("pcnbr_block_batch", self.points.data_ptr(), self.labels.data_ptr(), self.block_start.data_ptr(),
                  ids.data_ptr(), sel.data_ptr() if sel is not None else None, B, N, L, pts.data_ptr(), lab.data_ptr(),
                  lens.data_ptr(), _stream())
        return pts, lab, lens


def collate_blocks(batch, device="cuda"):
    """block_datasets.py:5-29 for a list of (points (n,9), labels (n,L)) samples that are NOT part of a packed split
    (e.g. blocks cut from a new scene): packs them and runs the same gather.  -> device tensors."""
    return PackedBlocks(batch, device).batch(range(len(batch)), None)


class BlockS3DISDataset:
    """The S3DIS block dataset (block_datasets.py:33-131), read once and kept packed in HBM.

    Same constructor, same file layout, same index order and the same exceptions as the reference; `__getitem__`
    returns device tensors."""

    def __init__(self, data_dir: str, included_areas: set[int], sampling: int | None = None, device="cuda"):
        if not os.path.exists(data_dir):
            raise FileNotFoundError(f'Data directory "{data_dir}" does not exist.')
        if any([a < 1 or a > 6 for a in included_areas]):
            raise ValueError(f'Included areas can only contain values from the range [1, 6], got {included_areas}.')
        self.blocks = self._create_block_index(data_dir, included_areas)
        self.data_dir = data_dir
        self.sampling = sampling
        records = []
        for area_index, room_index, block_index in self.blocks.tolist():
            records.append(torch.load(os.path.join(data_dir, f'area_{area_index}', f'room{room_index:02d}_block{block_index:03d}.pt')))
        self.packed = PackedBlocks(records, device)

    @staticmethod
    def _create_block_index(data_dir: str, included_areas: set[int]) -> torch.Tensor:
        """(nblocks,3) (area, room, block), sorted as block_datasets.py:56-93."""
        blocks = []
        for area_index in sorted(list(included_areas)):
            area_dir = os.path.join(data_dir, f'area_{area_index}')
            if not os.path.exists(area_dir):
                raise FileNotFoundError(f'Directory for area {area_index} does not exist.')
            indices = [name.replace('room', '').replace('block', '').replace('.pt', '').split('_') for name in os.listdir(area_dir)]
            if len(indices) == 0:
                raise FileNotFoundError(f'Directory for area {area_index} does not contain any blocks.')
            blocks += sorted((area_index, int(room), int(block)) for room, block in indices)
        return torch.tensor(blocks, dtype=torch.int32)

    def __len__(self) -> int:
        return self.blocks.shape[0]

    def __getitem__(self, index: int):
        pts, lab, _ = self.packed.batch([index], self.sampling)
        n = pts.shape[1]
        return pts.view(n, POINT_CHANNELS), lab.view(n, -1)


class BlockLoader:
    """Iterates like `DataLoader(dataset, batch_size, shuffle, collate_fn=collate_blocks)` of the reference
    (block_datasets.py:160-175) and yields device batches.  Index order and host-generator consumption are those of a
    num_workers=0 DataLoader (torch's own RandomSampler / BatchSampler produce the order)."""

    def __init__(self, dataset: BlockS3DISDataset, batch_size: int, shuffle: bool, device_sampling: bool = False,
                 generator: torch.Generator | None = None, rank: int = 0, world_size: int = 1, drop_last: bool | None = None):
        if not 0 <= rank < world_size:
            raise ValueError(f"rank {rank} outside [0, {world_size})")
        self.dataset, self.batch_size, self.shuffle = dataset, batch_size, shuffle
        self.device_sampling, self.generator = device_sampling, generator
        self.rank, self.world_size, self.drop_last = rank, world_size, drop_last

    def __len__(self) -> int:
        return plan_steps(len(self.dataset), self.batch_size, self.world_size, self.drop_last)

    def __iter__(self):
        packed, S = self.dataset.packed, self.dataset.sampling
        for ids, sel in loader_plan(len(packed), packed.counts_host, self.batch_size, self.shuffle, S,
                                    host_draws=not self.device_sampling, rank=self.rank, world_size=self.world_size,
                                    drop_last=self.drop_last):
            if S is not None and self.device_sampling:
                sel = packed.draw_device(ids, S, self.generator)
            yield packed.batch(ids, S, sel)


def create_block_dataloaders(data_dir: str, test_areas: set[int], train_batch_size: int = 4, test_batch_size: int = 4,
                             num_workers: int = 4, train_sampling: int | None = 4096, test_sampling: int | None = None,
                             train_shuffle: bool = True, test_shuffle: bool = False, device="cuda",
                             device_sampling: bool = False, rank: int = 0, world_size: int = 1):
    """block_datasets.py:134-177 -> (train_loader, test_loader).  `num_workers` is accepted and ignored: there is no
    per-step host work left to parallelise.  rank / world_size shard the TRAIN batches across data-parallel ranks (the
    test loader stays whole on every rank, as an unsharded evaluation would)."""
    areas = {1, 2, 3, 4, 5, 6}
    train = BlockS3DISDataset(data_dir, areas - test_areas, train_sampling, device)
    test = BlockS3DISDataset(data_dir, test_areas, test_sampling, device)
    return (BlockLoader(train, train_batch_size, train_shuffle, device_sampling, rank=rank, world_size=world_size),
            BlockLoader(test, test_batch_size, test_shuffle, device_sampling))
