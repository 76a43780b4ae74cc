"""CPU, world_size 2 over gloo: the data-parallel plumbing of the train step (batch sharding by cloud,
flat gradient bucket, one all-reduce, parameter broadcast).  The hot-path kernels need no collective."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out, steal=False, overlap=True):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import __graft_entry__ as ge
    pkg = ge.load_package()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                      # different init per rank on purpose
    net = torch.nn.Sequential(torch.nn.Linear(9, 16), torch.nn.ReLU(), torch.nn.Linear(16, 13))
    pkg.train.broadcast_parameters(net)                # -> rank 0's parameters everywhere
    bucket = pkg.train.FlatGradBucket(net, steal_grads=steal, overlap=overlap, late_fraction=0.5)
    assert bucket.overlap == overlap and (bucket.split == 2 if overlap else bucket.split == 0)     # late part = the first Linear
    pts, lab, lens = pkg.synthetic.s3dis_blocks(4, 64, seed=0)
    sl = pkg.train.shard_batch(4, rank, world)
    bucket.zero()
    loss = pkg.train.masked_onehot_cross_entropy(net(pts[sl]), lab[sl], lens[sl])
    loss.backward()
    assert (bucket._pending is not None) == overlap    # the head's all-reduce was started from the backward hook
    bucket.all_reduce_mean()
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(bucket.params, bucket.views))   # grads live in the bucket
    # a second step through the same bucket (hooks re-arm, views are reused)
    first = bucket.flat.clone()
    bucket.zero()
    pkg.train.masked_onehot_cross_entropy(net(pts[sl]), lab[sl], lens[sl]).backward()
    bucket.all_reduce_mean()
    assert torch.allclose(bucket.flat, first, rtol=1e-6, atol=1e-8)
    out[rank] = (bucket.flat.clone(), torch.cat([p.detach().flatten() for p in net.parameters()]), (sl.start, sl.stop))
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("steal,overlap", [(False, True), (True, True), (True, False)])
def test_flat_bucket_allreduce_matches_full_batch(pkg, steal, overlap):
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out, steal, overlap), nprocs=world, join=True)
        res = dict(out)
    g0, p0, s0 = res[0]
    g1, p1, s1 = res[1]
    assert torch.equal(p0, p1)                          # broadcast worked
    assert torch.equal(g0, g1)                          # both ranks hold the reduced gradient
    assert (s0, s1) == ((0, 2), (2, 4))
    # single-process reference: mean over the two shards' losses == full-batch loss (equal shard sizes)
    torch.manual_seed(100)
    net = torch.nn.Sequential(torch.nn.Linear(9, 16), torch.nn.ReLU(), torch.nn.Linear(16, 13))
    pts, lab, lens = pkg.synthetic.s3dis_blocks(4, 64, seed=0)
    pkg.train.masked_onehot_cross_entropy(net(pts), lab, lens).backward()
    full = torch.cat([p.grad.flatten() for p in net.parameters()])
    assert torch.allclose(g0, full, rtol=1e-5, atol=1e-7)


def test_bucket_views_and_zero(pkg):
    net = torch.nn.Linear(4, 3)
    b = pkg.train.FlatGradBucket(net)
    assert b.flat.numel() == 15 and all(p.grad.data_ptr() >= b.flat.data_ptr() for p in net.parameters())
    net(torch.ones(2, 4)).sum().backward()
    assert b.flat.abs().sum() > 0                       # autograd accumulated INTO the bucket views
    b.zero()
    assert all(float(p.grad.abs().sum()) == 0 for p in net.parameters())
    assert [pkg.train.shard_batch(16, r, 8) for r in (0, 7)] == [slice(0, 2), slice(14, 16)]
    assert pkg.train.shard_batch(5, 3, 4) == slice(5, 5)


def test_masked_loss_matches_reference_formula(pkg):
    """Training/train_model.py:15-57: mean of -sum(onehot * log_softmax) over positions < pad_start."""
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(3, 10, 13, generator=g)
    lab = torch.nn.functional.one_hot(torch.randint(0, 13, (3, 10), generator=g), 13).to(torch.uint8)
    lens = torch.tensor([10, 4, 0])
    want = sum(torch.nn.functional.cross_entropy(logits[b, :n], lab[b, :n].argmax(-1), reduction="sum")
               for b, n in enumerate(lens.tolist())) / lens.sum()
    got = pkg.train.masked_onehot_cross_entropy(logits, lab, lens)
    assert torch.allclose(got, want, rtol=1e-6)
