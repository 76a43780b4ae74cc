"""CPU: the oracle restatements of the block dataloader / sliding-window inference against the golden vectors made
from the unmodified reference (oracle/make_golden_blocks.py), and the HOST side of the packed loader (batch order and
row draws): with the same seed it must consume torch's generator exactly like the reference's DataLoader."""
import torch

from oracle import ref_ops as O


class TinyModel(torch.nn.Module):
    """The stand-in model of oracle/make_golden_blocks.py (weights from the fixture)."""
    num_classes = 13

    def __init__(self, w):
        super().__init__()
        self.w = torch.nn.Parameter(w.clone())

    def forward(self, x):
        x = x - x.mean(dim=2, keepdim=True)
        return torch.einsum("cf,bfn->bnc", self.w, x), None, None


def same_batch(a, b):
    return all(torch.equal(x, y) for x, y in zip(a, b)) and len(a) == len(b)


def test_oracle_collate_matches_reference(golden):
    g = golden("blocks")
    bs = g["test_batch_size"]
    for i, want in enumerate(g["test_batches"]):
        assert same_batch(O.collate_blocks(g["test_blocks"][bs * i:bs * (i + 1)]), want)
    assert [int(n) for n in g["test_batches"][-1][2]] == [g["test_blocks"][-1][0].shape[0]]        # ragged last batch


def test_oracle_getitem_draws_like_reference(golden):
    g = golden("blocks")
    # two branches of block_datasets.py:119-125: above the sampling size -> a subset, at or below -> with replacement
    big = next(b for b in g["train_blocks"] if b[0].shape[0] > g["sampling"])
    small = next(b for b in g["train_blocks"] if b[0].shape[0] <= g["sampling"])
    torch.manual_seed(1)
    p, l = O.block_getitem(*big, g["sampling"])
    assert p.shape == (g["sampling"], 9) and len({tuple(r.tolist()) for r in p}) == g["sampling"]
    p, l = O.block_getitem(*small, g["sampling"])
    assert p.shape == (g["sampling"], 9) and l.shape == (g["sampling"], 14)


def test_loader_plan_reproduces_reference_batches(pkg, golden):
    g = golden("blocks")
    blocks = g["train_blocks"]
    counts = [b[0].shape[0] for b in blocks]
    torch.manual_seed(g["seed"])
    got = []
    for _ in range(2):                                                          # two epochs, as the fixture
        for ids, sel in pkg.block_datasets.loader_plan(len(blocks), counts, g["train_batch_size"], True, g["sampling"]):
            assert sel.dtype == torch.int32 and sel.shape == (len(ids), g["sampling"])
            got.append(O.gather_block_batch(blocks, ids, sel))
    assert len(got) == len(g["train_batches"])
    for a, b in zip(got, g["train_batches"]):
        assert same_batch(a, b)


def test_loader_plan_sequential_no_sampling(pkg, golden):
    g = golden("blocks")
    blocks = g["test_blocks"]
    plan = list(pkg.block_datasets.loader_plan(len(blocks), [b[0].shape[0] for b in blocks], g["test_batch_size"], False, None))
    assert [ids for ids, _ in plan] == [[0, 1], [2]] and all(sel is None for _, sel in plan)
    for (ids, sel), want in zip(plan, g["test_batches"]):
        assert same_batch(O.gather_block_batch(blocks, ids, sel), want)


def test_oracle_scene_windows_match_reference(pkg, golden):
    g = golden("scene_windows")
    model = TinyModel(g["w"])
    for case in g["cases"]:
        mean, pred, conf = O.predict_single_scene(model, case["points"], case["window"], case["overlap"])
        assert torch.equal(pred, case["pred"]) and torch.equal(conf, case["conf"]) and torch.equal(mean, case["mean_logits"])
        # the window list of the host layer is the reference's loop
        n = case["points"].shape[0]
        if n > case["window"]:
            step = case["window"] - case["overlap"]
            assert pkg.dgcnn_utils.scene_windows(n, case["window"], case["overlap"]) == \
                [(s, min(s + case["window"], n)) for s in range(0, n, step)]


def test_block_store_and_scene_inference_need_cuda(pkg):
    import pytest
    pts = torch.rand(10, 9)
    lab = torch.zeros(10, 14, dtype=torch.uint8)
    with pytest.raises(RuntimeError):
        pkg.block_datasets.PackedBlocks([(pts, lab)], device="cpu")
    with pytest.raises(RuntimeError):
        pkg.dgcnn_utils.predict_single_scene(TinyModel(torch.zeros(13, 6)), torch.rand(10, 6), device="cpu")


def test_loader_plan_shards_batches_across_ranks(pkg, golden):
    """world_size 3: the ranks' batches are disjoint, interleave to the single-process epoch (up to the tail cut for equal
    step counts), and every rank leaves the host generator in the same state (lockstep draws)."""
    g = golden("blocks")
    blocks = g["train_blocks"]
    counts = [b[0].shape[0] for b in blocks]
    plan = pkg.block_datasets.loader_plan
    torch.manual_seed(7)
    whole = list(plan(len(blocks), counts, 2, True, g["sampling"]))
    state = torch.get_rng_state()
    per_rank = []
    for r in range(3):
        torch.manual_seed(7)
        per_rank.append(list(plan(len(blocks), counts, 2, True, g["sampling"], rank=r, world_size=3)))
        assert torch.equal(torch.get_rng_state(), state)
    steps = pkg.block_datasets.plan_steps(len(blocks), 2, 3)
    assert [len(p) for p in per_rank] == [steps] * 3 and steps == (len(blocks) // 2) // 3
    for i in range(3 * steps):
        ids, sel = whole[i]
        rid, rsel = per_rank[i % 3][i // 3]
        assert rid == ids and torch.equal(rsel, sel)
    seen = [tuple(ids) for p in per_rank for ids, _ in p]
    assert len(set(seen)) == len(seen)


def test_every_rank_runs_the_same_number_of_equally_shaped_steps(pkg):
    """with n_batches % world != 0 (or a short last batch) ranks used to run different step counts and the
    epoch ended in a hung all-reduce.  Every (blocks, batch, world) combination now gives all ranks the same number of
    batches, all of them full."""
    plan, steps_of = pkg.block_datasets.loader_plan, pkg.block_datasets.plan_steps
    for nblocks in (5, 7, 16, 23):
        counts = [100 + i for i in range(nblocks)]
        for bs in (2, 3, 4):
            for world in (2, 3, 4, 8):
                lens = []
                for r in range(world):
                    torch.manual_seed(1)
                    batches = list(plan(nblocks, counts, bs, True, 64, rank=r, world_size=world))
                    assert all(len(ids) == bs for ids, _ in batches)
                    lens.append(len(batches))
                assert lens == [steps_of(nblocks, bs, world)] * world
    # one rank keeps the reference's DataLoader semantics: the short last batch is yielded
    torch.manual_seed(1)
    single = list(plan(7, [50] * 7, 2, False, None))
    assert [len(ids) for ids, _ in single] == [2, 2, 2, 1] and steps_of(7, 2) == 4
