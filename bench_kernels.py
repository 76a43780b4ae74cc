#!/usr/bin/env python
"""bench_kernels.py -- the kernel sweep of BASELINE.json configs[4]: FPS / ball query / kNN / group / pool /
interpolate / edge features / scatter-add at N = 4k-100k, k = 16-32, C = 3-256, one GPU (replicas only at N GPUs).

    python bench_kernels.py [--quick] [--iters 10] [--md profiles/rX_kernel_sweep.md]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 bench_kernels.py --replicas    # "1 vs 8 GPUs"

Every libpcnbr KERNEL is timed by the library's own profiler (csrc/prof.cu: CUDA events on the launch stream
around each kernel), with L2 flushed between iterations (a 512 MB memset), after 3 warm-up iterations.  The roofline
fraction is ALGORITHMIC bytes (or flops) -- stated by the launch site from the SURVEY.md 8d formulas -- / time /
the measured peak in MEASURED_PEAKS.json.  Prints one JSON line per (op, shape, kernel) and a markdown table.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def replica_sweep(args):
    """BASELINE configs[4] "1 vs 8 GPUs": every op of the path is per-cloud, so N GPUs run N independent replicas with no
    exchange (SURVEY.md 8e "replicas only").  One process per GPU (python -m torch.distributed.run --nproc-per-node N
    bench_kernels.py --replicas); every rank times the same op calls on its own clouds, the reported time is the MAX over
    ranks (device events, L2 flushed between calls) and the throughput the aggregate clouds/s of all replicas."""
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import __graft_entry__ as ge
    pkg = ge.load_package()
    ops = pkg.ops
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator().manual_seed(rank)
    rows = []

    def timed(op, shape, B, fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = 0.0
        for _ in range(args.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
        t = torch.tensor([ms / args.iters], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        row = {"op": op, "shape": shape, "n_gpus": world, "us_per_call_max_over_ranks": round(1e3 * t.item(), 1),
               "clouds_per_s_aggregate": round(world * B / (t.item() / 1e3), 1)}
        rows.append(row)
        if rank == 0:
            print(json.dumps(row), flush=True)

    for N in (4096, 24000, 100000):
        B = 32 if N <= 4096 else max(2, min(8, (1 << 18) // N))
        pts, _, _ = pkg.synthetic.s3dis_blocks(B, N, seed=rank * 31 + N % 997)
        xyz = pts[:, :, :3].contiguous().to(dev)
        start = torch.zeros(B, dtype=torch.int32, device=dev)
        sh = f"B={B} N={N}"
        timed("fps C=1024", sh, B, lambda: ops.farthest_point_sample(xyz, 1024, start))
        _, cen = ops.farthest_point_sample(xyz, 1024, start, return_coords=True)
        timed("ball_query r=0.1 K=32", sh, B, lambda: ops.query_ball_point(0.1, 32, xyz, cen))
        nbr = ops.NeighborIndex(ops.query_ball_point(0.1, 32, xyz, cen), N)
        f = torch.randn(B, N, 64, generator=g).to(dev)
        timed("group D=64 K=32", sh, B, lambda: ops.group_points(xyz, f, cen, nbr, 0.1, pad4=True))
        timed("knn3 M=1024", sh, B, lambda: ops.knn_points(xyz, cen, 3))
        i3, d3 = ops.knn_points(xyz, cen, 3)
        coarse = torch.randn(B, 1024, 128, generator=g).to(dev)
        n3 = ops.NeighborIndex(i3, 1024)
        timed("interpolate D=128", sh, B, lambda: ops.three_interpolate(coarse, n3, d3))
        del xyz, f, coarse
    for (N, F, k) in ((4096, 64, 20), (16384, 64, 20), (4096, 64, 32), (4096, 3, 16)):
        B = 16 if N <= 4096 else 4
        xt = torch.randn(B, N, F, generator=g).to(dev)
        sh = f"B={B} N={N} F={F} k={k}"
        timed("knn_feature", sh, B, lambda: ops.knn_graph(xt.transpose(1, 2), k))
        nbr = ops.NeighborIndex(ops.knn_graph(xt.transpose(1, 2), k), N)
        nbr.csr()
        if F == 64:
            PQ = torch.randn(B, N, 128, generator=g).to(dev)
            bn = torch.nn.BatchNorm2d(64).to(dev)
            with torch.no_grad():
                timed("edgeconv_fused O=64", sh, B, lambda: ops.edgeconv_fused(PQ, nbr, bn, 0.2))
    if rank == 0 and args.md:
        with open(args.md, "w") as fh:
            fh.write(f"# kernel sweep, {world} replica(s) ({torch.cuda.get_device_name(0)}): op calls through the C ABI, max over ranks, aggregate clouds/s\n\n")
            fh.write("| op | shape per GPU | GPUs | us per call (max over ranks) | clouds/s (all replicas) |\n|---|---|---:|---:|---:|\n")
            for r in rows:
                fh.write(f"| {r['op']} | {r['shape']} | {r['n_gpus']} | {r['us_per_call_max_over_ranks']} | {r['clouds_per_s_aggregate']} |\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--replicas", action="store_true",
                    help="op-level sweep as independent replicas, one per GPU (run under torch.distributed.run for N > 1): aggregate clouds/s")
    ap.add_argument("--quick", action="store_true", help="BASELINE shapes only (no N sweep)")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--md", default="")
    ap.add_argument("--only", default="", help="comma-separated op groups: pointnet, dgcnn, gemm, io (default all)")
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("bench_kernels.py: no CUDA device; the hot path has no CPU fallback")
    if args.replicas:
        return replica_sweep(args)
    import __graft_entry__ as ge
    pkg = ge.load_package()
    lib, ops = pkg._lib, pkg.ops
    lib.load()
    dev = torch.device("cuda:0")
    peaks = bench.load_peaks()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    rows = []

    def run(op, shape, fn, bwd=False):
        """fn() -> tensor (its .sum().backward() is also run when bwd)."""
        def once():
            out = fn()
            if bwd:
                out.backward(torch.ones_like(out))
        for _ in range(3):
            once()
        torch.cuda.synchronize()
        lib.prof_enable(True)
        for _ in range(args.iters):
            flush.zero_()
            once()
        kernels = lib.prof_collect()
        lib.prof_enable(False)
        for name, d in kernels.items():
            us = 1e3 * d["ms"] / d["calls"]
            b = bench.kernel_bound(name, d, peaks)
            row = {"op": op, "shape": shape, "kernel": name, "us_per_launch": round(us, 2), "launches_per_call": d["calls"] // args.iters,
                   "bound": b["bound"], "achieved": round(b["achieved"], 2), "unit": b["unit"],
                   "frac": round(b["frac"], 4), "gflop_per_s": round(d["flops"] / 1e9 / (d["ms"] / 1e3), 1)}
            rows.append(row)
            print(json.dumps(row), flush=True)

    g = torch.Generator().manual_seed(0)
    only = set(filter(None, args.only.split(",")))
    want = lambda grp: not only or grp in only

    def cloud(B, N):
        pts, _, _ = pkg.synthetic.s3dis_blocks(B, N, seed=N % 997)
        return pts[:, :, :3].contiguous().to(dev)

    def feats(B, N, D):
        return torch.randn(B, N, D, generator=g).to(dev)

    # ---- PointNet++ SSG shapes at B=32 (configs[0]/[2]) and the N sweep (configs[4])
    pn_levels = [(4096, 1024, 0.1, 32, 6, 64), (1024, 256, 0.2, 32, 64, 128), (256, 64, 0.4, 32, 128, 256), (64, 16, 0.8, 32, 256, 512)]
    sweep_n = [] if args.quick else [8192, 16384, 32768, 65536, 100000]
    for (N, C, r, K, D, O) in (pn_levels + [(n, 1024, 0.1, 32, 6, 64) for n in sweep_n]) if want("pointnet") else []:
        B = 32 if N <= 4096 else max(1, min(8, (1 << 18) // N))
        xyz = cloud(B, N)
        f = feats(B, N, D).requires_grad_(True)
        start = torch.zeros(B, dtype=torch.int32, device=dev)
        sh = f"B={B} N={N} C={C} K={K} D={D}"
        run("fps", sh, lambda: ops.farthest_point_sample(xyz, C, start).float())
        _, cen = ops.farthest_point_sample(xyz, C, start, return_coords=True)
        run("ball_query", sh + f" r={r}", lambda: ops.query_ball_point(r, K, xyz, cen).float())
        if K // 2 <= N:                                    # the MSG pair of this level: (r/2, K/2) + (r, K) from one scan
            run("ball_query_msg", sh + f" r={r / 2}+{r} K={K // 2}+{K}",
                lambda: ops.query_ball_point_multi([r / 2, r], [K // 2, K], xyz, cen)[1].float())
        nbr = ops.NeighborIndex(ops.query_ball_point(r, K, xyz, cen), N)
        run("csr_build", sh, lambda: (setattr(nbr, "_csr", None), nbr.csr()[1].float())[1])
        run("group(+bwd)", sh, lambda: ops.group_points(xyz, f, cen, nbr, r), bwd=True)
        conv_out = torch.randn(B, C, K, O, generator=g).to(dev).requires_grad_(True)
        run("maxpool(+bwd)", f"B={B} C={C} K={K} D'={O}", lambda: ops.max_pool_neighbors(conv_out, 2), bwd=True)
        coarse = feats(B, C, O).requires_grad_(True)
        run("knn3", f"B={B} N={N} M={C}", lambda: ops.knn_points(xyz, cen, 3)[1])
        i3, d3 = ops.knn_points(xyz, cen, 3)
        n3 = ops.NeighborIndex(i3, C)
        n3.csr()
        run("interpolate(+bwd)", f"B={B} N={N} M={C} D={O}", lambda: ops.three_interpolate(coarse, n3, d3), bwd=True)
        del xyz, f, conv_out, coarse

    # ---- DGCNN shapes at B=16, k=20 (configs[1]) and the sweep
    dg_shapes = [(4096, 3, 20), (4096, 64, 20)] + [(n, 64, kk) for n in sweep_n[:4] for kk in (20,)] + ([] if args.quick else [(4096, 64, 16), (4096, 64, 32), (4096, 32, 20)])
    for (N, F, k) in dg_shapes if want("dgcnn") else []:
        B = 16 if N <= 4096 else max(1, min(8, (1 << 16) // N))
        xt = feats(B, N, F).requires_grad_(True)
        sh = f"B={B} N={N} F={F} k={k}"
        run("knn_feature", sh, lambda: ops.knn_graph(xt.detach().transpose(1, 2), k).float())
        nbr = ops.NeighborIndex(ops.knn_graph(xt.detach().transpose(1, 2), k), N)
        run("csr_build", sh, lambda: (setattr(nbr, "_csr", None), nbr.csr()[1].float())[1])
        if N <= 16384:
            run("edge_feature(+bwd)", sh, lambda: ops.edge_features(xt, nbr), bwd=True)
        if F == 64:
            bn = torch.nn.BatchNorm2d(64).to(dev)
            PQ = feats(B, N, 128).requires_grad_(True)
            run("edgeconv_fused(+bwd)", sh + " O=64", lambda: ops.edgeconv_fused(PQ, nbr, bn, 0.2), bwd=True)
        del xt

    # ---- the 1x1-convolution GEMMs of both models (rows x Cin -> Cout): output, input-gradient and weight-gradient GEMM
    gemm_shapes = [(32 * 1024 * 32, 12, 32), (32 * 1024 * 32, 32, 32), (32 * 1024 * 32, 32, 64),          # PointNet++ SA1, B=32
                   (32 * 256 * 32, 68, 64), (32 * 256 * 32, 64, 128), (32 * 64 * 32, 132, 128), (32 * 64 * 32, 128, 256),
                   (32 * 16 * 32, 260, 256), (32 * 16 * 32, 256, 512), (32 * 1024, 320, 256), (32 * 4096, 128, 128),
                   (16 * 4096, 64, 128), (16 * 4096, 64, 256), (16 * 4096, 384, 1024), (16 * 4096, 1408, 512), (16 * 4096, 512, 256)]   # DGCNN, B=16
    for (R, Cin, Cout) in gemm_shapes if want("gemm") else []:
        x = torch.randn(R, Cin, generator=g).to(dev)
        w = (torch.randn(Cout, Cin, generator=g) / Cin ** 0.5).to(dev)
        gy = torch.randn(R, Cout, generator=g).to(dev)
        sh = f"R={R} Cin={Cin} Cout={Cout}"
        # as the layer functions of ops.py issue them: on the fp16-split path every operand's absmax is scanned once per
        # tensor and the weight is pre-split once per GEMM (absmax_kernel / split_f16_kernel rows)
        h2 = ops._gemm_h2_wanted(R, Cout, Cin)

        def fwd():
            ax, aw = (ops._absmax(x), ops._absmax(w)) if h2 else (None, None)
            return ops._gemm3x(x, False, w, False, R, Cout, Cin, amax_a=ax, amax_b=aw, b_split=ops._wsplit(w, False, aw))

        def dgrad():
            ag, aw = (ops._absmax(gy), ops._absmax(w)) if h2 else (None, None)
            return ops._gemm3x(gy, False, w, True, R, Cin, Cout, amax_a=ag, amax_b=aw, b_split=ops._wsplit(w, True, aw))

        def wgrad():
            ag, ax = (ops._absmax(gy), ops._absmax(x)) if h2 else (None, None)
            return ops._gemm3x(gy, True, x, True, Cout, Cin, R, amax_a=ag, amax_b=ax)
        run("conv1x1 y=xW^T", sh, fwd)
        run("conv1x1 dx=gyW", sh, dgrad)
        run("conv1x1 dW=gy^Tx", sh, wgrad)
        del x, w, gy

    # ---- the data formats either side of the path: packed block batches, sliding-window merge
    if want("io"):
        BD = pkg.block_datasets
        sizes = torch.randint(3000, 12000, (512,), generator=g).tolist()          # ~3.8 M points packed in HBM
        store = BD.PackedBlocks([(torch.randn(n, 9, generator=g), torch.zeros(n, 14, dtype=torch.uint8)) for n in sizes], dev)
        for B in (16, 32, 256):
            ids = torch.randint(0, len(store), (B,), generator=g).tolist()
            sel = store.draw_device(ids, 4096)
            run("block_batch", f"B={B} S=4096 L=14 ({len(store)} blocks, {store.nbytes() >> 20} MB packed)",
                lambda: store.batch(ids, 4096, sel)[0])
        for n_scene in (100_000, 1_000_000):
            wins = pkg.dgcnn_utils.scene_windows(n_scene, 4096, 512)
            total = sum(e - s for s, e in wins)
            logits = torch.randn(total, 13, generator=g).to(dev)
            offs = torch.tensor([0] + [e - s for s, e in wins[:-1]], dtype=torch.int64).cumsum(0).to(dev)
            pred = torch.empty(n_scene, dtype=torch.int64, device=dev)
            conf = torch.empty(n_scene, dtype=torch.float32, device=dev)

            def merge():
                lib.call("pcnbr_window_merge_f32", logits.data_ptr(), offs.data_ptr(), len(wins), n_scene, 4096, 3584, 13, None,
                         pred.data_ptr(), conf.data_ptr(), ops._stream())
                return conf
            run("window_merge", f"N={n_scene} window=4096 overlap=512 C=13", merge)

    md = ["| op | shape | kernel | µs/launch | algorithmic | roofline | frac |", "|---|---|---|---:|---:|---|---:|"]
    for r in rows:
        md.append(f"| {r['op']} | {r['shape']} | `{r['kernel']}` | {r['us_per_launch']} | {r['achieved']} {r['unit']} | {r['bound']} | {r['frac']} |")
    text = "\n".join(md)
    if args.md:
        with open(args.md, "w") as fh:
            fh.write(f"# kernel sweep ({torch.cuda.get_device_name(0)}; peaks: {peaks['source']}: HBM {peaks['hbm_gbs']} GB/s, TF32 {peaks['tf32_tflops']} TFLOP/s)\n\n")
            fh.write("Per-kernel CUDA-event timings from libpcnbr's profiler, L2 flushed between iterations; ALGORITHMIC bytes/flops per launch.\n\n")
            fh.write(text + "\n")
    print(text)


if __name__ == "__main__":
    main()
