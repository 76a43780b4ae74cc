#!/usr/bin/env python
"""bench.py -- train points/s of the neighbourhood hot path behind the reference's model interface.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--model dgcnn|pointnetpp|pointnetpp_msg|pointnext]
    python bench.py --impl reference ...        # the reference's CPU path (oracle port) on host cores
    python bench.py --impl reference --ref-device cuda ...   # the same op sequence through stock ATen kernels on cuda:0
    torchrun --nproc-per-node N bench.py --gpus N ...   (one rank per GPU, NCCL)

A step = one full train step (forward + backward + Adam, lr 1e-3 as the reference's train.py:17,79) of
the model on one batch of synthetic S3DIS-shaped blocks; every neighbourhood op runs in libpcnbr
(hand-written sm_100a kernels through the C ABI), the 1x1 convolutions / BatchNorm are library calls.
Default workload = BASELINE.json configs[1]: DGCNN (DGCNNWithColor, the class train.py builds) k=20,
batch 16 x 4096 points per GPU, 13 classes, strict fp32 (TF32 off).  Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

N_POINTS = 4096
N_CLASSES = 13
METRIC = "train_points_per_sec"
UNIT = "points/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="dgcnn", choices=["dgcnn", "pointnetpp", "pointnetpp_msg", "pointnext"])
    ap.add_argument("--batch", type=int, default=0, help="clouds per GPU (default 16 dgcnn / 32 pointnet++)")
    ap.add_argument("--points", type=int, default=N_POINTS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of the captured CUDA graph")
    ap.add_argument("--cpu-batch", type=int, default=2, help="clouds per CPU-baseline step (bounded sample)")
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference only: 'cuda' runs the same oracle port on cuda:0 with stock ATen / cuBLAS / cuDNN "
                         "kernels at the full per-GPU batch (SURVEY 8d 'also time': the reference's own GPU path)")
    return ap.parse_args()


def default_batch(model):
    return {"dgcnn": 16, "pointnetpp": 32, "pointnetpp_msg": 32, "pointnext": 4}[model]


def workload_name(model, B, N):
    return {
        "dgcnn": f"DGCNNWithColor semseg k=20 emb=1024, batch {B} x {N} pts x 6 ch per GPU, {N_CLASSES} classes, fwd+bwd+Adam",
        "pointnetpp": f"PointNet++ SSG semseg, batch {B} x {N} pts x 9 ch per GPU, {N_CLASSES} classes, fwd+bwd+Adam",
        "pointnetpp_msg": f"PointNet++ MSG semseg (two radii per level, one multi-radius ball query), batch {B} x {N} pts x 9 ch per GPU, {N_CLASSES} classes, fwd+bwd+Adam",
        "pointnext": f"PointNeXt semseg, batch {B} x {N} pts x 9 ch per GPU, {N_CLASSES} classes, fwd+bwd+Adam",
    }[model]


# --------------------------------------------------------------------------- clocks


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# --------------------------------------------------------------------------- data / models


def model_input(model, pts):
    # train.py feeds (B,N,9); DGCNNWithColor wants (B,6,N) (adapter outside the reference files)
    return pts[:, :, :6].transpose(1, 2) if model == "dgcnn" else pts


def build_model(impl_pkg, model):
    if model == "dgcnn":
        return impl_pkg.DGCNNWithColor(num_classes=N_CLASSES, k=20)
    if model == "pointnetpp":
        return impl_pkg.PointNetpp(N_CLASSES)
    if model == "pointnetpp_msg":
        return impl_pkg.PointNetppMSG(N_CLASSES)
    return impl_pkg.PointNeXt(N_CLASSES)


def logits_of(out):
    return out[0] if isinstance(out, tuple) else out


# --------------------------------------------------------------------------- reference arm / cpu baseline


def cpu_reference_steps(model, cloud_batch, N, steps, warmup, device="cpu"):
    """The reference's own implementation of the path (oracle/ref_ops.py: the same ATen op sequence, raw torch.topk
    selection): fwd + bwd + Adam on `cloud_batch` clouds, timed on the host cores -- or, with device="cuda", on cuda:0
    through the stock ATen / cuBLAS / cuDNN kernels the reference would launch there (torch defaults, as train.py)."""
    from oracle import ref_ops as O
    s3dis_blocks = O.s3dis_blocks
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    dev = torch.device(device)
    net = {"dgcnn": lambda: O.DGCNNWithColor(N_CLASSES, k=20, tie="raw"),
           "pointnetpp": lambda: O.PointNetpp(N_CLASSES, tie="raw"),
           "pointnetpp_msg": lambda: O.PointNetppMSG(N_CLASSES, tie="raw"),
           "pointnext": lambda: O.PointNeXt(N_CLASSES, tie="raw")}[model]().to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    pts, lab, lens = (t.to(dev) for t in s3dis_blocks(cloud_batch, N, 0, N_CLASSES))

    def ce(logits, onehot, lens):
        logp = torch.log_softmax(logits, dim=-1)
        tok = -(onehot.float() * logp).sum(-1)
        mask = (torch.arange(logits.shape[1], device=dev).unsqueeze(0) < lens.unsqueeze(1)).float()
        return (tok * mask).sum() / mask.sum()

    times = []
    for i in range(warmup + steps):
        if dev.type == "cuda":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss = ce(logits_of(net(model_input(model, pts))), lab, lens)
        loss.backward()
        opt.step()
        if dev.type == "cuda":
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    on_gpu = args.ref_device == "cuda"
    B = (args.batch or default_batch(args.model)) if on_gpu else args.cpu_batch
    steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    times, cores = cpu_reference_steps(args.model, B, args.points, steps, warmup, args.ref_device)
    ms = 1e3 * sum(times) / len(times)
    value = B * args.points / (ms / 1e3)
    sample = (f"{steps} steps of {B} clouds x {args.points} pts (" + ("the full per-GPU batch on cuda:0, stock ATen kernels"
              if on_gpu else "bounded sample of the per-GPU batch") + f"), {warmup} warm-up, mean")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic S3DIS-shaped blocks, random-init weights",
        "config": {"workload": workload_name(args.model, args.batch or default_batch(args.model), args.points),
                   "reference_path": "oracle port of the reference's torch " + ("CUDA path (stock ATen / cuBLAS / cuDNN kernels, torch defaults)"
                                     if on_gpu else "CPU path") + " (the Python reference does not travel to the GPU box)",
                   "ref_device": args.ref_device},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- our arm


FP32_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12                  # CUDA-core FMA peak of one B200 at the sustained SM clock
TENSOR_KERNELS = {"knn_tc_kernel", "gemm3x_kernel"}          # tcgen05 kernels: roofline = tensor pipe (TF32); everything else moves bytes


def ncu_traffic(kernel):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed `ncu --set full`
    capture of this kernel at the bench shapes (profiles/ncu_traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    ent = json.load(open(p)).get(kernel)
    return ent if ent else None


# kernels whose stated flops are CUDA-core work (lane-ops / flops on the FP32 pipe): the roofline that can bind them
# besides HBM is the FP32 issue rate.  FPS occupies one SM per cloud, so its ceiling is scaled by the SMs it uses.
ALU_KERNELS = {"fps_reg_kernel", "fps_big_kernel", "select_xyz_kernel<ball>", "select_xyz_kernel<knn>", "knn_expand_kernel"}


def kernel_bound(name, d, peaks):
    """Which roofline binds this kernel and how close it runs to it: time bound = max(algorithmic bytes / HBM peak,
    algorithmic flops / peak of the pipe that executes them); frac = bound / measured.  `d` = {"ms","bytes","flops","calls"}
    summed over the launches (ALGORITHMIC work as stated by the launch sites from the SURVEY.md 8d formulas)."""
    sec = d["ms"] / 1e3
    t_hbm = d["bytes"] / (peaks["hbm_gbs"] * 1e9)
    if name in TENSOR_KERNELS:
        t_fl, fl_unit, fl_peak, fl_name = d["flops"] / (peaks["tf32_tflops"] * 1e12), "TFLOP/s", peaks["tf32_tflops"], "tensor"
    elif name in ALU_KERNELS:
        t_fl, fl_unit, fl_peak, fl_name = d["flops"] / (peaks["fp32_tflops"] * 1e12), "TFLOP/s", peaks["fp32_tflops"], "alu"
    else:
        t_fl, fl_unit, fl_peak, fl_name = 0.0, "TFLOP/s", 1.0, "alu"
    if t_fl > t_hbm:
        return {"bound": fl_name, "achieved": d["flops"] / 1e12 / sec, "peak": fl_peak, "unit": fl_unit, "frac": t_fl / sec,
                "algorithmic": d["flops"]}
    return {"bound": "hbm", "achieved": d["bytes"] / 1e9 / sec, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": t_hbm / sec,
            "algorithmic": d["bytes"]}


def roofline_for(kernels, peaks):
    """Dominant libpcnbr KERNEL of the profiled steps against the roofline that bounds it.
    `achieved` = algorithmic bytes (or flops) of its launches, as stated by the launch sites from the SURVEY.md 8d
    formulas, / their summed duration (CUDA events on the launch stream, csrc/prof.cu)."""
    if not kernels:
        return None
    name, d = max(kernels.items(), key=lambda kv: kv[1]["ms"])
    b = kernel_bound(name, d, peaks)
    out = {"kernel": name, "bound": b["bound"], "achieved": b["achieved"], "peak": b["peak"], "unit": b["unit"], "frac": b["frac"],
           "traffic": (ncu_traffic(name) or {}).get("dram_bytes_per_launch"),
           "traffic_shape": (ncu_traffic(name) or {}).get("shape"),
           "traffic_algorithmic_bytes_same_shape": (ncu_traffic(name) or {}).get("algorithmic_bytes_same_shape"),
           "avg_launch_ms": d["ms"] / d["calls"], "calls": d["calls"], "peak_source": peaks["source"],
           "algorithmic_per_launch": b["algorithmic"] / d["calls"]}
    if name == "gemm3x_kernel" and b["bound"] == "tensor":
        # fp32-grade accuracy costs three TF32 instructions per algorithmic flop: the kernel's own ceiling is peak / 3
        out["issued_frac"] = 3.0 * b["frac"]
    out["note"] = ("flops counted once (2 M N K per GEMM, 2 N^2 F per kNN cloud), whatever the 3xTF32 kernel issues" if b["bound"] == "tensor" else
                   "algorithmic (compulsory) bytes; gathers that hit L2 are not counted" if b["bound"] == "hbm" else
                   "algorithmic lane-ops / flops on the FP32 pipe")
    return out


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf32_tflops": d["bf16_tflops_sustained"] / 2, "fp32_tflops": FP32_TFLOPS,
                "source": "MEASURED_PEAKS.json (tf32 = sustained bf16 / 2; fp32 = 148 SMs x 128 lanes x 2 x 1.965 GHz)"}
    return {"hbm_gbs": 6650.0, "tf32_tflops": 700.0, "fp32_tflops": FP32_TFLOPS, "source": "fallback (B200_PROFILING.md)"}


def run_ours(args):
    import torch.distributed as dist
    import __graft_entry__ as ge

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep NCCL's version / debug lines off stdout (one JSON line)
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = False       # strict fp32, as the north-star parity bar
    torch.backends.cudnn.allow_tf32 = False

    pkg = ge.load_package()
    pkg._lib.load()
    s3dis_blocks = pkg.synthetic.s3dis_blocks
    B = args.batch or default_batch(args.model)
    N = args.points
    torch.manual_seed(0)
    net = build_model(pkg, args.model).to(dev)
    pkg.train.broadcast_parameters(net)
    bucket = pkg.train.FlatGradBucket(net, steal_grads=True)
    # torch's fused Adam: one multi-tensor kernel per step instead of ~8 (same update rule as the reference's Adam(lr=1e-3))
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, capturable=not args.no_graph, fused=True)

    n_batches = 4                                        # rotate inputs; activations (GBs) >> L2 anyway
    host, devb = [], []
    for i in range(n_batches):
        pts, lab, lens = s3dis_blocks(B, N, seed=1000 * rank + i, classes=N_CLASSES)
        host.append((pts.pin_memory(), lab.pin_memory(), lens.pin_memory()))
        devb.append((pts.to(dev), lab.to(dev), lens.to(dev)))

    def loss_of(model, pts, lab, lens):
        return pkg.train.masked_onehot_cross_entropy(logits_of(model(model_input(args.model, pts))), lab, lens)

    def eager_step(pts, lab, lens):
        bucket.zero()
        loss = loss_of(net, pts, lab, lens)
        loss.backward()
        pkg.ops.join_aux()
        bucket.all_reduce_mean()
        opt.step()
        return loss

    for i in range(args.warmup):
        eager_step(*devb[i % n_batches])
    step = eager_step if args.no_graph else pkg.train.GraphedTrainStep(net, opt, bucket, loss_of, devb[0], warmup=2)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(region_steps, from_host):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        last = None
        for i in range(region_steps):
            if from_host:
                hp, hl, hn = host[i % n_batches]         # pinned host batch -> H2D copies inside the timed region
                if args.no_graph:
                    hp, hl, hn = hp.to(dev, non_blocking=True), hl.to(dev, non_blocking=True), hn.to(dev, non_blocking=True)
                last = step(hp, hl, hn).item()           # (graph: copied straight into the static inputs); D2H loss read
            else:
                last = step(*devb[i % n_batches])
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, last

    for i in range(args.warmup):
        step(*devb[i % n_batches])
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_total, _ = timed(args.steps, from_host=False)
    ms_e2e, last_loss = timed(args.steps, from_host=True)
    clk = clocks.stop() if rank == 0 else None
    # per-kernel durations: the same kernels launched eagerly with CUDA events around every libpcnbr call
    # (events cannot be recorded inside a graph replay); also counts the libpcnbr launches of one step
    prof_steps = 3
    pkg._lib.prof_enable(True)
    for i in range(prof_steps):
        eager_step(*devb[i % n_batches])
    kernels = pkg._lib.prof_collect()
    pkg._lib.prof_enable(False)
    launches = sum(d["calls"] for d in kernels.values()) // prof_steps * args.steps

    pts_per_step = B * N * world
    value = pts_per_step * args.steps / (ms_total / 1e3)
    e2e = pts_per_step * args.steps / (ms_e2e / 1e3)
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    peaks = load_peaks()

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        times, cores = cpu_reference_steps(args.model, args.cpu_batch, N, 2, 1)
        cms = sum(times) / len(times)
        cpu_base = {"value": args.cpu_batch * N / cms, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": f"2 steps of {args.cpu_batch} clouds x {N} pts (oracle port of the reference's torch CPU path), 1 warm-up, mean"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic S3DIS-shaped blocks (SURVEY.md 8d generator), random-init weights",
            "config": {"workload": workload_name(args.model, B, N), "global_batch": B * world, "parallelism": f"dp{world}",
                       "precision": "strict fp32 (TF32 disabled for cuBLAS and cuDNN)",
                       "l2": "no explicit flush: per-step activations (GBs) far exceed the 126 MB L2; 4 input batches rotate"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps,
                    "loss": last_loss},
            "gpu_launches": launches,
            "clocks": clk,
            "roofline": roofline_for(kernels, peaks),
            "cpu_baseline": cpu_base,
            "kernel_ms_per_step": {k: round(v["ms"] / prof_steps, 4) for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]["ms"])},
            "kernel_roofline_frac": {k: [kernel_bound(k, v, peaks)["bound"], round(kernel_bound(k, v, peaks)["frac"], 4)]
                                     for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]["ms"])},
            "launch_mode": "eager" if args.no_graph else "whole train step captured in one CUDA graph, replayed per batch",
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # a captured graph holds NCCL work: tear down in order (graph, then a device sync, then the group) and
        # leave without the interpreter's atexit pass, which can block on the communicator
        del step
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
