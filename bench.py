#!/usr/bin/env python
"""bench.py -- train points/s of the neighbourhood hot path behind the reference's model interface.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--model dgcnn|pointnetpp|pointnetpp_msg|pointnext]
                    [--scaling weak|strong] [--no-extras]
    python bench.py --impl reference ...                     # the UNMODIFIED reference (oracle/_ref) on the host cores
    python bench.py --impl reference --ref-device cuda ...   # the same reference code through stock ATen kernels on cuda:0
    torchrun --nproc-per-node N bench.py --gpus N ...        # one rank per GPU, NCCL

A step = one full train step (forward + backward + Adam, lr 1e-3 as the reference's train.py:17,79) of the model on one
batch of synthetic S3DIS-shaped blocks; every op of the step runs in libpcnbr (hand-written sm_100a kernels through the
C ABI) except Adam (torch's fused multi-tensor kernel) and the NCCL all-reduce.  Default workload = BASELINE.json
configs[1]: DGCNN (DGCNNWithColor, the class train.py builds) k=20, batch 16 x 4096 points per GPU, 13 classes, strict
fp32.  Prints ONE JSON line.  With the default model the line also carries (unless --no-extras)
  "pointnetpp"     the co-headline model of the north star (PointNet++ SSG, 32 x 4096 per GPU), same measurement;
  "reference_gpu"  SURVEY 8d "also time": the unmodified reference's train step through stock ATen / cuBLAS / cuDNN on
                   the same B200, same batch (N=1 only);
  "strong"         N>1 only: the same models with the GLOBAL batch fixed at the BASELINE batch (16 / 32 clouds split over
                   the ranks), i.e. the strong-scaling point next to the weak one.
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

def _argv_value(flag, default):
    for i, a in enumerate(sys.argv):
        if a == flag and i + 1 < len(sys.argv):
            return sys.argv[i + 1]
        if a.startswith(flag + "="):
            return a.split("=", 1)[1]
    return default


# the reference's CPU path must not see a GPU: models/dgcnn/dgcnn.py:39 picks its device from torch.cuda.is_available(),
# not from its input, so on a CUDA box the unmodified DGCNN cannot run on CPU tensors unless CUDA is hidden
if _argv_value("--impl", "ours") == "reference" and _argv_value("--ref-device", "cpu") == "cpu":
    os.environ["CUDA_VISIBLE_DEVICES"] = ""

import torch  # noqa: E402

N_POINTS = 4096
N_CLASSES = 13
METRIC = "train_points_per_sec"
UNIT = "points/s"
REF_TREE = os.path.join(ROOT, "oracle", "_ref")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="dgcnn", choices=["dgcnn", "pointnetpp", "pointnetpp_msg", "pointnext"])
    ap.add_argument("--batch", type=int, default=0, help="clouds per GPU (default 16 dgcnn / 32 pointnet++ / 4 pointnext)")
    ap.add_argument("--points", type=int, default=N_POINTS)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch clouds per GPU; strong: the BASELINE batch split over the ranks")
    ap.add_argument("--no-extras", action="store_true", help="only the requested model (no pointnetpp / reference_gpu / strong records)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of the captured CUDA graph")
    ap.add_argument("--prefetch", action="store_true",
                    help="software-pipeline the step: the NEXT batch's geometry (FPS / ball query / kNN tables + CSR inverses) is computed on a "
                         "side stream during the current step (train.GraphedTrainStep(geometry_fn)); measured neutral on B200 (DESIGN.md 7), off by default")
    ap.add_argument("--cpu-batch", type=int, default=0, help="clouds per CPU-reference step (0 = sized to ~120 s of CPU work)")
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference only: 'cuda' runs the same reference code on cuda:0 with stock ATen / cuBLAS / cuDNN "
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


def config_for(model, B, N, world):
    """The `config` object: identical for our arm and the reference arm of the same command line."""
    return {"workload": workload_name(model, B, N), "global_batch": B * world, "parallelism": f"dp{world}",
            "precision": "strict fp32 (TF32 disabled for cuBLAS and cuDNN)",
            "l2": "no explicit flush: per-step activations (GBs) far exceed the 126 MB L2; 4 input batches rotate"}


# --------------------------------------------------------------------------- clocks


class ClockSampler:
    """SM clock / throttle reasons WHILE the timed region runs.  In-process NVML (the library nvidia-smi reads: nvml
    DeviceGetClockInfo / CurrentClocksEventReasons, one sample every 20 ms with its time stamp), so that each timed region of
    50 ms is covered by 2-3 samples; `nvidia-smi -lms` as a child process (the fallback when pynvml cannot be loaded)
    needs ~100 ms to start and mostly sampled the idle GPU behind the region.  mark(t0, t1): the wall-clock windows of the
    timed regions -- only samples inside them count as "under load"."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.windows = index, [], None, []
        self.nvml, self.handle, self._stop, self.thread = None, None, threading.Event(), None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            pr = torch.cuda.get_device_properties(self.index)      # CUDA_VISIBLE_DEVICES may renumber: go by PCI address
            h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0".encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        return pynvml, h

    def start(self):
        try:
            self.nvml, self.handle = self._nvml_handle()
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return self
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None
        return self

    def _poll(self):
        n = self.nvml
        bits = [n.nvmlClocksEventReasonHwSlowdown, n.nvmlClocksEventReasonHwThermalSlowdown,
                n.nvmlClocksEventReasonSwThermalSlowdown, n.nvmlClocksEventReasonSwPowerCap]
        try:
            mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        except Exception:
            mx = None
        while not self._stop.is_set():
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                rs = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                self.rows.append((time.time(), sm, mx, [nm for nm, b in zip(self.NAMES, bits) if rs & b]))
            except Exception:
                pass
            # every 20 ms: 2-3 samples per 50 ms timed region.  (Every 4 ms -- 75 NVML calls under a 100 ms measurement -- two of
            # fourteen end-to-end regions showed a 3-5 ms host stall: NVML queries share driver locks with the CUDA calls of
            # the thread that launches and synchronises every step.)
            time.sleep(0.02)

    def _read(self):
        for line in self.proc.stdout:
            c = [x.strip() for x in line.split(",")]
            if len(c) >= 6 and c[0].isdigit() and c[1].isdigit():
                self.rows.append((time.time(), int(c[0]), int(c[1]), [nm for nm, v in zip(self.NAMES, c[2:6]) if v.lower().startswith("active")]))

    def mark(self, t0: float, t1: float):
        self.windows.append((t0, t1))

    def stop(self):
        if self.nvml is None and self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if self.nvml is not None:
            self._stop.set()
            self.thread.join(timeout=1.0)
        else:
            time.sleep(0.05)
            self.proc.terminate()
        rows = list(self.rows)
        inside = [r for r in rows if any(t0 <= r[0] <= t1 for t0, t1 in self.windows)] if self.windows else rows
        use = inside if len(inside) >= 2 else rows                 # (a region shorter than two sampling periods: every sample of the run)
        sm = sorted(r[1] for r in use)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max((r[2] for r in use if r[2] is not None), default=None),
                "reasons": sorted({nm for r in use for nm in r[3]}), "samples": len(use),
                "samples_in_timed_region": len(inside), "source": "nvml" if self.nvml is not None else "nvidia-smi -lms 20"}


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


# --------------------------------------------------------------------------- the reference (oracle/_ref, else the oracle port)


def reference_factory(model):
    """-> (constructor, criterion, kind).  kind "reference": the UNMODIFIED reference classes from oracle/_ref (a verbatim
    copy made by oracle/make_ref.py; /root/reference itself does not exist on the GPU box) with the reference's own
    criterion (Training/train_model.py:15-57).  kind "port": oracle/ref_ops.py (same ATen op sequence, raw torch.topk
    selection) -- only when the copy is missing, or for the MSG network, which the reference does not ship."""
    have = os.path.isdir(os.path.join(REF_TREE, "models", "utils"))
    if have and model in ("dgcnn", "pointnetpp", "pointnext"):
        if REF_TREE not in sys.path:
            sys.path.insert(0, REF_TREE)
        from Training.train_model import masked_onehot_cross_entropy as crit
        if model == "dgcnn":
            from models.dgcnn.dgcnn import DGCNNWithColor
            return (lambda: DGCNNWithColor(num_classes=N_CLASSES, k=20)), crit, "reference"
        if model == "pointnetpp":
            from models.PointNetpp.PointNetpp import PointNetpp
            return (lambda: PointNetpp(N_CLASSES)), crit, "reference"
        from models.PointNeXt.PointNeXt import PointNeXt
        return (lambda: PointNeXt(N_CLASSES)), crit, "reference"
    from oracle import ref_ops as O

    def ce(logits, onehot, lens):
        logp = torch.log_softmax(logits, dim=-1)
        tok = -(onehot.float() * logp).sum(-1)
        mask = (torch.arange(logits.shape[1], device=logits.device).unsqueeze(0) < lens.to(logits.device).unsqueeze(1)).float()
        return (tok * mask).sum() / mask.sum()
    ctor = {"dgcnn": lambda: O.DGCNNWithColor(N_CLASSES, k=20, tie="raw"), "pointnetpp": lambda: O.PointNetpp(N_CLASSES, tie="raw"),
            "pointnetpp_msg": lambda: O.PointNetppMSG(N_CLASSES, tie="raw"), "pointnext": lambda: O.PointNeXt(N_CLASSES, tie="raw")}[model]
    return ctor, ce, "port"


def reference_steps(model, cloud_batch, N, steps, warmup, device="cpu"):
    """fwd + bwd + Adam of the reference on `cloud_batch` clouds, wall-clock per step (device synchronised on CUDA).
    -> (times [s], threads, kind)."""
    from oracle.ref_ops import s3dis_blocks
    ctor, crit, kind = reference_factory(model)
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    dev = torch.device(device)
    net = ctor().to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    pts, lab, lens = s3dis_blocks(cloud_batch, N, 0, N_CLASSES)
    pts, lab = pts.to(dev), lab.to(dev)
    times = []
    for i in range(warmup + steps):
        if dev.type == "cuda":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss = crit(logits_of(net(model_input(model, pts))), lab, lens)
        loss.backward()
        opt.step()
        if dev.type == "cuda":
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times, torch.get_num_threads(), kind


def cpu_sample_batch(model, full_batch, N, steps, warmup, budget_s=75.0):
    """Clouds per CPU step so that warmup + steps fit `budget_s` of host time: one calibration step on 2 clouds (points/s is
    batch-independent on the CPU within ~10 %: the ops are per cloud), then the largest batch <= the full per-GPU batch."""
    t, _, _ = reference_steps(model, 2, N, 1, 1)
    per_cloud = t[0] / 2
    fit = int(budget_s / max(1, steps + warmup) / max(per_cloud, 1e-6))
    return max(2, min(full_batch, fit)), per_cloud


def cpu_baseline_subprocess(model, clouds, N):
    """The `cpu_baseline` leg of our arm: the reference arm of this script on a bounded sample, in a child process that
    cannot see the GPU (the unmodified reference picks 'cuda' whenever it is available, dgcnn.py:39)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--model", model, "--cpu-batch", str(clouds),
           "--points", str(N), "--steps", "2", "--warmup", "1"]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
        line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
        return line["cpu_baseline"]
    except Exception as e:
        return {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {e!r}"[:300]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    on_gpu = args.ref_device == "cuda"
    full = args.batch or default_batch(args.model)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    steps, warmup = max(1, args.steps), max(1, args.warmup)
    if on_gpu:
        B = full
        steps, warmup = min(steps, 5), min(warmup, 2)
    elif args.cpu_batch:
        B = args.cpu_batch
    else:
        B, _ = cpu_sample_batch(args.model, full, args.points, steps, warmup)
    clocks = ClockSampler(0).start() if on_gpu else None
    times, cores, kind = reference_steps(args.model, B, args.points, steps, warmup, args.ref_device)
    clk = clocks.stop() if clocks else None
    ms = 1e3 * sum(times) / len(times)
    value = B * args.points / (ms / 1e3)
    what = ("the unmodified reference (oracle/_ref: verbatim copy of /root/reference's models/ + Training/)" if kind == "reference"
            else "oracle port of the reference (oracle/ref_ops.py)")
    sample = (f"{steps} steps of {B} clouds x {args.points} pts" + (" = the full per-GPU batch" if B == full else f" (bounded sample of the {full}-cloud batch)")
              + f", {warmup} warm-up, mean; {what}; " + ("cuda:0, stock ATen / cuBLAS / cuDNN kernels, torch defaults" if on_gpu else f"{cores} host threads"))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic S3DIS-shaped blocks (SURVEY.md 8d generator), random-init weights",
        "config": config_for(args.model, full, args.points, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "ref_device": args.ref_device,
    }
    if clk:
        line["clocks"] = clk
    emit_line(line)


# --------------------------------------------------------------------------- roofline bookkeeping


FP32_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12                  # CUDA-core FMA peak of one B200 at the sustained SM clock
# tcgen05 kernels whose binding roofline can be the tensor pipe; gemm kernels are profiled per shape class ("[tensor]" = the
# launch's 2MNK / tensor peak exceeds its bytes / HBM peak, "[hbm]" otherwise), so one model's wide layers and another
# model's narrow layers are never averaged into one figure
TENSOR_KERNELS = {"knn_tc_kernel", "gemm3x_kernel[tensor]", "gemm3x_kernel[hbm]", "gemm2h_kernel[tensor]", "gemm2h_kernel[hbm]"}
# kernels whose stated flops are CUDA-core work (lane-ops / flops on the FP32 pipe): the roofline that can bind them
# besides HBM is the FP32 issue rate.  FPS occupies one SM (or one cluster) per cloud: its stated work is already scaled
# to the SMs it uses.
ALU_KERNELS = {"fps_reg_kernel", "fps_big_kernel", "fps_cluster_kernel", "select_xyz_kernel<ball>", "select_xyz_kernel<knn>",
               "knn_expand_kernel", "ball_grid_kernel", "knn3_grid_kernel"}


def ncu_traffic(kernel):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed `ncu --set full`
    capture of this kernel at the bench shapes (profiles/ncu_traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    table = json.load(open(p))
    return table.get(kernel) or table.get(kernel.split("[")[0]) or None


# tensor instructions ISSUED per algorithmic flop: the fp32-accurate GEMMs run three products per product (hi.hi' + lo.hi' +
# hi.lo'); the kNN kernel runs two passes over the tiles plus a 16-wide tail K-slice per 64-wide feature slab
ISSUE_FACTOR = {"gemm3x_kernel": 3.0, "gemm2h_kernel": 3.0, "knn_tc_kernel": 2.5}


def kernel_bound(name, d, peaks):
    """Which roofline binds this kernel and how close it runs to it.  `d` = {"ms","bytes","flops","calls"} summed over the
    launches (ALGORITHMIC work as stated by the launch sites from the SURVEY.md 8d formulas).  A tensor-core kernel is
    tensor-bound when the tensor time of the instructions it ISSUES (ISSUE_FACTOR x algorithmic flops / peak of the pipe it
    uses: f16 for the fp16-split GEMM and the kNN kernel, tf32 for the 3xTF32 GEMM) exceeds its HBM time; its `achieved` /
    `frac` still count every flop ONCE (so a three-product GEMM cannot exceed 1/3), `issued_frac` says how busy the pipe is.
    Everything else: max(bytes / HBM peak, flops / FP32 peak) / measured."""
    sec = d["ms"] / 1e3
    t_hbm = d["bytes"] / (peaks["hbm_gbs"] * 1e9)
    base = name.split("[")[0]
    if name in TENSOR_KERNELS:
        pk = peaks["tf32_tflops"] if base == "gemm3x_kernel" else peaks["f16_tflops"]
        t_once = d["flops"] / (pk * 1e12)
        t_issued = ISSUE_FACTOR.get(base, 1.0) * t_once
        if t_issued > t_hbm:
            return {"bound": "tensor", "achieved": d["flops"] / 1e12 / sec, "peak": pk, "unit": "TFLOP/s", "frac": t_once / sec,
                    "algorithmic": d["flops"], "issued_frac": t_issued / sec, "hbm_frac": t_hbm / sec}
        return {"bound": "hbm", "achieved": d["bytes"] / 1e9 / sec, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": t_hbm / sec,
                "algorithmic": d["bytes"], "issued_frac": t_issued / sec}
    if name in ALU_KERNELS:
        t_fl = d["flops"] / (peaks["fp32_tflops"] * 1e12)
        if t_fl > t_hbm:
            return {"bound": "alu", "achieved": d["flops"] / 1e12 / sec, "peak": peaks["fp32_tflops"], "unit": "TFLOP/s",
                    "frac": t_fl / sec, "algorithmic": d["flops"]}
    return {"bound": "hbm", "achieved": d["bytes"] / 1e9 / sec, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": t_hbm / sec,
            "algorithmic": d["bytes"]}


def roofline_for(kernels, peaks):
    """Dominant libpcnbr KERNEL (per shape class) of the profiled steps against the roofline that bounds it.
    `achieved` = algorithmic bytes (or flops) of its launches, as stated by the launch sites from the SURVEY.md 8d
    formulas, / their summed duration (CUDA events on the launch stream, csrc/prof.cu)."""
    if not kernels:
        return None
    name, d = max(kernels.items(), key=lambda kv: kv[1]["ms"])
    b = kernel_bound(name, d, peaks)
    t = ncu_traffic(name) or {}
    out = {"kernel": name, "bound": b["bound"], "achieved": b["achieved"], "peak": b["peak"], "unit": b["unit"], "frac": b["frac"],
           "traffic": t.get("dram_bytes_per_launch"), "traffic_shape": t.get("shape"),
           "traffic_algorithmic_bytes_same_shape": t.get("algorithmic_bytes_same_shape"),
           "avg_launch_ms": d["ms"] / d["calls"], "calls": d["calls"], "peak_source": peaks["source"],
           "algorithmic_per_launch": b["algorithmic"] / d["calls"]}
    if "issued_frac" in b:
        out["issued_per_algorithmic_flop"] = ISSUE_FACTOR.get(name.split("[")[0], 1.0)
        out["issued_frac"] = b["issued_frac"]              # share of the tensor pipe's time the issued instructions need
    if "hbm_frac" in b:
        out["hbm_frac"] = b["hbm_frac"]                    # the same launches against the HBM roofline (algorithmic bytes)
    out["note"] = ("flops counted once (2 M N K per GEMM, 2 N^2 F per kNN cloud), whatever the split kernel issues" if b["bound"] == "tensor" else
                   "algorithmic (compulsory) bytes; gathers that hit L2 are not counted" if b["bound"] == "hbm" else
                   "algorithmic lane-ops / flops on the FP32 pipe")
    return out


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf32_tflops": d["bf16_tflops_sustained"] / 2, "f16_tflops": d["bf16_tflops_sustained"],
                "fp32_tflops": FP32_TFLOPS,
                "source": "MEASURED_PEAKS.json (f16 = sustained bf16; tf32 = sustained bf16 / 2; fp32 = 148 SMs x 128 lanes x 2 x 1.965 GHz)"}
    return {"hbm_gbs": 6650.0, "tf32_tflops": 700.0, "f16_tflops": 1400.0, "fp32_tflops": FP32_TFLOPS, "source": "fallback (B200_PROFILING.md)"}


# --------------------------------------------------------------------------- our arm


class Harness:
    """Process-wide state of our arm: device, ranks, package."""

    def __init__(self, args):
        import torch.distributed as dist
        import __graft_entry__ as ge
        self.args, self.dist = args, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep NCCL's version / debug lines off stdout (one JSON line)
            dist.init_process_group("nccl", device_id=self.dev)
        torch.backends.cuda.matmul.allow_tf32 = False       # strict fp32, as the north-star parity bar
        torch.backends.cudnn.allow_tf32 = False
        self.pkg = ge.load_package()
        self.pkg._lib.load()
        self.peaks = load_peaks()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def measure(self, model, B, N, steps, warmup, profile=True, cpu_baseline=False):
        """Train-step throughput of `model` at B clouds per GPU: device-resident (`value`) and end to end from pinned host
        batches (`e2e`), clocks sampled over both timed regions, then the per-kernel map from eager profiled steps."""
        args, pkg, dev, world, rank = self.args, self.pkg, self.dev, self.world, self.rank
        torch.manual_seed(0)
        net = build_model(pkg, model).to(dev)
        pkg.train.broadcast_parameters(net)
        bucket = pkg.train.FlatGradBucket(net, steal_grads=True)
        # torch's fused Adam: one multi-tensor kernel per step (same update rule as the reference's Adam(lr=1e-3))
        opt = torch.optim.Adam(net.parameters(), lr=1e-3, capturable=not args.no_graph, fused=True)
        n_batches = 4                                        # rotate inputs; activations (GBs) >> L2 anyway
        host, devb = [], []
        for i in range(n_batches):
            pts, lab, lens = pkg.synthetic.s3dis_blocks(B, N, seed=1000 * rank + i, classes=N_CLASSES)
            host.append((pts.pin_memory(), lab.pin_memory(), lens.pin_memory()))
            devb.append((pts.to(dev), lab.to(dev), lens.to(dev)))

        def loss_of(m, pts, lab, lens, geometry=None):
            out = m(model_input(model, pts), geometry=geometry) if geometry is not None else m(model_input(model, pts))
            return pkg.train.masked_onehot_cross_entropy(logits_of(out), lab, lens)

        # index-only work of the NEXT batch (FPS picks, ball-query / first-layer kNN tables, their CSR inverses) runs on a side
        # stream during the current step: the captured step is software-pipelined over two batches (train.GraphedTrainStep)
        geo_fn = None
        if model in ("dgcnn", "pointnetpp") and args.prefetch and not args.no_graph:
            geo_fn = lambda m, pts, lab, lens, stream=None: m.prepare_geometry(model_input(model, pts), stream=stream)
            if os.environ.get("PCNBR_FAKE_PREFETCH"):            # diagnostic only: geometry computed once, never again (an INVALID
                cache = {}                                         # measurement: shows the step time with no geometry work at all)

                def geo_fn(m, pts, lab, lens, stream=None, _real=geo_fn):
                    if "g" not in cache:
                        cache["g"] = _real(m, pts, lab, lens, stream=stream)
                    elif stream is not None:
                        stream.wait_event(torch.cuda.current_stream().record_event())     # keep the fork / join structure
                    return cache["g"]

        def eager_step(pts, lab, lens):
            bucket.zero()
            loss = loss_of(net, pts, lab, lens)
            loss.backward()
            pkg.ops.join_aux()
            bucket.all_reduce_mean()
            opt.step()
            return loss

        pkg.ops.reset_fallbacks()
        for i in range(max(warmup, 3)):
            eager_step(*devb[i % n_batches])
        step = eager_step if args.no_graph else pkg.train.GraphedTrainStep(net, opt, bucket, loss_of, devb[0], warmup=2, geometry_fn=geo_fn)

        clocks = ClockSampler(self.local).start() if rank == 0 else None     # sampling runs from before the warm-up on

        def timed(region_steps, from_host, mark=True):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.barrier()
            t_wall0 = time.time()
            ev0.record()
            last = None
            for i in range(region_steps):
                if from_host:
                    hp, hl, hn = host[i % n_batches]         # pinned host batch -> H2D copies inside the timed region
                    if args.no_graph:
                        hp, hl, hn = hp.to(dev, non_blocking=True), hl.to(dev, non_blocking=True), hn.to(dev, non_blocking=True)
                    loss = step(hp, hl, hn)                  # (graph: into the static inputs -- the staged copy when prefetched)
                    if not args.no_graph and not os.environ.get("PCNBR_BENCH_NO_PREFETCH"):   # next batch's H2D copy on the copy stream, under this step
                        step.prefetch(*host[(i + 1) % n_batches])
                    last = loss.item()                       # D2H loss read
                else:
                    last = step(*devb[i % n_batches])
            ev1.record()
            self.barrier()
            if clocks and mark:
                clocks.mark(t_wall0, time.time())            # the samples inside this window are the ones "under load"
            ms = ev0.elapsed_time(ev1)
            if world > 1:
                t = torch.tensor([ms], device=dev)
                self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
                ms = t.item()
            return ms, last

        for i in range(max(warmup, 3)):
            step(*devb[i % n_batches])
        ms_total, _ = timed(steps, from_host=False)
        # the end-to-end loop has host code of its own (pinned-memory copies, the prefetch, .item()): W untimed steps of exactly
        # that loop first -- on a fresh box its first executions page in library code and cost milliseconds (tools/e2e_jitter.py:
        # the first seconds of a host-driven loop run 5.7-6.0 ms per step with 7-8 ms outliers, 5.15-5.2 after that)
        timed(max(warmup, 3), from_host=True, mark=False)
        ms_e2e, last_loss = timed(steps, from_host=True)
        clk = clocks.stop() if clocks else None
        out = {"ms_per_step": ms_total / steps, "value": B * N * world * steps / (ms_total / 1e3),
               "e2e": {"value": B * N * world * steps / (ms_e2e / 1e3), "unit": UNIT,
                       "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in host[0]), "d2h_bytes_per_step": 4,
                       "h2d": ("eager launches: copied at the head of the step" if args.no_graph else
                               "every step's pinned host batch is copied inside the timed region; the NEXT step's copy is issued on a "
                               "copy stream under the running step (GraphedTrainStep.prefetch) and handed over device-side"),
                       "ms_per_step": ms_e2e / steps, "loss": last_loss},
               "clocks": clk, "library_fallbacks": pkg.ops.fallbacks(),
               "geometry": ("next batch's FPS / ball-query / kNN tables + CSR inverses computed on a side stream during the current step "
                            "(one H2D batch copy and one full train step per timed step; a call returns the previous batch's loss)"
                            if geo_fn is not None else "computed inside the step")}
        if profile:
            # per-kernel durations: the same kernels launched eagerly with CUDA events around every libpcnbr kernel
            # (events cannot be recorded inside a graph replay); also counts the libpcnbr launches of one step
            prof_steps = 3
            pkg._lib.prof_enable(True)
            for i in range(prof_steps):
                eager_step(*devb[i % n_batches])
            kernels = pkg._lib.prof_collect()
            pkg._lib.prof_enable(False)
            out["gpu_launches"] = sum(d["calls"] for d in kernels.values()) // prof_steps * steps
            out["roofline"] = roofline_for(kernels, self.peaks)
            order = sorted(kernels.items(), key=lambda kv: -kv[1]["ms"])
            out["kernel_ms_per_step"] = {k: round(v["ms"] / prof_steps, 4) for k, v in order}
            out["kernel_roofline_frac"] = {k: [kernel_bound(k, v, self.peaks)["bound"], round(kernel_bound(k, v, self.peaks)["frac"], 4)] for k, v in order}
        if cpu_baseline and rank == 0 and world == 1:
            out["cpu_baseline"] = cpu_baseline_subprocess(model, args.cpu_batch or 2, N)
        if world > 1 and not args.no_graph:
            del step                                         # a captured graph holds NCCL work: drop it before the next capture
        del net, opt, bucket
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        return out

    def reference_on_gpu(self, model, B, N):
        """SURVEY 8d 'also time': the unmodified reference's own train step on this B200 (stock ATen / cuBLAS / cuDNN
        kernels, torch defaults as train.py), full per-GPU batch, wall clock around synchronised steps."""
        tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = False, True      # torch's defaults
        try:
            clocks = ClockSampler(self.local).start()
            times, _, kind = reference_steps(model, B, N, 3, 2, device=str(self.dev))
            clk = clocks.stop()
        finally:
            torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
        torch.cuda.empty_cache()
        ms = 1e3 * sum(times) / len(times)
        return {"ms_per_step": ms, "value": B * N / (ms / 1e3), "unit": UNIT, "kind": kind, "steps": 3, "warmup": 2,
                "batch": B, "clocks": clk, "timing": "wall clock around device-synchronised steps (the reference syncs per FPS pick itself)"}


def run_ours(args):
    h = Harness(args)
    world, rank = h.world, h.rank
    N = args.points
    full = args.batch or default_batch(args.model)
    strong = args.scaling == "strong"
    B = max(1, full // world) if strong else full
    extras = (not args.no_extras) and args.model == "dgcnn" and not args.batch and N == N_POINTS and not strong
    main = h.measure(args.model, B, N, args.steps, args.warmup, profile=True, cpu_baseline=not args.no_cpu_baseline)
    line = {
        "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic S3DIS-shaped blocks (SURVEY.md 8d generator), random-init weights",
        "config": config_for(args.model, B, N, world),
        "e2e": main["e2e"], "gpu_launches": main.get("gpu_launches"), "clocks": main["clocks"], "roofline": main.get("roofline"),
        "cpu_baseline": main.get("cpu_baseline"), "kernel_ms_per_step": main.get("kernel_ms_per_step"),
        "kernel_roofline_frac": main.get("kernel_roofline_frac"), "library_fallbacks": main["library_fallbacks"],
        "geometry": main["geometry"],
        "launch_mode": "eager" if args.no_graph else "whole train step captured in one CUDA graph, replayed per batch",
    }
    if extras:
        pb = default_batch("pointnetpp")
        pn = h.measure("pointnetpp", pb, N, args.steps, args.warmup, profile=True)
        pn["config"] = config_for("pointnetpp", pb, N, world)
        pn["unit"] = UNIT
        line["pointnetpp"] = pn
        if world == 1:
            try:
                line["reference_gpu"] = {"dgcnn": h.reference_on_gpu("dgcnn", full, N), "pointnetpp": h.reference_on_gpu("pointnetpp", pb, N)}
            except Exception as e:                           # never lose the headline line to the side measurement
                line["reference_gpu"] = {"error": repr(e)[:300]}
        else:
            st = {}
            for m in ("dgcnn", "pointnetpp"):
                g = default_batch(m)
                if g % world:
                    continue
                r = h.measure(m, g // world, N, args.steps, args.warmup, profile=False)
                st[m] = {"global_batch": g, "clouds_per_gpu": g // world, "ms_per_step": r["ms_per_step"], "value": r["value"],
                         "e2e": r["e2e"], "unit": UNIT, "clocks": r["clocks"]}
            line["strong"] = st
    if rank == 0:
        emit_line(line)
    if world > 1:
        # captured graphs held NCCL work: tear down in order (device sync, then the group) and leave without the
        # interpreter's atexit pass, which can block on the communicator
        torch.cuda.synchronize()
        h.dist.barrier()
        sys.stdout.flush()
        os._exit(0)


class _QuietStdout:
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on file descriptor
    1 whatever NCCL_DEBUG_FILE says on some builds; subprocesses inherit it), so the descriptor itself is pointed at stderr
    for the whole run and only emit() writes to the real stdout."""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, text: str) -> None:
        sys.stdout.flush()
        os.write(self.real, (text.rstrip("\n") + "\n").encode())


OUT = None


def emit_line(line: dict) -> None:
    text = json.dumps(line)
    if OUT is not None:
        OUT.emit(text)
    else:
        print(text, flush=True)


def main():
    global OUT
    args = parse()
    OUT = _QuietStdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
