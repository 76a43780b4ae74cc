"""Which host-side helper disturbs the end-to-end region?  The DGCNN step (captured graph) run as bench.py's e2e loop (pinned
host batch, prefetch of the next one, .item() per step), 40 regions of 10 steps per condition: NVML polling off / every 4 ms /
every 20 ms, prefetch on / off.  Prints median and worst ms per step of each condition."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
pkg = ge.load_package(); dev = torch.device("cuda:0")
net = pkg.DGCNNWithColor(13, k=20).to(dev)
bucket = pkg.train.FlatGradBucket(net, steal_grads=True)
opt = torch.optim.Adam(net.parameters(), lr=1e-3, capturable=True, fused=True)
host = [tuple(t.pin_memory() for t in pkg.synthetic.s3dis_blocks(16, 4096, s)) for s in range(4)]
devb = [tuple(t.to(dev) for t in h) for h in host]
inp = lambda p: p[:, :, :6].transpose(1, 2)
def loss_of(m, p, l, n): return pkg.train.masked_onehot_cross_entropy(m(inp(p))[0], l, n)
for i in range(3):
    bucket.zero(); loss_of(net, *devb[i]).backward(); pkg.ops.join_aux(); bucket.all_reduce_mean(); opt.step()
step = pkg.train.GraphedTrainStep(net, opt, bucket, loss_of, devb[0], warmup=2)

import pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
stop = threading.Event()
def poll(dt):
    while not stop.is_set():
        pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM); pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
        time.sleep(dt)

def region(prefetch):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(10):
        loss = step(*host[i % 4])
        if prefetch: step.prefetch(*host[(i + 1) % 4])
        loss.item()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10

for name, dt, pf in [("no nvml, no prefetch", None, False), ("no nvml, prefetch", None, True), ("nvml 20 ms, prefetch", 0.02, True),
                     ("nvml 4 ms, prefetch", 0.004, True), ("no nvml, prefetch", None, True), ("no nvml, no prefetch", None, False),
                     ("nvml 20 ms, no prefetch", 0.02, False)]:
    stop.clear(); th = None
    if dt: th = threading.Thread(target=poll, args=(dt,), daemon=True); th.start()
    for _ in range(3): region(pf)
    ts = sorted(region(pf) for _ in range(40))
    stop.set()
    if th: th.join()
    print(f"{name:26s} median {ts[20]:.3f}  p90 {ts[36]:.3f}  worst {ts[-1]:.3f} ms/step   regions above median + 0.1: {sum(t > ts[20] + 0.1 for t in ts)}")
