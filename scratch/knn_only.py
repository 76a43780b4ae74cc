import sys, torch
sys.path.insert(0, '/root/repo')
import __graft_entry__ as ge
pkg = ge.load_package(); dev = torch.device('cuda:0')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
x = (torch.randn(B, 4096, 64, device=dev) * 0.65 + 0.36).transpose(1, 2)     # point-major memory, (B,F,N) view
for _ in range(4): idx = pkg.ops.knn_graph(x, 20)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(10): idx = pkg.ops.knn_graph(x, 20)
ev1.record(); torch.cuda.synchronize()
print("knn_graph ms/call", ev0.elapsed_time(ev1) / 10)
