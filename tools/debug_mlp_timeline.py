import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pql_b200 import _kernels as K, _lib
lib = _lib.load()
dev = "cuda:0"
M = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ng = int(sys.argv[2]) if len(sys.argv) > 2 else 1
k_in = 104
g = torch.Generator(device=dev).manual_seed(0)
groups, keep = [], []
for i in range(ng):
    x = torch.randn(M, k_in, device=dev, generator=g); w1 = torch.randn(512, k_in, device=dev, generator=g) * 0.1
    w2 = torch.randn(256, 512, device=dev, generator=g) * 0.05; w3 = torch.randn(128, 256, device=dev, generator=g) * 0.05
    b = [torch.zeros(n, device=dev) for n in (512, 256, 128)]
    h = [torch.zeros(M, n, device=dev) for n in (512, 256, 128)]
    q = torch.zeros(M, device=dev); w4 = torch.randn(128, device=dev, generator=g); b4 = torch.zeros(1, device=dev)
    st = int(os.environ.get("STORE", "1"))
    groups.append(dict(x=K.addr(x), ldx=k_in, w1=K.addr(w1), ldw1=k_in, w2=K.addr(w2), w3=K.addr(w3), b1=K.addr(b[0]), b2=K.addr(b[1]),
                       b3=K.addr(b[2]), head_w=K.addr(w4), head_b=K.addr(b4), q=K.addr(q),
                       h1=K.addr(h[0]) if st else 0, h2=K.addr(h[1]) if st else 0, h3=K.addr(h[2]) if st else 0))
    keep += [x, w1, w2, w3, b, h, q, w4, b4]
lib.pqlb_mlp_forward_cluster(int(os.environ.get("CLUSTER", "0")))
call = K.MlpForward(M, k_in, groups)
call(); torch.cuda.synchronize()
dbg = torch.zeros(64, dtype=torch.int64, device=dev)
lib.pqlb_mlp_forward_debug.argtypes = [C.c_void_p]; lib.pqlb_mlp_forward_debug.restype = None
lib.pqlb_mlp_forward_debug(C.c_void_p(dbg.data_ptr()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); call(); e1.record(); torch.cuda.synchronize()
lib.pqlb_mlp_forward_debug(None)
t = dbg.cpu().tolist()
t0 = min(x for x in t if x > 0)
print("kernel us (with timeline stamps)", e0.elapsed_time(e1) * 1e3)
e0.record()
for _ in range(20): call()
e1.record(); torch.cuda.synchronize()
print("kernel us (avg of 20)", e0.elapsed_time(e1) * 1e3 / 20)
names_m = ["start", "x_full", "L1q0", "L1q1", "c0 wait", "c0 done", "L1q2", "c1 wait", "c1 done", "L1q3", "c2 wait", "c2 done", "c3 wait", "c3 done", "L3 done"]
print("MMA thread:")
for n, v in zip(names_m, t[:len(names_m)]):
    print(f"  {n:10s} {v - t0:8d}")
names_e = ["start", "q0 ready", "q0 conv", "q1 ready", "q1 conv", "q2 ready", "q2 conv", "q3 ready", "q3 conv", "y ready", "y0 conv", "y1 conv", "z ready", "end"]
print("epilogue warp 0:")
for n, v in zip(names_e, t[32:32 + len(names_e)]):
    print(f"  {n:10s} {v - t0:8d}")
