"""One insert and one gather per size (K1 / K2 alone) between cudaProfilerStart/Stop, for
`ncu --set full --profile-from-start off` (DRAM bytes of the replay kernels; tools/ncu_summary.py condenses it)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pql_b200 import _lib
from pql_b200.replay import ReplayBuffer
O, A, CAP, E, B = 88, 16, 1_000_000, 4096, 8192
dev = torch.device("cuda:0")
mem = ReplayBuffer(capacity=CAP, obs_dim=O, action_dim=A, device=dev)
gen = torch.Generator(device=dev).manual_seed(1)
def rows(n):
    return (torch.randn(n, O, device=dev, generator=gen), torch.rand(n, A, device=dev, generator=gen),
            torch.randn(n, 1, device=dev, generator=gen), torch.randn(n, O, device=dev, generator=gen), torch.zeros(n, 1, device=dev))
ins = {n: rows(n) for n in (E, 30 * E, 120 * E)}
idx = {n: torch.randint(CAP, (n,), device=dev) for n in (B, 8 * B, 32 * B)}
out = {n: mem.gather(idx[n]) for n in idx}
p = 777
for n, r in ins.items():          # warm-up (un-profiled)
    _lib.call("pqlb_ring_insert", _lib.ptr(mem.ring), CAP, O, A, *(_lib.ptr(x) for x in r), n, p); p += n
torch.cuda.synchronize()
torch.cuda.profiler.start()
for n, r in ins.items():
    _lib.call("pqlb_ring_insert", _lib.ptr(mem.ring), CAP, O, A, *(_lib.ptr(x) for x in r), n, p % (CAP - n)); p += n
for n in idx:
    _lib.call("pqlb_sample_gather", _lib.ptr(mem.ring), CAP, O, A, _lib.ptr(idx[n]), n, *(_lib.ptr(x) for x in out[n]))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
