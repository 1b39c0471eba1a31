"""Replay kernels alone (K1 insert, K2 gather, fused critic-batch gather) at bench.py's sizes."""
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
def ev_time(fn, n):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
walk = {"p": 12345}
def insert(r, n):
    # consecutive calls land on consecutive slot ranges and walk the whole 1 GB ring (nothing is re-written while still in L2)
    _lib.call("pqlb_ring_insert", _lib.ptr(mem.ring), CAP, O, A, *(_lib.ptr(x) for x in r), n, walk["p"])
    walk["p"] = (walk["p"] + n) % (CAP - n)
for n in (E, 30 * E, 120 * E):
    r = rows(n)
    t = ev_time(lambda: insert(r, n), 50)
    print(f"insert {n:7d} rows: {t*1e3:8.2f} us  {n*1549/(t*1e-3)/1e9:8.1f} GB/s  ({n*1549/(t*1e-3)/1e9/6547.5:.3f} of HBM copy peak)")
turn = {"i": 0}
for n in (B, 8 * B, 32 * B):
    pool = [torch.randint(CAP, (n,), device=dev) for _ in range(8)]          # fresh indices every call
    out = mem.gather(pool[0])
    def gather():
        turn["i"] += 1
        idx = pool[turn["i"] % 8]
        _lib.call("pqlb_sample_gather", _lib.ptr(mem.ring), CAP, O, A, _lib.ptr(idx), n, *(_lib.ptr(x) for x in out))
    t = ev_time(gather, 50)
    print(f"gather {n:7d} rows: {t*1e3:8.2f} us  {n*1557/(t*1e-3)/1e9:8.1f} GB/s  ({n*1557/(t*1e-3)/1e9/6547.5:.3f} of HBM copy peak)")
