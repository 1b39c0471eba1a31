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
for n in (E, 30 * E, 120 * E):
    r = rows(n)
    t = ev_time(lambda: _lib.call("pqlb_ring_insert", _lib.ptr(mem.ring), CAP, O, A, *(_lib.ptr(x) for x in r), n, 12345), 50)
    print(f"insert {n:7d} rows: {t*1e3:8.2f} us  {n*1549/(t*1e-3)/1e9:8.1f} GB/s  ({n*1549/(t*1e-3)/1e9/6547.5:.3f} of HBM copy peak)")
for n in (B, 8 * B, 32 * B):
    idx = torch.randint(CAP, (n,), device=dev)
    out = mem.gather(idx)
    t = ev_time(lambda: _lib.call("pqlb_sample_gather", _lib.ptr(mem.ring), CAP, O, A, _lib.ptr(idx), n, *(_lib.ptr(x) for x in out)), 50)
    print(f"gather {n:7d} rows: {t*1e3:8.2f} us  {n*1557/(t*1e-3)/1e9:8.1f} GB/s  ({n*1557/(t*1e-3)/1e9/6547.5:.3f} of HBM copy peak)")
