"""Launch-time study of the fused forward kernels (CUDA events, 30 launches each): how a launch's time
depends on the number of 128-row tiles in flight (1 tile = latency of one tile, 148 = one full wave,
more = waves / contention), for the split-fp16 kernel (terms 1 and 3) and the TF32 kernel.

    python tools/fwd_bench.py
    python tools/fwd_bench.py --k-in 231 --quick      # the wide-input instantiation (ShadowHand critics), three launch shapes
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pql_b200 import _kernels as K  # noqa: E402
from pql_b200 import _lib  # noqa: E402

DEV = "cuda:0"


def net(k_in, g):
    ld = (k_in + 7) // 8 * 8 if k_in > 128 else (k_in + 3) // 4 * 4
    dims = [(512, ld), (256, 512), (128, 256), (1, 128)]
    ws = [torch.randn(o, l, device=DEV, generator=g) * 0.05 for o, l in dims]
    bs = [torch.randn(o, device=DEV, generator=g) * 0.05 for o, _ in dims]
    hs, ls = [], []
    for w in ws:
        hi = torch.zeros(w.numel() + 8, dtype=torch.float16, device=DEV)
        lo = torch.zeros(w.numel() + 8, dtype=torch.float16, device=DEV)
        _lib.call("pqlb_split_f16", _lib.ptr(w), _lib.ptr(hi), _lib.ptr(lo), w.numel())
        hs.append(hi); ls.append(lo)
    return ws, bs, hs, ls


def timeit(call, n=int(os.environ.get("FWD_BENCH_N", 30))):
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        call()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--k-in", type=int, default=104, help="input width (> 128: the wide-input kernel; no TF32 variant there)")
    ap.add_argument("--quick", action="store_true", help="the launch shapes of a training step only")
    args = ap.parse_args()
    g = torch.Generator(device=DEV).manual_seed(0)
    k_in = args.k_in
    ld = (k_in + 7) // 8 * 8 if k_in > 128 else (k_in + 3) // 4 * 4
    keep = []
    print(f"{'kernel':10s} {'tiles':>6s} {'groups':>6s} {'store':>5s} {'us':>8s} {'us/wave':>8s}")
    shapes = ((1, 1, 0), (1, 1, 1), (16, 1, 0), (74, 1, 0), (148, 1, 0), (148, 1, 1), (64, 2, 0), (64, 4, 0), (64, 4, 1),
              (296, 1, 0), (64, 2, 1))
    if args.quick:      # store = 2: the first half of the groups stores its activations (a critic update: current critics store, targets do not)
        shapes = ((64, 4, 2), (64, 2, 1), (64, 1, 0))
    for tiles, groups, store in shapes:
        M = 128 * tiles
        x = torch.zeros(M, ld, device=DEV)
        x[:, :k_in] = torch.randn(M, k_in, device=DEV, generator=g)
        nets = [net(k_in, g) for _ in range(groups)]
        h = [[torch.zeros(M, n, device=DEV) for n in (512, 256, 128)] for _ in range(groups)]
        q = [torch.zeros(M, device=DEV) for _ in range(groups)]
        keep.append((x, nets, h, q))
        for name, terms in (("f16x3", 3), ("f16x1", 1), ("tf32", 0)):
            if terms == 0 and k_in > 128:
                continue
            grp = []
            for i, (ws, bs, hs, ls) in enumerate(nets):
                if terms:
                    d = dict(x=K.addr(x), ldx=ld, w1h=hs[0].data_ptr(), ldw1=ld, w2h=hs[1].data_ptr(), w3h=hs[2].data_ptr(),
                             b1=K.addr(bs[0]), b2=K.addr(bs[1]), b3=K.addr(bs[2]), head_w=K.addr(ws[3]), head_b=K.addr(bs[3]),
                             q=K.addr(q[i]), terms=terms)
                    if terms == 3:
                        d.update(w1l=ls[0].data_ptr(), w2l=ls[1].data_ptr(), w3l=ls[2].data_ptr())
                else:
                    d = dict(x=K.addr(x), ldx=ld, w1=K.addr(ws[0]), ldw1=ld, w2=K.addr(ws[1]), w3=K.addr(ws[2]),
                             b1=K.addr(bs[0]), b2=K.addr(bs[1]), b3=K.addr(bs[2]), head_w=K.addr(ws[3]), head_b=K.addr(bs[3]),
                             q=K.addr(q[i]))
                if store == 1 or (store == 2 and i < groups // 2):
                    d.update(h1=K.addr(h[i][0]), h2=K.addr(h[i][1]), h3=K.addr(h[i][2]))
                grp.append(d)
            call = K.MlpForwardH(M, k_in, grp) if terms else K.MlpForward(M, k_in, grp)
            us = timeit(call)
            waves = -(-tiles * groups // 148)
            print(f"{name:10s} {tiles * groups:6d} {groups:6d} {store:5d} {us:8.2f} {us / waves:8.2f}", flush=True)


if __name__ == "__main__":
    main()
