"""What does the fused tanh policy head cost?  64 tiles of a one-term (policy) network through pqlb_mlp_forward_h
with (a) the scalar Q head, (b) the policy head, (c) the policy head + noise, each with and without activation stores."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pql_b200 import _kernels as K  # noqa: E402
from pql_b200 import _lib  # noqa: E402
from tools.fwd_bench import timeit  # noqa: E402

DEV = "cuda:0"
g = torch.Generator(device=DEV).manual_seed(0)
k_in, A, M = 88, 16, 8192
dims = [(512, k_in), (256, 512), (128, 256), (A, 128)]
ws = [torch.randn(o, l, device=DEV, generator=g) * 0.05 for o, l in dims]
bs = [torch.randn(o, device=DEV, generator=g) * 0.05 for o, _ in dims]
hs = []
for w in ws:
    hi = torch.zeros(w.numel() + 8, dtype=torch.float16, device=DEV)
    _lib.call("pqlb_split_f16", _lib.ptr(w), _lib.ptr(hi), None, w.numel())
    hs.append(hi)
x = torch.randn(M, 104, device=DEV, generator=g)
h = [torch.zeros(M, n, device=DEV) for n in (512, 256, 128)]
q = torch.zeros(M, device=DEV)
out, out2 = torch.zeros(M, 104, device=DEV), torch.zeros(M, 104, device=DEV)
noise = torch.randn(M, A, device=DEV, generator=g)
qw = torch.randn(128, device=DEV, generator=g)
for head in ("q", "policy", "policy+noise", "none"):
    for store in (0, 1):
        d = dict(x=K.addr(x), ldx=104, w1h=hs[0].data_ptr(), ldw1=k_in, w2h=hs[1].data_ptr(), w3h=hs[2].data_ptr(),
                 b1=K.addr(bs[0]), b2=K.addr(bs[1]), b3=K.addr(bs[2]), terms=1, k_in=k_in)
        if head == "q":
            d.update(head_w=K.addr(qw), head_b=K.addr(bs[3]), q=K.addr(q))
        elif head.startswith("policy"):
            d.update(act_wh=hs[3].data_ptr(), act_b=K.addr(bs[3]), act_n=A, act_out=K.addr(out, 88), act_ldo=104,
                     act_out2=K.addr(out2, 88), act_ldo2=104)
            if head.endswith("noise"):
                d.update(act_noise=K.addr(noise), act_ldnoise=A, noise_std=0.8, noise_bound=0.2)
        if store or head == "none":
            d.update(h1=K.addr(h[0]), h2=K.addr(h[1]), h3=K.addr(h[2]))
        us = timeit(K.MlpForwardH(M, k_in, [d]))
        print(f"head {head:13s} store {store}: {us:7.2f} us", flush=True)
