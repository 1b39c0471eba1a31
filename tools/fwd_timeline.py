"""clock64 timeline of one tile of the split-fp16 forward (CTA-per-tile schedule, CTA (0,0)):
when the MMA issuer finished issuing each phase and when the first conversion warp saw each
accumulator complete / finished each conversion.  python tools/fwd_timeline.py [terms] [tiles] [store]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pql_b200 import _kernels as K, _lib
from tools.fwd_bench import net
DEV = "cuda:0"
terms = int(sys.argv[1]) if len(sys.argv) > 1 else 3
tiles = int(sys.argv[2]) if len(sys.argv) > 2 else 1
store = int(sys.argv[3]) if len(sys.argv) > 3 else 0
g = torch.Generator(device=DEV).manual_seed(0)
k_in, M = 104, 128 * tiles
x = torch.randn(M, k_in, device=DEV, generator=g)
ws, bs, hs, ls = net(k_in, g)
q = torch.zeros(M, device=DEV)
h = [torch.zeros(M, n, device=DEV) for n in (512, 256, 128)]
d = dict(x=K.addr(x), ldx=k_in, w1h=hs[0].data_ptr(), ldw1=k_in, w2h=hs[1].data_ptr(), w3h=hs[2].data_ptr(), b1=K.addr(bs[0]),
         b2=K.addr(bs[1]), b3=K.addr(bs[2]), head_w=K.addr(ws[3]), head_b=K.addr(bs[3]), q=K.addr(q), terms=terms)
if terms == 3:
    d.update(w1l=ls[0].data_ptr(), w2l=ls[1].data_ptr(), w3l=ls[2].data_ptr())
if store:
    d.update(h1=K.addr(h[0]), h2=K.addr(h[1]), h3=K.addr(h[2]))
lib = _lib.load()
lib.pqlb_mlp_forward_h_mode(1)
call = K.MlpForwardH(M, k_in, [d])
for _ in range(3):
    call()
buf = torch.zeros(64, dtype=torch.int64, device=DEV)
lib.pqlb_mlp_forward_h_debug(buf.data_ptr())
call()
torch.cuda.synchronize()
lib.pqlb_mlp_forward_h_debug(None)
t = buf.tolist()
t0 = min(v for v in t if v)
names_m = ["start", "x_conv seen", "L1q0 issued", "L1q1 issued", "p_conv0 seen", "L2c0 issued", "L1q2 issued", "p_conv1 seen", "L2c1 issued",
           "L1q3 issued", "p_conv2", "L2c2 issued", "p_conv3", "L2c3 issued", "L3 issued"]
names_e = ["start", "x_full seen", "x converted", "p_full q0", "conv q0 done", "p_full q1", "conv q1 done", "p_full q2", "conv q2 done",
           "p_full q3", "conv q3 done", "y_full", "conv Y0 done", "conv Y1 done", "z_full", "end"]
print(f"terms={terms} tiles={tiles} store={store}  (cycles since the first stamp; 1965 cycles = 1 us)")
for lab, off, names in (("MMA issuer", 0, names_m), ("conversion warp 0", 32, names_e)):
    print(lab)
    for i, n in enumerate(names):
        if t[off + i]:
            print(f"  {n:16s} {t[off + i] - t0:8d}")
