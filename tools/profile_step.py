"""One super-step (1 insert + 8 critic + 4 actor updates, batch 8192, AllegroHand shape) with
CUDA graphs off, bracketed by cudaProfilerStart/Stop, for `ncu --profile-from-start off`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PQLB_NO_GRAPH"] = "1"
import numpy as np
import torch
import bench
from pql_b200.algo import PQLPLearner, PQLVLearner
from pql_b200.replay import NStepReplay
from pql_b200.utils import default_pql_cfg

O, A, E, B = bench.O, bench.A, bench.E, bench.B
CAP = int(os.environ.get("CAP", 200_000))
dev = torch.device("cuda:0")
cfg = default_pql_cfg(batch_size=B, memory_size=CAP, num_envs=E)
v, p = PQLVLearner(O, A, cfg), PQLPLearner(O, A, cfg)
ns = NStepReplay(O, A, num_envs=E, nstep=3, device=dev)
rs = np.random.RandomState(0)
blocks = [tuple(torch.from_numpy(x).to(dev) for x in bench.synth_block(rs)) for _ in range(4)]
norm = (torch.zeros(O, device=dev), torch.ones(O, device=dev), 1e-4)
traj = ns.add_to_buffer(*(torch.from_numpy(x).to(dev) for x in bench.synth_block(rs, 32)))
critic, actor = v.start()[0], p.start()[0]
v.update(actor, traj, norm, 0); p.update(critic, traj[0], norm, 0)
while not v.memory.if_full:
    v.memory.add_to_buffer(ns.add_to_buffer(*blocks[0]))

def super_step(k):
    tr = ns.add_to_buffer(*blocks[k % 4])
    v.update(actor, tr, norm, 0); p.update(critic, tr[0], norm, 0)
    for j in range(8):
        v.learn()
        if j % 2 == 1:
            p.learn()

super_step(0); super_step(1)
torch.cuda.synchronize()
torch.cuda.profiler.start()
super_step(2)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
