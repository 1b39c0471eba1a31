"""torchrun --nproc-per-node N tools/dp_check.py: the fused exchange (all-reduce inside the optimiser
kernel, symmetric memory) against the NCCL path on identical data - parameters must agree (bitwise at
N = 2, where the sum order cannot differ) and stay bit-identical across ranks."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=dev)
from pql_b200.algo import PQLPLearner, PQLVLearner, _dp
from pql_b200.replay import NStepReplay
from pql_b200.utils import default_pql_cfg


def run(mode, streams, B=1024, steps=3, distl=False):
    O, A, E = 88, 16, 256
    torch.manual_seed(5)
    cfg = default_pql_cfg(batch_size=B, memory_size=8192, num_envs=E, v_learner_gpu=local, p_learner_gpu=local, distl=distl)
    cfg.data_parallel = True
    cfg.learner_streams = streams
    cfg.dp_fused = mode == "fused"
    v = PQLVLearner(O, A, cfg, process_group=dist.new_group(list(range(world))))
    p = PQLPLearner(O, A, cfg, process_group=dist.new_group(list(range(world))))
    dist.broadcast(v.critic.arena.flat, 0); dist.broadcast(p.actor.arena.flat, 0)
    ns = NStepReplay(O, A, num_envs=E, nstep=3, device=dev)
    g = torch.Generator(device=dev).manual_seed(100 + rank)          # every rank its own envs / replay shard
    torch.manual_seed(1000 + rank)                                    # ... and its own sampling stream
    norm = (torch.zeros(O, device=dev), torch.ones(O, device=dev), 1e-4)
    critic, actor = v.start()[0], p.start()[0]
    for k in range(steps):
        T = 8 if k == 0 else 1
        blk = (torch.randn(E, T, O, device=dev, generator=g), torch.rand(E, T, A, device=dev, generator=g) * 2 - 1,
               torch.randn(E, T, 1, device=dev, generator=g) * 0.01, torch.randn(E, T, O, device=dev, generator=g),
               (torch.rand(E, T, 1, device=dev, generator=g) < 0.05).float())
        tr = ns.add_to_buffer(*blk)
        critic, vl, _ = v.update(actor, tr, norm, 0)
        actor, pl, _ = p.update(critic, tr[0], norm, 0)
        for j in range(4):
            v.learn()
            if j % 2 == 1:
                p.learn()
    torch.cuda.synchronize()
    c, a = v.critic.arena.flat.clone(), p.actor.arena.flat.clone()
    sync = _dp.params_in_sync(c) and _dp.params_in_sync(a)
    for l in (v, p):
        l._plan.graphs = None
    return c, a, sync

ok = True
for distl in (False, True):
    ref_c, ref_a, ref_sync = run("nccl", False, distl=distl)
    for streams in (False, True):
        c, a, sync = run("fused", streams, distl=distl)
        dc = ((c - ref_c).abs().max() / ref_c.abs().max()).item(); da = ((a - ref_a).abs().max() / ref_a.abs().max()).item()
        bit = torch.equal(c, ref_c) and torch.equal(a, ref_a)
        good = sync and ref_sync and (bit if world == 2 else max(dc, da) < 1e-5) and bool(torch.isfinite(c).all())
        ok &= good
        if rank == 0:
            print(f"distl={distl} streams={streams}: fused vs nccl max rel diff critic {dc:.2e} actor {da:.2e} bitwise={bit} "
                  f"ranks in sync: fused {sync} nccl {ref_sync} -> {'OK' if good else 'FAIL'}", flush=True)
t = torch.tensor([1.0 if ok else 0.0], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("DP CHECK", "PASSED" if t.item() == 1.0 else "FAILED", flush=True)
torch.cuda.synchronize(); dist.barrier()
os._exit(0 if t.item() == 1.0 else 1)
