"""torchrun --nproc-per-node N tools/dp_oracle_check.py: the N-rank data-parallel critic update (fused
exchange inside the optimiser kernel, or --dp nccl) against ONE reference update on the concatenated
batch (oracle/learner.py, fp32), SURVEY 8e:
  * every gradient tensor of the exchanged mean gradient within 1e-3 of the oracle's,
  * parameters, Polyak target and operand copies bit-identical across ranks after the step,
  * the data-parallel RunningMeanStd equal to the reference update on the concatenated observations
    and bit-identical across ranks.
Prints one line 'DP ORACLE CHECK PASSED|FAILED' from rank 0 (tests/test_gpu_dp.py asserts on it)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
from oracle import actor as OA  # noqa: E402
from oracle import learner as L  # noqa: E402
from pql_b200.algo import PQLVLearner, _dp  # noqa: E402
from pql_b200.models import TanhMLPPolicy  # noqa: E402
from pql_b200.utils import RunningMeanStd  # noqa: E402
from tests import parity  # noqa: E402
from tests.golden import inputs  # noqa: E402

mode = "nccl" if "--dp" in sys.argv and sys.argv[sys.argv.index("--dp") + 1] == "nccl" else "fused"
ok = True
report = []
for distl, Bl in ((False, 2048), (True, 1024)):
    O, A = 88, 16
    B = Bl * world
    case = inputs.learner_case(77 + int(distl), B, O, A, distl)
    sl = slice(rank * Bl, (rank + 1) * Bl)
    cfg = parity.make_cfg(Bl, distl, local, memory=Bl)
    cfg.data_parallel, cfg.dp_fused = True, mode == "fused"
    torch.manual_seed(1 + rank)              # different default initialisation per rank: the constructor must broadcast rank 0's
    v = PQLVLearner(O, A, cfg, process_group=dist.new_group(list(range(world))))
    parity.load_params(v.critic.net_q1, case["q1"]); parity.load_params(v.critic.net_q2, case["q2"])
    actor = TanhMLPPolicy(O, A).to(dev)
    parity.load_params(actor, case["actor"])
    norm = (case["norm"][0].to(dev), case["norm"][1].to(dev), case["norm"][2])
    v.update(actor, tuple(x[sl].to(dev) for x in case["batch"]), norm, 0)
    plan = v._plan
    idx = torch.arange(Bl)
    with parity.injected_draws(idx, case["noises"][0][sl]):
        v.learn()
    torch.cuda.synchronize()
    ov = L.VLearnerOracle(case["q1"], case["q2"], distl=distl)
    ov.learn(case["batch"], case["noises"][0], case["actor"], case["norm"])
    summed = plan.dp.red if plan.dp is not None else plan.opt.grad
    worst = 0.0
    gi = 0
    for net in range(2):
        for layer in range(4):
            gw, gb = parity.unflatten(plan.Lc, summed / world, net, layer)
            for got in (gw, gb):
                ref = ov.last["grads"][gi]
                if ref.numel() > 1:
                    worst = max(worst, parity.rel(got, ref))
                gi += 1
    in_sync = all(_dp.params_in_sync(t) for t in (v.critic.arena.flat, plan.t_flat, plan.c_tf)) and \
        (plan.c_h is None or _dp.params_in_sync(plan.c_h.view(torch.int16).float()))
    good = worst <= 1e-3 and in_sync and plan.dp is not None if mode == "fused" else worst <= 1e-3 and in_sync
    ok &= bool(good)
    report.append(f"{'C51' if distl else 'twin-Q'} B={Bl}x{world} [{mode}, fwd {plan.fwd_mode}]: max gradient error vs the oracle on the "
                  f"concatenated batch {worst:.2e}, replicas in sync {in_sync}")
    plan.graphs = None

# ---- data-parallel observation normaliser
E, O = 512, 88
g = torch.Generator().manual_seed(3)
xs = [torch.randn(world * E, O, generator=g) * 2 + 0.5 for _ in range(3)]
rms = RunningMeanStd(shape=(O,), device=dev, process_group=dist.group.WORLD)
ref = OA.RunningMeanStdOracle(shape=(O,))
for x in xs:
    rms.update(x[rank * E:(rank + 1) * E].to(dev))
    ref.update(x)
torch.cuda.synchronize()
em = (rms.mean.cpu() - ref.mean).abs().max().item()
ev = ((rms.var.cpu() - ref.var).abs() / ref.var).max().item()
rms_ok = em <= 2e-6 and ev <= 1e-5 and rms.count == ref.count and _dp.params_in_sync(rms.mean) and _dp.params_in_sync(rms.var)
ok &= bool(rms_ok)
report.append(f"RunningMeanStd over {world} ranks vs the reference update on the concatenated batch: mean {em:.1e}, var {ev:.1e} "
              f"(rel), count {rms.count} == {ref.count}, ranks identical: {rms_ok}")
t = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    for line in report:
        print(line, flush=True)
    print("DP ORACLE CHECK", "PASSED" if t.item() == 1.0 else "FAILED", flush=True)
torch.cuda.synchronize()
dist.barrier()
os._exit(0 if t.item() == 1.0 else 1)
