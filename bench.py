#!/usr/bin/env python
"""Benchmark of the PQL learner hot path (BASELINE.json: critic updates/s at batch 8192 + replay GB/s).

One STEP is one env step's worth of learner work at the reference's default ratios
(pql_algo.yaml:17-18): n-step push + ring insert of 4096 env transitions, obs-ring insert,
8 critic updates (PQLVLearner.learn, batch 8192) and 4 actor updates (PQLPLearner.learn),
on synthetic AllegroHand-shaped transitions (obs 88, act 16), 1M-slot replay (800 MB > L2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 is launched under torchrun (one rank per GPU, NCCL): every rank owns its envs and replay
shard, critic / actor gradients are all-reduced before clip + AdamW (weak scaling).
``--impl reference`` times the CPU restatement of the reference path (oracle/, torch fp32 on
all host threads) on a bounded sample of the same workload.
"""
import argparse
import gc
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if "reference" in sys.argv:
    # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers: the CPU arm must use every host
    # core regardless (round-1 VERDICT: the N > 1 reference lines ran single-threaded)
    for _k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = str(os.cpu_count() or 1)

import numpy as np  # noqa: E402
import torch  # noqa: E402

O, A, E, B, CAP, NSTEP = 88, 16, 4096, 8192, 1_000_000, 3
V_PER_STEP, P_PER_STEP = 8, 4
FLOP_V = 3_684_352          # GEMM FLOPs per sample of one critic update
FLOP_P = 2_913_280          # ... of one actor update
# the dominant kernel, mlp_fwd_kernel (pqlb_mlp_forward): MACs per sample of its launches in one update
MAC_ACTOR_FWD = O * 512 + 512 * 256 + 256 * 128 + 128 * A                 # trunk + fused tanh policy head
MAC_CRITIC_FWD = (O + A) * 512 + 512 * 256 + 256 * 128 + 128              # trunk + scalar Q head
FWD_MAC_V = MAC_ACTOR_FWD + 4 * MAC_CRITIC_FWD       # critic update: target policy + 2 target + 2 current nets
FWD_MAC_P = MAC_ACTOR_FWD + 2 * MAC_CRITIC_FWD       # actor update: policy + frozen twin critics
BYTES_INSERT = 1549         # algorithmic bytes per inserted transition
BYTES_SAMPLE = 1557         # ... per sampled transition
N_BLOCKS = 16


def synth_block(rs, T=1):
    """One [E, T, *] block like PQLActor.explore_env produces (SURVEY 8d synthetic inputs)."""
    obs = rs.standard_normal((E, T, O)).astype(np.float32)
    nxt = rs.standard_normal((E, T, O)).astype(np.float32)
    act = rs.uniform(-1, 1, (E, T, A)).astype(np.float32)
    rew = (rs.standard_normal((E, T, 1)) * 0.01).astype(np.float32)
    done = (rs.uniform(0, 1, (E, T, 1)) < 0.01).astype(np.float32)
    return obs, act, rew, nxt, done


class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING the timed region through NVML (a thread
    polling every 10 ms: the timed region of a default run is ~50 ms, too short for nvidia-smi -lms)."""

    def __init__(self, index):
        self.rows, self.index, self.stop, self.th, self.max_mhz = [], index, False, None, None

    def __enter__(self):
        try:
            import pynvml as N
            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.max_mhz = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
            names = {"hw_slowdown": N.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": N.nvmlClocksEventReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": N.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": N.nvmlClocksEventReasonSwPowerCap}

            def poll():
                while not self.stop:
                    try:
                        mask = N.nvmlDeviceGetCurrentClocksEventReasons(h)
                        self.rows.append((float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)), N.nvmlDeviceGetPowerUsage(h) / 1000.0,
                                          [k for k, bit in names.items() if mask & bit]))
                    except Exception:
                        pass
                    time.sleep(0.01)
            self.th = threading.Thread(target=poll, daemon=True)
            self.th.start()
        except Exception:
            self.th = None
        return self

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x for x in vis.split(",") if x.strip() != ""]
            if self.index < len(ids) and ids[self.index].strip().isdigit():
                return int(ids[self.index])
        return self.index

    def __exit__(self, *a):
        self.stop = True
        if self.th is not None:
            self.th.join(timeout=1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        sm = [r[0] for r in self.rows]
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": self.max_mhz, "reasons": sorted({x for r in self.rows for x in r[2]}),
                "power_w_max": max(r[1] for r in self.rows), "samples": len(sm)}


def cpu_reference_rate(seconds_budget=20.0, threads=None):
    """The reference's CPU torch path (oracle restatement, every host thread) on a bounded sample:
    a few full-size updates (batch 8192, AllegroHand shape) at the 8 V : 4 P : 1 insert ratio, sampling
    from a full 1M-slot ring like the GPU arm."""
    from oracle import learner as L
    from oracle import replay as R
    torch.set_num_threads(int(threads or os.cpu_count() or 1))
    g = torch.Generator().manual_seed(0)
    q1, q2 = L.init_mlp(O + A, 1, g), L.init_mlp(O + A, 1, g)
    actor = L.init_mlp(O, A, g)
    v, p = L.VLearnerOracle(q1, q2), L.PLearnerOracle(actor)
    rs = np.random.RandomState(0)
    cap = CAP
    ring = R.RingOracle(cap, O, A)
    ns = R.NStepOracle(O, A, E, NSTEP, 0.99)
    ring.insert(*ns.push(*synth_block(rs, 8)))
    norm = (torch.zeros(O), torch.ones(O), 1e-4)
    blocks = [synth_block(rs) for _ in range(4)]
    fill = ns.push(*synth_block(rs, 16))             # 16 * E rows per insert until every slot is live
    while not ring.if_full:
        ring.insert(*fill)
    t_ins = t_v = t_p = 0.0
    n_steps = 0
    t_start = time.perf_counter()
    while True:
        t0 = time.perf_counter()
        ring.insert(*ns.push(*blocks[n_steps % 4]))
        t1 = time.perf_counter()
        for _ in range(V_PER_STEP):
            idx = rs.randint(0, ring.cur_capacity, size=B)
            batch = tuple(torch.from_numpy(x) for x in ring.gather(idx))
            v.learn(batch, torch.randn(B, A) * 0.8, actor, norm)
        t2 = time.perf_counter()
        for _ in range(P_PER_STEP):
            idx = rs.randint(0, ring.cur_capacity, size=B)
            p.learn(torch.from_numpy(ring.buf_obs[idx]), q1, q2, norm)
        t3 = time.perf_counter()
        t_ins += t1 - t0; t_v += t2 - t1; t_p += t3 - t2
        n_steps += 1
        if time.perf_counter() - t_start > seconds_budget or n_steps >= 50:
            break
    total = t_ins + t_v + t_p
    return dict(value=V_PER_STEP * n_steps / total, unit="critic updates/s", cores=torch.get_num_threads(), kind="port",
                sample=f"{n_steps} steps of (1 insert of {E} rows + {V_PER_STEP} critic + {P_PER_STEP} actor updates, batch {B}) "
                       f"through oracle/ (torch-CPU fp32 restatement of the reference path; full {cap}-slot numpy ring, "
                       f"numpy RandomState indices, {torch.get_num_threads()} torch threads)",
                ring_slots=cap, index_source="numpy RandomState.randint", torch_threads=torch.get_num_threads(),
                host_cpus=os.cpu_count(),
                ms_per_step=1e3 * total / n_steps, ms_per_critic_update=1e3 * t_v / (n_steps * V_PER_STEP),
                ms_per_actor_update=1e3 * t_p / (n_steps * P_PER_STEP), ms_per_insert=1e3 * t_ins / n_steps)


def cuda_eager_rate(dev, seconds_budget=6.0):
    """the reference's own arithmetic as eager PyTorch on the same B200 - the oracle
    learners with every tensor on ``dev``, fp32 SGEMM (allow_tf32 off, as the reference never enables
    it), torch.randint / torch.normal draws on the device, five index gathers per sample, five slice
    copies per insert and the reference's one host sync per update (``loss.item()``,
    pql_v_learner.py:111).  Baseline leg only: nothing of this is on the product path."""
    from oracle import learner as L
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator().manual_seed(0)
    mv = lambda ps: [(w.to(dev), b.to(dev)) for w, b in ps]          # noqa: E731
    q1, q2, actor = mv(L.init_mlp(O + A, 1, g)), mv(L.init_mlp(O + A, 1, g)), mv(L.init_mlp(O, A, g))
    v, p = L.VLearnerOracle(q1, q2), L.PLearnerOracle(actor)
    gen = torch.Generator(device=dev).manual_seed(0)
    widths = (O, A, 1, O, 1)
    ring = [torch.randn(CAP, w, device=dev, generator=gen) * (0.01 if i == 2 else 1.0) for i, w in enumerate(widths)]
    ring[4] = (torch.rand(CAP, 1, device=dev, generator=gen) < 0.01).float()
    rows = [torch.randn(E, w, device=dev, generator=gen) for w in widths]
    norm = (torch.zeros(O, device=dev), torch.ones(O, device=dev), 1e-4)

    def step(k):
        p0 = (k * E) % (CAP - E)
        for buf, x in zip(ring, rows):                                   # simple_replay.py:40-83
            buf[p0:p0 + E] = x
        for j in range(V_PER_STEP):
            idx = torch.randint(CAP, (B,), device=dev)                   # simple_replay.py:87
            batch = tuple(buf[idx] for buf in ring)
            noise = torch.normal(torch.zeros(B, A, device=dev), torch.full((B, A), 0.8, device=dev))
            v.learn(batch, noise, actor, norm)                           # ends with float(loss): the reference's .item()
            if (j + 1) % (V_PER_STEP // P_PER_STEP) == 0:
                idx = torch.randint(CAP, (B,), device=dev)
                p.learn(ring[0][idx], v.q1, v.q2, norm)
    for k in range(3):
        step(k)
    torch.cuda.synchronize(dev)
    n, t0 = 0, time.perf_counter()
    while True:
        step(3 + n)
        n += 1
        torch.cuda.synchronize(dev)
        if time.perf_counter() - t0 > seconds_budget or n >= 200:
            break
    dt = time.perf_counter() - t0
    return {"value": V_PER_STEP * n / dt, "unit": "critic updates/s", "ms_per_step": 1e3 * dt / n, "steps": n,
            "kind": "oracle/ learners as eager PyTorch on the same GPU: fp32 SGEMM (allow_tf32 = False), device-side "
                    "torch.randint / torch.normal, index gathers from five column tensors, one host sync per update",
            "torch": torch.__version__}


def config_dict(n_gpus, dp="fused"):
    return {"workload": "configs[1]: DoubleQ V-learner + P-learner, AllegroHand shape (obs 88, act 16), batch 8192, "
                        "1M-slot replay, n-step 3, 4096-env synthetic insert stream, Polyak target update",
            "step": f"1 n-step push + ring insert of {E} transitions, {V_PER_STEP} critic updates, {P_PER_STEP} actor updates",
            "critic_updates_per_step": V_PER_STEP, "actor_updates_per_step": P_PER_STEP, "batch_per_gpu": B,
            "num_envs_per_gpu": E, "replay_slots_per_gpu": CAP, "parallelism": f"dp{n_gpus}",
            "gradient_exchange": "none (1 GPU)" if n_gpus == 1 else
                                 {"fused": "two-shot all-reduce over symmetric memory fused into the optimiser kernel "
                                           "(pqlb_adamw_polyak_dp); --dp nccl / nccl-graph for the NCCL paths",
                                  "nccl": "ncclAllReduce between two CUDA graphs", "nccl-graph": "ncclAllReduce captured in the update's CUDA graph"}[dp],
            "settle_steps": 30,
            "learner_streams": "V-learner and P-learner each enqueue on their own CUDA stream (the reference runs them as "
                               "two concurrent Ray actors); update() is the exchange/join point",
            "cache": "replay ring 1 GB per GPU > 126 MB L2 (random gathers miss L2); weights/activations are the "
                     "step's own working set and are not flushed"}


def reference_config(r):
    """What the CPU arm actually runs (same workload and step definition as the B200 arm)."""
    return {"workload": "configs[1]: DoubleQ V-learner + P-learner, AllegroHand shape (obs 88, act 16), batch 8192, "
                        "1M-slot replay, n-step 3, 4096-env synthetic insert stream, Polyak target update",
            "step": f"1 n-step push + ring insert of {E} transitions, {V_PER_STEP} critic updates, {P_PER_STEP} actor updates",
            "critic_updates_per_step": V_PER_STEP, "actor_updates_per_step": P_PER_STEP, "batch": B, "num_envs": E,
            "replay_slots": r["ring_slots"], "parallelism": "one process, all host threads (no GPU, no data parallelism)",
            "implementation": "oracle/ : torch-CPU fp32 restatement of pql_v_learner.py:73-133 / pql_p_learner.py:47-96 / "
                              "simple_replay.py / nstep_replay.py (the reference itself needs Ray/Hydra/gym and cannot travel)",
            "index_source": r["index_source"], "torch_threads": r["torch_threads"], "host_cpus": r["host_cpus"],
            "sample": r["sample"]}


def run_reference(args, rank):
    if rank != 0:
        return
    r = cpu_reference_rate(seconds_budget=max(5.0, min(120.0, 4.0 * (args.steps + args.warmup))))
    line = {"metric": "critic updates/s (batch 8192)", "value": r["value"], "unit": "critic updates/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": reference_config(r), "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "critic updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "ms_per_critic_update": r["ms_per_critic_update"], "ms_per_actor_update": r["ms_per_actor_update"],
            "ms_per_insert": r["ms_per_insert"], "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="skip the configs[2] / configs[3] sub-records and the eager-CUDA leg")
    ap.add_argument("--repeats", type=int, default=5, help="timed blocks of --steps steps each; value = the median block")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--one-stream", action="store_true", help="both learners on the caller's stream (no overlap)")
    ap.add_argument("--dp", default="fused", choices=["fused", "nccl", "nccl-graph"],
                    help="gradient exchange: fused into the optimiser kernel over symmetric memory (default), "
                         "NCCL all-reduce between two CUDA graphs, or NCCL captured inside the update graph")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3
    import __graft_entry__ as entry
    if rank == 0 and not os.path.exists(entry.LIB):
        entry.build()
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries the one JSON line; NCCL_DEBUG is whatever the launcher set (never touched here)
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
    from pql_b200 import _kernels as K
    from pql_b200 import _lib
    from pql_b200.algo import PQLPLearner, PQLVLearner
    from pql_b200.replay import NStepReplay
    from pql_b200.utils import default_pql_cfg

    torch.manual_seed(42 + rank)
    cfg = default_pql_cfg(batch_size=B, memory_size=CAP, num_envs=E, v_learner_gpu=local, p_learner_gpu=local)
    cfg.data_parallel = world > 1
    cfg.use_cuda_graph = not args.no_graph
    cfg.learner_streams = not args.one_stream      # V- and P-learner overlap like the reference's two Ray actors
    cfg.dp_fused = world > 1 and args.dp == "fused"             # all-reduce inside the optimiser kernel (symmetric memory)
    cfg.dp_graph_allreduce = world > 1 and args.dp == "nccl-graph"
    # one communicator per learner: their updates replay on two streams, and an all-reduce captured in
    # a CUDA graph must not share NCCL's per-communicator ordering with the other learner's
    pg_v = dist.new_group(list(range(world))) if world > 1 else None
    pg_p = dist.new_group(list(range(world))) if world > 1 else None
    v = PQLVLearner(O, A, cfg, process_group=pg_v)      # the constructors broadcast rank 0's initial weights
    p = PQLPLearner(O, A, cfg, process_group=pg_p)
    ns = NStepReplay(O, A, num_envs=E, nstep=NSTEP, device=dev, gamma=0.99)
    rs = np.random.RandomState(42 + rank)
    host_blocks = [tuple(torch.from_numpy(x).pin_memory() for x in synth_block(rs)) for _ in range(N_BLOCKS)]
    dev_blocks = [tuple(x.to(dev) for x in blk) for blk in host_blocks]
    h2d_bytes = sum(x.numel() * 4 for x in host_blocks[0])
    all_obs = torch.cat([b[0].reshape(-1, O) for b in host_blocks])
    norm = (all_obs.mean(0).to(dev), all_obs.var(0).to(dev), 1e-4)      # RunningMeanStd.get_states stand-in
    if world > 1:
        # one normaliser for the whole job: the per-rank (mean, var, count) merged with the
        # reference's parallel-variance formula (torch_util.py:91-103) - pql_b200.algo._dp.merge_moments
        from pql_b200.algo._dp import merge_moments
        m, var_, _ = merge_moments(norm[0], norm[1], float(all_obs.shape[0]))
        norm = (m, var_, 1e-4)

    # warm-up block of 32 steps (train_pql.py:58) then fill the ring so that sampling spans all 800 MB
    warm = tuple(torch.from_numpy(x).to(dev) for x in synth_block(rs, 32))
    traj = ns.add_to_buffer(*warm)
    critic, _, _ = v.start()
    actor, _, _ = p.start()
    v.update(actor, traj, norm, 0)
    p.update(critic, traj[0], norm, 0)
    i = 0
    while not v.memory.if_full:
        traj = ns.add_to_buffer(*dev_blocks[i % N_BLOCKS]); i += 1
        v.memory.add_to_buffer(traj)
    state = {"actor": actor, "critic": critic, "d2h": 0}

    def super_step(k, host_inputs):
        blk = host_blocks[k % N_BLOCKS] if host_inputs else dev_blocks[k % N_BLOCKS]
        if host_inputs:
            blk = tuple(x.to(dev, non_blocking=True) for x in blk)
        tr = ns.add_to_buffer(*blk)
        state["critic"], vloss, _ = v.update(state["actor"], tr, norm, 0)     # returns the loss mean: D2H
        state["actor"], ploss, _ = p.update(state["critic"], tr[0], norm, 0)
        for j in range(V_PER_STEP):
            v.learn()
            if (j + 1) % (V_PER_STEP // P_PER_STEP) == 0:
                p.learn()
        return vloss, ploss

    def timed(n, host_inputs, base):
        # no cyclic-GC pass inside a timed region: a generation-2 collection over the interpreter's heap
        # takes 10-20 ms - a third of a 20-step region (seen as one rank stalling the others at 2 GPUs)
        gc.collect()
        gc.disable()
        try:
            return _timed(n, host_inputs, base)
        finally:
            gc.enable()

    def _timed(n, host_inputs, base):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for k in range(n):
            losses = super_step(base + k, host_inputs)
        for lrn in (v, p):                       # the step ends when both learners' streams have drained
            if lrn.stream is not None:
                torch.cuda.current_stream(dev).wait_stream(lrn.stream)
        e1.record()
        torch.cuda.synchronize(dev)
        wall = time.perf_counter() - t0
        if world > 1:
            dist.barrier()
        ms = max(e0.elapsed_time(e1), 0.0)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, wall, losses

    timed(args.warmup, False, 0)
    # settling beyond the W warm-up steps (reported in config.settle_steps): GPUs 1..N-1 idle through
    # the set-up and have not reached their boost clock, and the peer-memory / NCCL paths are cold;
    # without it the first timed block of an 8-GPU run measured 6 % below the second
    SETTLE = 30
    timed(SETTLE, False, args.warmup)
    def n_launches():      # C-ABI launches + the kernels replayed from the learners' CUDA graphs
        return _lib.launch_count() + sum(getattr(l._plan, "graph_launches", 0) for l in (v, p) if l._plan is not None)
    launches0 = n_launches()
    reps = max(1, args.repeats)
    blocks_ms, blocks_wall = [], []
    with ClockSampler(local) as clk:
        for r in range(reps):        # each block: EXACTLY --steps steps between barrier + synchronize, max over ranks
            ms_r, wall_r, losses = timed(args.steps, False, args.warmup + r * args.steps)
            blocks_ms.append(ms_r); blocks_wall.append(wall_r)
    launches = (n_launches() - launches0) // reps
    clocks = clk.summary()
    order = sorted(range(reps), key=lambda i: blocks_ms[i])
    ms, wall = blocks_ms[order[reps // 2]], blocks_wall[order[reps // 2]]       # the median block is the reported one
    timed(2, True, 0)
    e2e_blocks = [timed(args.steps, True, args.warmup + r * args.steps)[0] for r in range(min(reps, 3))]
    ms_e2e = sorted(e2e_blocks)[len(e2e_blocks) // 2]
    # like-for-like with the reference's blocking loss read: update() waits for the learner's stream
    # and returns the current loss mean (cfg.sync_loss = True) instead of the previous step's
    v._sync_loss = p._sync_loss = True
    timed(2, False, 0)
    sync_blocks = [timed(args.steps, False, args.warmup + r * args.steps)[0] for r in range(min(reps, 3))]
    ms_sync = sorted(sync_blocks)[len(sync_blocks) // 2]
    v._sync_loss = p._sync_loss = False
    # data parallel: every rank must hold bit-identical parameters after the run (checked on the device)
    sync_check = None
    if world > 1:
        from pql_b200.algo import _dp
        torch.cuda.synchronize(dev)
        ok_c = _dp.params_in_sync(v.critic.arena.flat); ok_a = _dp.params_in_sync(p.actor.arena.flat)
        ok_t = _dp.params_in_sync(v._plan.t_flat)
        sync_check = {"critic": ok_c, "actor": ok_a, "critic_target": ok_t,
                      "critic_checksum": int(v.critic.arena.flat.view(torch.int32).to(torch.int64).sum().item()),
                      "actor_checksum": int(p.actor.arena.flat.view(torch.int32).to(torch.int64).sum().item()),
                      "updates": int(v._plan.opt.step)}
        assert ok_c and ok_a and ok_t, f"data-parallel replicas diverged: {sync_check}"

    # ---- per-kernel device time over the same super-step (CUDA events around every launch)
    K.PROFILE = {}
    prev_graph = getattr(v, "_graph_off", None)
    v.disable_graph(); p.disable_graph()
    prof_steps = 3
    for k in range(prof_steps):
        super_step(k, False)
    torch.cuda.synchronize(dev)
    per_kernel = {name: sum(a.elapsed_time(b) for a, b in evs) / prof_steps for name, evs in K.PROFILE.items()}
    n_gemm = sum(len(K.PROFILE.get(k, [])) for k in ("pqlb_gemm_tf32", "pqlb_mlp_forward", "pqlb_mlp_forward_h", "pqlb_mlp_backward")) // prof_steps
    len_fwd = {k: len(x) // prof_steps for k, x in K.PROFILE.items()}
    K.PROFILE = None
    v.enable_graph(); p.enable_graph()

    # ---- replay kernels alone (HBM roofline): single calls and L2-exceeding calls
    def ev_time(fn, n):
        fn(); torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / n
    rows1 = tuple(x.reshape(E, -1) for x in dev_blocks[0])
    big_n = 30 * E
    gen = torch.Generator(device=dev).manual_seed(1)
    rows_big = (torch.randn(big_n, O, device=dev, generator=gen), torch.rand(big_n, A, device=dev, generator=gen),
                torch.randn(big_n, 1, device=dev, generator=gen), torch.randn(big_n, O, device=dev, generator=gen),
                torch.zeros(big_n, 1, device=dev))
    mem = v.memory
    idx1s = [torch.randint(CAP, (B,), device=dev) for _ in range(8)]         # fresh indices every call
    idx8s = [torch.randint(CAP, (8 * B,), device=dev) for _ in range(8)]
    out1 = mem.gather(idx1s[0]); out8 = mem.gather(idx8s[0])
    turn = {"i": 0}

    def next_idx(pool):
        turn["i"] += 1
        return pool[turn["i"] % len(pool)]

    walk = {"p": 12345}

    def raw_insert(rows):
        # consecutive calls land on consecutive slot ranges and walk the whole 1 GB ring, like the real
        # insert stream: nothing is re-written while it is still in L2
        n = rows[0].shape[0]
        _lib.call("pqlb_ring_insert", _lib.ptr(mem.ring), CAP, O, A, *(_lib.ptr(x) for x in rows), n, walk["p"])
        walk["p"] = (walk["p"] + n) % (CAP - n)

    def raw_gather(idx, out):
        _lib.call("pqlb_sample_gather", _lib.ptr(mem.ring), CAP, O, A, _lib.ptr(idx), idx.numel(), *(_lib.ptr(x) for x in out))
    replay = {}
    for name, fn, units, per in (("insert_4096", lambda: raw_insert(rows1), E, BYTES_INSERT),
                                 ("insert_122880", lambda: raw_insert(rows_big), big_n, BYTES_INSERT),
                                 ("sample_8192", lambda: raw_gather(next_idx(idx1s), out1), B, BYTES_SAMPLE),
                                 ("sample_65536", lambda: raw_gather(next_idx(idx8s), out8), 8 * B, BYTES_SAMPLE)):
        t_ms = ev_time(fn, 200 if units <= 8 * B else 50)
        replay[name] = {"us": round(t_ms * 1e3, 2), "gbs": round(units * per / (t_ms * 1e-3) / 1e9, 1)}

    def leave():
        """Tear-down of a data-parallel run.  The learners' CUDA graphs hold captured NCCL kernels;
        destroying the communicators under them hung in round 1, so: drop the graphs, drain the
        device, meet at a barrier and leave without NCCL's destructor."""
        if world > 1:
            for l in (v, p):
                if l._plan is not None:
                    l._plan.graphs = None
            torch.cuda.synchronize(dev)
            dist.barrier()
            sys.stdout.flush(); sys.stderr.flush()
            os._exit(0)

    # ---- tcgen05 issue throughput with all SMs busy and resident operands (csrc/peak.cu): the roofline denominator
    def mma_peak(kind, n, iters=4000):
        t_ms = ev_time(lambda: _lib.call("pqlb_mma_peak", kind, n, iters), 5)
        return 148 * iters * 4 * 2.0 * 128 * n * (8 if kind == 0 else 16) / (t_ms * 1e-3) / 1e12
    mma_peaks = {"tf32_m128_n256": mma_peak(0, 256), "tf32_m128_n128": mma_peak(0, 128),
                 "f16_m128_n256": mma_peak(1, 256), "f16_m128_n128": mma_peak(1, 128)}

    dp_phases = None
    if world > 1 and v._plan.dp is not None:
        dp_phases = {name: dict(zip(("exchanges", "mean_ns"), l._plan.dp.phase_ns())) for name, l in (("critic", v), ("actor", p))}
        dp_phases["phases"] = ["wait for every rank's gradient", "reduce own slice + deliver", "wait for the other slices",
                               "norm + clip + AdamW + Polyak"]
    if rank != 0:
        leave()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    # figures that only a profiler gives (DRAM bytes, tensor-pipe activity), from the committed capture
    ncu = {}
    try:
        ncu = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))
    except OSError:
        pass
    # TF32 dense peak: the larger of half the BURST cuBLAS bf16 figure (the bench runs at the maximum SM
    # clock, the sustained figure was taken at 1282 MHz) and this run's own tcgen05 kind::tf32 issue rate
    tf32_peak = max(peaks.get("bf16_tflops", 1675.7) / 2.0, mma_peaks["tf32_m128_n256"])
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    TC_KERNELS = ("pqlb_gemm_tf32", "pqlb_mlp_forward", "pqlb_mlp_forward_h", "pqlb_mlp_backward")   # every tcgen05 launch of the step
    gemm_ms = sum(per_kernel.get(k, 0.0) for k in TC_KERNELS)
    flops_step = B * (V_PER_STEP * FLOP_V + P_PER_STEP * FLOP_P)
    achieved = flops_step / (gemm_ms * 1e-3) / 1e12
    fwd_mode = v._plan.fwd_mode
    fwd_name = "pqlb_mlp_forward_h" if fwd_mode == "f16x3" else "pqlb_mlp_forward"
    fwd_ms = per_kernel.get(fwd_name, 0.0)
    fwd_flops_step = 2.0 * B * (V_PER_STEP * FWD_MAC_V + P_PER_STEP * FWD_MAC_P)
    # split-fp16: the critic nets issue three kind::f16 MMAs per product, the policy nets one
    fwd_issued_step = 2.0 * B * (V_PER_STEP * (MAC_ACTOR_FWD + 12 * MAC_CRITIC_FWD) + P_PER_STEP * (MAC_ACTOR_FWD + 6 * MAC_CRITIC_FWD)) \
        if fwd_mode == "f16x3" else fwd_flops_step
    fwd_achieved = fwd_flops_step / (max(fwd_ms, 1e-9) * 1e-3) / 1e12
    fwd_issued = fwd_issued_step / (max(fwd_ms, 1e-9) * 1e-3) / 1e12
    # dense peak of the forward's MMA kind: kind::f16 -> the burst cuBLAS bf16 figure (the bench runs at the maximum
    # SM clock with no power cap active; the sustained figure was taken at 1282 MHz), kind::tf32 -> half of it
    dense_peak = peaks.get("bf16_tflops", 1590.0) / (1.0 if fwd_mode == "f16x3" else 2.0)
    n_fwd = len_fwd.get(fwd_name, 0)
    value = world * V_PER_STEP * args.steps / (ms * 1e-3)
    e2e = world * V_PER_STEP * args.steps / (ms_e2e * 1e-3)
    line = {"metric": "critic updates/s (batch 8192)", "value": value, "unit": "critic updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "split-fp16 x3 forward (22 significand bits) + tf32 backward, fp32 accumulate" if fwd_mode == "f16x3" else "tf32 (fp32 accumulate)", "data": "synthetic",
            "config": dict(config_dict(world, args.dp), **({"gradient_exchange": "ncclAllReduce (fused exchange unavailable on this box)"}
                                                  if world > 1 and args.dp == "fused" and v._plan.dp is None else {})),
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "critic updates/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": 2 * 5 * 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "timing": {"blocks": reps, "steps_per_block": args.steps, "ms_per_step_blocks": [x / args.steps for x in blocks_ms],
                       "ms_per_step_median": ms / args.steps, "ms_per_step_min": min(blocks_ms) / args.steps,
                       "value_best_block": world * V_PER_STEP * args.steps / (min(blocks_ms) * 1e-3),
                       "e2e_ms_per_step_blocks": [x / args.steps for x in e2e_blocks], "reported": "median block"},
            "sync_loss": {"value": world * V_PER_STEP * args.steps / (ms_sync * 1e-3), "unit": "critic updates/s",
                          "ms_per_step": ms_sync / args.steps,
                          "note": "cfg.sync_loss = True: update() blocks on the learner's stream and returns the current loss "
                                  "mean, as the reference's update() does; the headline uses the non-blocking read-back "
                                  "(DeviceTracker.mean_lagged: the mean as of the previous env step)"},
            "roofline": {"kernel": f"{'mlp_fwd_h_kernel' if fwd_mode == 'f16x3' else 'mlp_fwd_kernel'} ({fwd_name}: layer-fused Linear+ELU x3 "
                                   f"trunk + Q / policy head, tcgen05 {'kind::f16 on split-fp16 operands, three MMAs per product in the critics' if fwd_mode == 'f16x3' else 'kind::tf32'}) "
                                   "- the dominant kernel of the step",
                         "bound": "tensor", "achieved": fwd_achieved, "peak": dense_peak, "unit": "TFLOP/s",
                         "frac": fwd_achieved / dense_peak,
                         "traffic": ncu.get("fwd_dram_bytes_per_launch"),
                         "traffic_source": ncu.get("source"),
                         "launches_per_step": n_fwd, "algorithmic_flops_per_launch": fwd_flops_step / max(n_fwd, 1),
                         "us_per_launch": 1e3 * fwd_ms / max(n_fwd, 1), "ms_per_step_in_kernel": fwd_ms,
                         "peak_source": f"{peak_src}: burst cuBLAS bf16 figure for kind::f16 (half of it for kind::tf32); the run's own "
                                        "tcgen05 issue rates are listed beside it",
                         "issued_mma": {"tflops": fwd_issued, "frac_of_peak": fwd_issued / dense_peak,
                                        "note": "FLOPs the tensor pipe executes: a critic product is a_hi.w_hi + a_lo.w_hi + a_hi.w_lo "
                                                "(three fp16 MMAs, 22 significand bits; DESIGN.md section 4), a policy product one; "
                                                "'achieved' counts every product once (SURVEY 8d's algorithmic FLOPs)"},
                         "tcgen05_issue_rate_tflops": {k: round(x, 1) for k, x in mma_peaks.items()},
                         "all_tcgen05_kernels": {"kernels": "gemm_tf32_kernel + mlp_fwd(_h)_kernel + mlp_bwd_kernel (every dense-layer "
                                                            "launch of the step: forward, dgrad, wgrad, heads)",
                                                 "achieved": achieved, "frac_of_tf32_peak": achieved / tf32_peak, "launches_per_step": n_gemm,
                                                 "ms_per_step_in_kernel": gemm_ms, "algorithmic_flops_per_step": flops_step},
                         "ncu_tensor_pipe_active_pct": ncu.get("tensor_pipe_active_pct")},
            "kernel_ms_per_step": {k: round(x, 4) for k, x in sorted(per_kernel.items(), key=lambda kv: -kv[1])},
            "replay": {"kernels": replay, "hbm_peak_gbs": hbm_peak, "peak_source": peak_src,
                       "insert_frac_of_hbm": replay["insert_122880"]["gbs"] / hbm_peak,
                       "sample_frac_of_hbm": replay["sample_65536"]["gbs"] / hbm_peak,
                       "bytes_per_unit": {"insert": BYTES_INSERT, "sample": BYTES_SAMPLE},
                       "traffic": ncu.get("replay_traffic")},
            "host_wall_ms_per_step": 1e3 * wall / args.steps,
            "losses": {"critic": float(losses[0]), "actor": float(losses[1])}}
    if dp_phases is not None:
        line["dp_exchange_phases"] = dp_phases
    if sync_check is not None:
        line["params_in_sync"] = sync_check
    if world == 1 and not args.no_sub:
        # the other single-GPU configurations of BASELINE.json, measured in the same run through the full
        # lock-step loop (pql_b200.train.LockStepTrainer on a stub env): sub-records, not the headline
        del v, p
        gc.collect(); torch.cuda.empty_cache()
        from tools.bench_loop import measure_loop
        subs = {}
        for name, kw in (("configs[2]_c51_b16384", dict(envs=E, obs=O, act=A, memory=CAP, batch=16384, distl=True, iters=30, warmup=8, profile=True)),
                         ("configs[3]_shadowhand_loop", dict(envs=16384, obs=211, act=20, memory=5_000_000, batch=8192, iters=30, warmup=8, profile=True)),
                         # the same loop with one TF32 MMA per product everywhere (cfg.forward_mode = "tf32": the round-1 number format)
                         ("configs[3]_shadowhand_loop_tf32_mode", dict(envs=16384, obs=211, act=20, memory=5_000_000, batch=8192, iters=30, warmup=8, forward_mode="tf32")),
                         ("configs[1]_allegro_loop", dict(envs=E, obs=O, act=A, memory=CAP, batch=B, iters=30, warmup=8))):
            try:
                subs[name] = measure_loop(**kw)
            except Exception as e:      # a sub-record must never take the headline down
                subs[name] = {"error": f"{type(e).__name__}: {e}"}
            try:
                gc.collect(); torch.cuda.empty_cache()
            except Exception as e:      # a sticky CUDA error from a sub-record: keep what was measured, stop here
                subs[name + "_cleanup"] = {"error": f"{type(e).__name__}: {e}"}
                break
        line["sub_records"] = subs
        try:
            line["cuda_eager_baseline"] = cuda_eager_rate(dev)
        except Exception as e:
            line["cuda_eager_baseline"] = {"error": f"{type(e).__name__}: {e}"}
    if not args.no_cpu_baseline and world == 1:
        r = cpu_reference_rate(seconds_budget=15.0)
        line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line), flush=True)
    leave()


if __name__ == "__main__":
    main()
