"""BASELINE.json configs[3]: the full actor / P-learner / V-learner loop with a synthetic vectorised
env stub - 16384 envs, ShadowHand shape (obs 211, act 20), 5M-slot replay, batch 8192, the reference's
8 : 4 : 1 schedule - through pql_b200.train.LockStepTrainer.  Prints one JSON line:
env steps/s (transitions/s), critic updates/s, ms per loop iteration.

    python tools/bench_loop.py [--envs 16384] [--obs 211] [--act 20] [--memory 5000000] [--iters 50]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


class _Space:
    def __init__(self, shape):
        self.shape = shape


class StubEnv:
    """reset()/step(a) -> (obs, reward, done, info) on the GPU: next_obs = 0.9 * obs + noise block,
    reward = -|a|^2 mean, 1 % random episode ends; a handful of launches, no physics."""

    def __init__(self, E, O, A, device):
        self.E, self.O, self.A, self.dev = E, O, A, device
        g = torch.Generator(device=device).manual_seed(0)
        self.noise = torch.randn(64, E, O, device=device, generator=g)
        self.dones = (torch.rand(64, E, device=device, generator=g) < 0.01).float()
        self.observation_space, self.action_space = _Space((O,)), _Space((A,))
        self.t = 0

    def reset(self):
        self.obs = self.noise[0].clone()
        return self.obs

    def step(self, action):
        self.t += 1
        self.obs = 0.9 * self.obs + self.noise[self.t % 64]
        return self.obs, -(action * action).mean(dim=1), self.dones[self.t % 64], {}


def measure_loop(envs=16384, obs=211, act=20, memory=5_000_000, batch=8192, iters=50, warmup=10, distl=False,
                 profile=False, device="cuda:0", forward_mode=None):
    """Time ``iters`` iterations of LockStepTrainer.step() on the stub env; returns the result dict."""
    from pql_b200.train import LockStepTrainer
    from pql_b200.utils import default_pql_cfg
    dev = torch.device(device)
    torch.manual_seed(42)
    cfg = default_pql_cfg(num_envs=envs, sim_device=str(dev), batch_size=batch, memory_size=memory, distl=distl,
                          v_learner_gpu=dev.index or 0, p_learner_gpu=dev.index or 0)
    cfg.learner_streams = True
    if forward_mode is not None:
        cfg.forward_mode = forward_mode          # "tf32": one TF32 MMA per product everywhere (the round-1 number format)
    tr = LockStepTrainer(StubEnv(envs, obs, act, dev), cfg)
    tr.warm_up()
    for _ in range(warmup):
        tr.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(iters):
        info = tr.step()
    for l in (tr.v_learner, tr.p_learner):
        if l.stream is not None:
            torch.cuda.current_stream().wait_stream(l.stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    wall = (time.perf_counter() - t0) / iters * 1e3
    per_kernel = None
    if profile:
        from pql_b200 import _kernels as K
        tr.v_learner.disable_graph(); tr.p_learner.disable_graph()
        K.PROFILE = {}
        for _ in range(3):
            tr.step()
        torch.cuda.synchronize()
        per_kernel = {k: [round(sum(a.elapsed_time(b) for a, b in ev) / 3, 4), len(ev) // 3]
                      for k, ev in sorted(K.PROFILE.items(), key=lambda kv: -sum(a.elapsed_time(b) for a, b in kv[1]))}
        K.PROFILE = None
    return {"workload": f"full loop, {envs} envs, obs {obs}, act {act}, {memory}-slot replay, batch {batch}, "
                        f"{'C51' if distl else 'twin-Q'}, {tr.v_per_step} critic : {tr.v_per_step // tr.p_every} actor : 1 env step",
            "ms_per_iteration": ms, "host_wall_ms_per_iteration": wall, "iterations": iters,
            "env_transitions_per_s": envs / (ms * 1e-3),
            "critic_updates_per_s": tr.v_per_step / (ms * 1e-3),
            "actor_updates_per_s": tr.v_per_step / tr.p_every / (ms * 1e-3),
            "replay_gb": tr.v_learner.memory.ring.numel() * 4 / 1e9,
            "forward_mode": tr.v_learner._plan.fwd_mode,
            "losses": {"critic": info["train/critic_loss"], "actor": info["train/actor_loss"]},
            "kernel_ms_and_launches_per_iteration": per_kernel}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=16384)
    ap.add_argument("--obs", type=int, default=211)
    ap.add_argument("--act", type=int, default=20)
    ap.add_argument("--memory", type=int, default=5_000_000)
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--distl", action="store_true")
    ap.add_argument("--profile", action="store_true", help="per-kernel device time (CUDA events around every launch, graphs off)")
    ap.add_argument("--fwd-mode", default=None, choices=["f16x3", "tf32"], help="cfg.forward_mode (default: the library's choice)")
    args = ap.parse_args()
    print(json.dumps(measure_loop(args.envs, args.obs, args.act, args.memory, args.batch, args.iters, args.warmup,
                                  args.distl, args.profile, forward_mode=args.fwd_mode)))


if __name__ == "__main__":
    main()
