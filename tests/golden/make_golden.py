#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the REFERENCE itself.

Runs only in the build container (needs /root/reference, which is absent on the GPU box):
    python tests/golden/make_golden.py
It imports the unmodified reference classes under import stubs for the third-party
packages that are not installed (gym, omegaconf, escnn, morpho_symm, ray),
feeds them the deterministic inputs of ``inputs.py`` and stores only their *outputs*.
Random draws are injected by temporarily replacing ``torch.randint`` / ``torch.normal``
so the reference's own ``sample_batch`` / ``add_normal_noise`` code still runs.
"""
import hashlib
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from tests.golden import inputs  # noqa: E402


# ------------------------------------------------------------------ import stubs (App. C)
class _Meta(type):
    def __getattr__(cls, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Meta(k, (), {})


class _Permissive(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Meta(k, (), {})


for name in ["gym", "gym.spaces", "omegaconf", "omegaconf.dictconfig", "escnn", "escnn.nn",
             "morpho_symm", "morpho_symm.nn", "morpho_symm.nn.EquivariantModules", "ray"]:
    if name not in sys.modules:
        sys.modules[name] = _Permissive(name)
sys.modules["ray"].remote = (lambda *a, **k: a[0] if (len(a) == 1 and callable(a[0]) and not k)
                             else (lambda c: c))

import pql.algo.pql_p_learner as P  # noqa: E402
import pql.algo.pql_v_learner as V  # noqa: E402
from pql.replay.nstep_replay import NStepReplay  # noqa: E402
from pql.replay.simple_replay import ReplayBuffer  # noqa: E402
from pql.utils.distl_util import projection  # noqa: E402


class _TorchCPU:
    """Forwards to torch but pins torch.device(...) to the CPU (the learners hard-code cuda:k)."""

    def __getattr__(self, k):
        return getattr(torch, k)

    @staticmethod
    def device(*a, **k):
        return torch.device("cpu")


V.torch = _TorchCPU()
P.torch = _TorchCPU()


class AD(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def make_cfg(distl, batch, memory):
    return AD(available_gpus=1, artifact=None, num_envs=16, sim_device="cpu", info_track_keys=None,
              algo=AD(v_learner_gpu=0, p_learner_gpu=0, distl=distl, cri_class="DoubleQ",
                      act_class="TanhMLPPolicy", v_min=-10, v_max=10, num_atoms=51,
                      critic_lr=5e-4, actor_lr=5e-4, memory_size=memory, batch_size=batch,
                      obs_norm=True, gamma=0.99, nstep=3, tau=0.05, max_grad_norm=0.5,
                      tracker_len=100, reward_scale=1.0, handle_timeout=True, warm_up=32,
                      horizon_len=1,
                      noise=AD(type="mixed", decay=None, std_max=0.8, std_min=0.05,
                               tgt_pol_std=0.8, tgt_pol_noise_bound=0.2)))


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


# ------------------------------------------------------------------ replay / n-step goldens
def replay_small():
    C, O, A = 50, 5, 2
    rb = ReplayBuffer(C, O, A, device="cpu")
    sizes = [16, 16, 16, 16, 7, 50, 1, 29, 50, 3]
    ptrs, out = [], {}
    for i, n in enumerate(sizes):
        rows = inputs.flat_rows(100 + i, n, O, A)
        rb.add_to_buffer(tuple(t(x) for x in rows))
        ptrs.append([rb.next_p, int(rb.if_full), rb.cur_capacity])
    idx = inputs.indices(7, rb.cur_capacity, 37)
    real = torch.randint
    torch.randint = lambda *a, **k: t(idx)
    try:
        s = rb.sample_batch(37, device="cpu")
    finally:
        torch.randint = real
    out.update(ptrs=np.array(ptrs), sizes=np.array(sizes), idx=idx,
               buf_obs=rb.buf_obs.numpy(), buf_action=rb.buf_action.numpy(),
               buf_reward=rb.buf_reward.numpy(), buf_next_obs=rb.buf_next_obs.numpy(),
               buf_done=rb.buf_done.numpy(),
               s_obs=s[0].numpy(), s_action=s[1].numpy(), s_reward=s[2].numpy(),
               s_next_obs=s[3].numpy(), s_done=s[4].numpy())
    np.savez(os.path.join(HERE, "replay_small.npz"), **out)


def nstep_small():
    out = {}
    for tag, (E, n, O, A, Ts) in {"n3": (8, 3, 5, 2, [1, 1, 1, 5, 32, 1]),
                                  "n5": (6, 5, 4, 3, [7, 1, 9]),
                                  "n1": (4, 1, 3, 2, [2])}.items():
        ns = NStepReplay(O, A, num_envs=E, nstep=n, device="cpu", gamma=0.99)
        for j, T in enumerate(Ts):
            blk = inputs.transition_stream(200 + j, E, T, O, A, p_done=0.3)
            try:
                res = ns.add_to_buffer(*(t(x) for x in blk))
            except (RuntimeError, ValueError):
                out[f"{tag}_push{j}_empty"] = np.array(1)
                continue
            for name, r in zip(("obs", "act", "rew", "next", "done"), res):
                out[f"{tag}_push{j}_{name}"] = r.float().numpy()
    np.savez(os.path.join(HERE, "nstep_small.npz"), **out)


def replay_allegro_stream():
    """AllegroHand-shaped pipeline: n-step push -> ring insert (wrapping, capacity not a
    multiple of E) -> sample; only sha256 digests are stored."""
    E, O, A, C = 256, 88, 16, 10_000
    ns = NStepReplay(O, A, num_envs=E, nstep=3, device="cpu", gamma=0.99)
    rb = ReplayBuffer(C, O, A, device="cpu")
    digests, ptrs = [], []
    for step, T in enumerate([32] + [1] * 40):
        blk = inputs.transition_stream(300 + step, E, T, O, A, p_done=0.02)
        traj = ns.add_to_buffer(*(t(x) for x in blk))
        rb.add_to_buffer(traj)
        ptrs.append([rb.next_p, int(rb.if_full), rb.cur_capacity])
        digests.append(sha(*(x.float().numpy() for x in traj)))
    idx = inputs.indices(11, rb.cur_capacity, 8192)
    real = torch.randint
    torch.randint = lambda *a, **k: t(idx)
    try:
        s = rb.sample_batch(8192, device="cpu")
    finally:
        torch.randint = real
    meta = dict(E=E, O=O, A=A, C=C, ptrs=ptrs, push_digests=digests,
                ring_digest=sha(rb.buf_obs.numpy(), rb.buf_action.numpy(), rb.buf_reward.numpy(),
                                rb.buf_next_obs.numpy(), rb.buf_done.numpy()),
                sample_digest=sha(*(x.numpy() for x in s)))
    with open(os.path.join(HERE, "replay_allegro_stream.json"), "w") as f:
        json.dump(meta, f, indent=1)


def projection_kat():
    g = torch.Generator().manual_seed(5)
    B, N = 96, 51
    dist = torch.softmax(torch.randn(B, N, generator=g) * 2, dim=1)
    reward = torch.randn(B, 1, generator=g) * 4
    done = (torch.rand(B, 1, generator=g) < 0.3).float()
    # edge rows: exact atom hits (l == u), both clamps, terminal
    reward[0], done[0] = 0.0, 1.0        # all mass on the atom at 0 (b integral)
    reward[1], done[1] = 25.0, 0.0       # clamp to v_max
    reward[2], done[2] = -25.0, 0.0      # clamp to v_min
    reward[3], done[3] = -10.0, 1.0      # b == 0 exactly (u == 0 branch)
    reward[4], done[4] = 10.0, 1.0       # b == N-1 exactly
    reward[5], done[5] = 0.4, 1.0        # atom spacing multiple
    support = torch.linspace(-10, 10, N)
    out = projection(dist, reward, done, 0.99 ** 3, -10, 10, N, support=support, device="cpu")
    np.savez(os.path.join(HERE, "projection_kat.npz"), dist=dist.numpy(), reward=reward.numpy(),
             done=done.numpy(), out=out.numpy())


# ------------------------------------------------------------------ learner goldens
def load_mlp(module_net, params):
    sd = {}
    for k, (w, b) in zip((0, 2, 4, 6), params):
        sd[f"net.{k}.weight"], sd[f"net.{k}.bias"] = w.clone(), b.clone()
    module_net.load_state_dict(sd)


def tensor_digest(x):
    x = x.detach().float().reshape(-1)
    return np.array([x.sum().item(), x.abs().sum().item(), x.norm().item(),
                     *x[:6].tolist(), *x[-6:].tolist()], dtype=np.float64)


def learner_case(tag, seed, B, O, A, distl, steps=3):
    case = inputs.learner_case(seed, B, O, A, distl)
    cfg = make_cfg(distl, B, B)
    v = V.PQLVLearner(O, A, cfg)
    p = P.PQLPLearner(O, A, cfg)
    load_mlp(v.critic.net_q1, case["q1"]); load_mlp(v.critic.net_q2, case["q2"])
    v.critic_target.load_state_dict(v.critic.state_dict())
    load_mlp(p.actor, case["actor"])
    obs, action, reward, next_obs, done = case["batch"]
    # the replay holds exactly the batch; indices are the identity permutation reversed
    v.memory.add_to_buffer((obs, action, reward, next_obs, done))
    idx = torch.arange(B - 1, -1, -1)
    import copy
    v.actor = copy.deepcopy(p.actor)
    v.normalize_tuple = case["norm"]
    rec = {"idx": idx.numpy()}
    grads_log = []
    real_clip = V.clip_grad_norm_

    def spy_clip(parameters, max_norm):
        params = list(parameters)
        grads_log.append(np.array([q.grad.norm().item() for q in params]))
        return real_clip(parameters=params, max_norm=max_norm)

    real_randint, real_normal = torch.randint, torch.normal
    V.clip_grad_norm_ = spy_clip
    try:
        for s in range(steps):
            torch.randint = lambda *a, **k: idx.clone()
            torch.normal = lambda mean, std, s=s: case["noises"][s].clone()
            v.learn()
    finally:
        torch.randint, torch.normal = real_randint, real_normal
        V.clip_grad_norm_ = real_clip
    rec["v_losses"] = np.array(list(v.loss_tracker.moving_average)[-steps:])
    rec["v_grad_norms"] = np.stack(grads_log)
    for name, q in v.critic.state_dict().items():
        rec[f"critic.{name}"] = tensor_digest(q)
    for name, q in v.critic_target.state_dict().items():
        rec[f"target.{name}"] = tensor_digest(q)
    # P-learner: frozen copy of the *initial* critic, obs ring == batch obs
    pcrit = type(v.critic)(O, A, **(dict(v_min=-10, v_max=10, num_atoms=51, device="cpu") if distl else {}))
    load_mlp(pcrit.net_q1, case["q1"]); load_mlp(pcrit.net_q2, case["q2"])
    p.update(pcrit, obs, case["norm"], 0)
    grads_log.clear()
    P.clip_grad_norm_ = spy_clip
    try:
        for s in range(steps):
            torch.randint = lambda *a, **k: idx.clone()
            p.learn()
    finally:
        torch.randint = real_randint
        P.clip_grad_norm_ = real_clip
    rec["p_losses"] = np.array(list(p.loss_tracker.moving_average)[-steps:])
    rec["p_grad_norms"] = np.stack(grads_log)
    for name, q in p.actor.state_dict().items():
        rec[f"actor.{name}"] = tensor_digest(q)
    np.savez(os.path.join(HERE, f"learner_{tag}.npz"), **rec)


def actor_small():
    """The unmodified PQLActor (pql/algo/pql_actor.py) on the scripted env: warm-up with random
    actions, then policy steps with mixed exploration noise; everything it returns or tracks."""
    import pql.algo.pql_actor as ACT
    from pql.models.mlp import TanhMLPPolicy
    ACT.torch = _TorchCPU()
    for noise_type, obs_norm, timeout in (("mixed", True, True), ("fixed", True, True), ("mixed", False, False)):
        c = inputs.ACTOR_CASE
        E, O, A = c["E"], c["O"], c["A"]
        env = inputs.ScriptedEnv(c["seed"], E, O, A, c["warm_up"] + sum(c["calls"]))
        cfg = make_cfg(False, 64, 1000)
        cfg.num_envs = E
        cfg.algo.tracker_len = c["tracker_len"]
        cfg.algo.reward_scale = 0.01
        cfg.algo.noise.type = noise_type
        cfg.algo.obs_norm, cfg.algo.handle_timeout = obs_norm, timeout
        actor = ACT.PQLActor(env, cfg)
        pol = TanhMLPPolicy(O, A)
        load_mlp(pol, inputs.actor_case_params(c["seed"], O, A))
        actor.actor = pol
        torch.manual_seed(c["seed"])          # after the module constructors, which consume the generator
        actor.reset_agent()
        rec = {}

        def record(tag, res):
            p_data, v_data, steps = res
            rec[f"{tag}_p"] = p_data.numpy()
            for name, x in zip(("obs", "act", "rew", "next", "done"), v_data):
                rec[f"{tag}_{name}"] = x.float().numpy()
            rec[f"{tag}_steps"] = np.array(steps)
            if actor.obs_rms is not None:
                rec[f"{tag}_rms_mean"] = actor.obs_rms.mean.numpy().copy()
                rec[f"{tag}_rms_var"] = actor.obs_rms.var.numpy().copy()
                rec[f"{tag}_rms_count"] = np.array(actor.obs_rms.count, dtype=np.float64)
            rec[f"{tag}_ret_window"] = np.array(list(actor.return_tracker.moving_average), dtype=np.float64)
            rec[f"{tag}_len_window"] = np.array(list(actor.step_tracker.moving_average), dtype=np.float64)
            rec[f"{tag}_returns"] = actor.current_returns.numpy().copy()
            rec[f"{tag}_lengths"] = actor.current_lengths.numpy().copy()

        record("warm", actor.explore_env(env, c["warm_up"], random=True))
        for j, T in enumerate(c["calls"]):
            record(f"call{j}", actor.explore_env(env, T, random=False))
        tag = noise_type if obs_norm else f"{noise_type}_raw"       # _raw: no normaliser, no timeout handling
        np.savez(os.path.join(HERE, f"actor_small_{tag}.npz"), **rec)


def schedules():
    """pql/utils/schedule_util.py: the value sequences of the two exploration-noise schedules."""
    from pql.utils.schedule_util import ExponentialSchedule, LinearSchedule
    out = {}
    for name, sch in (("linear_0.8_0.05_7", LinearSchedule(0.8, 0.05, 7)), ("linear_1_0_3", LinearSchedule(1.0, 0.0, 3)),
                      ("exp_0.8_0.9_0.05", ExponentialSchedule(0.8, 0.9, 0.05)), ("exp_0.5_0.5_none", ExponentialSchedule(0.5, 0.5))):
        vals = [sch.val()]
        for _ in range(40):
            vals.append(sch.step())
            vals.append(sch.val())
        out[name] = vals
    json.dump(out, open(os.path.join(HERE, "schedules.json"), "w"))


def reference_checkpoint():
    """A checkpoint as the reference writes it (pql/utils/model_util.py:24-41 from evaluator.py:112-119):
    {'obs_rms': (mean, var, eps), 'actor': state_dict, 'critic': state_dict} of the reference's OWN modules,
    default-initialised under a fixed seed, plus what those modules compute on a small input.  The weights
    themselves are not stored (2 MB): torch's seeded nn.Linear initialisation is reproducible, so the fixture
    keeps a sha256 of every tensor - the test rebuilds the state_dict from the seed, proves it is the
    reference's bit for bit, writes the checkpoint file and loads it into the pql_b200 modules."""
    from pql.models.mlp import DistributionalDoubleQ, DoubleQ, TanhMLPPolicy
    O, A, n, seed = 24, 4, 96, 2024
    torch.manual_seed(seed)
    actor, critic = TanhMLPPolicy(O, A), DoubleQ(O, A)
    critic_d = DistributionalDoubleQ(O, A, v_min=-10, v_max=10, num_atoms=51, device="cpu")
    g = torch.Generator().manual_seed(seed + 1)
    obs, act = torch.randn(n, O, generator=g), torch.rand(n, A, generator=g) * 2 - 1
    rms = (torch.randn(O, generator=g) * 0.3, torch.rand(O, generator=g) + 0.5, 1e-4)
    with torch.no_grad():
        q1, q2 = critic.get_q1_q2(obs, act)
        p1, p2 = critic_d.get_q1_q2(obs, act)
        out = dict(action=actor(obs), q1=q1, q2=q2, q_min=critic.get_q_min(obs, act), p1=p1, p2=p2,
                   qd_min=critic_d.get_q_min(obs, act))
    digests = {name: {k: hashlib.sha256(v.detach().numpy().tobytes()).hexdigest() for k, v in m.state_dict().items()}
               for name, m in (("actor", actor), ("critic", critic), ("critic_c51", critic_d))}
    np.savez(os.path.join(HERE, "ref_checkpoint.npz"), obs=obs.numpy(), act=act.numpy(), rms_mean=rms[0].numpy(),
             rms_var=rms[1].numpy(), **{k: v.numpy() for k, v in out.items()})
    json.dump(dict(seed=seed, obs_dim=O, act_dim=A, order=["actor", "critic", "critic_c51"], digests=digests),
              open(os.path.join(HERE, "ref_checkpoint.json"), "w"), indent=1)


if __name__ == "__main__":
    torch.set_num_threads(8)
    if "--checkpoint-only" in sys.argv:
        reference_checkpoint()
        sys.exit(0)
    reference_checkpoint()
    schedules()
    actor_small()
    replay_small()
    nstep_small()
    replay_allegro_stream()
    projection_kat()
    learner_case("doubleq", seed=1234, B=512, O=88, A=16, distl=False)
    learner_case("c51", seed=4321, B=256, O=88, A=16, distl=True)
    learner_case("shadow", seed=77, B=128, O=211, A=20, distl=False, steps=2)
    print("golden fixtures written to", HERE)
