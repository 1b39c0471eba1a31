"""Deterministic input generators shared by make_golden.py (reference side, build container
only) and the parity tests (oracle / CUDA side).  numpy RandomState and torch's CPU
generator are both platform-stable, so the inputs never need to be stored."""
import numpy as np
import torch

F32 = np.float32


def transition_stream(seed, E, T, obs_dim, act_dim, p_done=0.05, reward_scale=0.01):
    """[E, T, *] float32 blocks like PQLActor.explore_env produces (pql_actor.py:111-120)."""
    rs = np.random.RandomState(seed)
    obs = rs.standard_normal((E, T, obs_dim)).astype(F32)
    nxt = rs.standard_normal((E, T, obs_dim)).astype(F32)
    act = rs.uniform(-1, 1, (E, T, act_dim)).astype(F32)
    rew = (rs.standard_normal((E, T, 1)) * reward_scale).astype(F32)
    done = (rs.uniform(0, 1, (E, T, 1)) < p_done).astype(F32)
    return obs, act, rew, nxt, done


def flat_rows(seed, n, obs_dim, act_dim, p_done=0.1):
    """n flat transitions (what ReplayBuffer.add_to_buffer receives)."""
    o, a, r, no, d = transition_stream(seed, n, 1, obs_dim, act_dim, p_done)
    return tuple(x.reshape(n, -1) for x in (o, a, r, no, d))


def indices(seed, high, n):
    return np.random.RandomState(seed).randint(0, high, size=n).astype(np.int64)


def learner_case(seed, B, obs_dim, act_dim, distl=False, num_atoms=51):
    """Weights (nn.Linear-style init from a seeded CPU generator), a batch, target-policy
    noise and observation-normaliser statistics for one learner parity case."""
    from oracle.learner import init_mlp
    g = torch.Generator().manual_seed(seed)
    out_dim = num_atoms if distl else 1
    q1 = init_mlp(obs_dim + act_dim, out_dim, g)
    q2 = init_mlp(obs_dim + act_dim, out_dim, g)
    actor = init_mlp(obs_dim, act_dim, g)
    obs = torch.randn(B, obs_dim, generator=g) * 1.5 + 0.3
    next_obs = torch.randn(B, obs_dim, generator=g) * 1.5 + 0.3
    action = torch.rand(B, act_dim, generator=g) * 2 - 1
    reward = torch.randn(B, 1, generator=g) * (1.0 if distl else 0.05)
    done = (torch.rand(B, 1, generator=g) < 0.1).float()
    mean = torch.randn(obs_dim, generator=g) * 0.2
    var = torch.rand(obs_dim, generator=g) * 2 + 0.5
    noises = [torch.randn(B, act_dim, generator=g) * 0.8 for _ in range(8)]
    return dict(q1=q1, q2=q2, actor=actor, batch=(obs, action, reward, next_obs, done),
                norm=(mean, var, 1e-4), noises=noises)


class _Space:
    def __init__(self, shape):
        self.shape = shape


class ScriptedEnv:
    """Deterministic vectorised env stub (``reset()`` / ``step(a) -> obs, reward, done, info``) for the
    actor-path parity cases: observations, base rewards, dones and time-limit flags follow a
    pre-generated script that does not depend on the actions - so two implementations whose actions
    differ by rounding stay on the same trajectory - while the reward adds ``0.1 * sum(action)`` so the
    actions still flow into the stored transitions."""

    def __init__(self, seed, E, O, A, T, device="cpu", p_done=0.08, p_trunc=0.5):
        rs = np.random.RandomState(seed)
        self.obs_seq = torch.from_numpy((rs.standard_normal((T + 1, E, O)) * 1.3 + 0.4).astype(F32)).to(device)
        self.rew_seq = torch.from_numpy(rs.standard_normal((T, E)).astype(F32)).to(device)
        done = rs.uniform(size=(T, E)) < p_done
        trunc = done & (rs.uniform(size=(T, E)) < p_trunc)
        self.done_seq = torch.from_numpy(done.astype(F32)).to(device)
        self.trunc_seq = torch.from_numpy(trunc).to(device)
        self.observation_space, self.action_space = _Space((O,)), _Space((A,))
        self.t = 0

    def reset(self):
        self.t = 0
        return self.obs_seq[0]

    def step(self, action):
        t = self.t
        self.t += 1
        reward = self.rew_seq[t] + 0.1 * action.sum(dim=1)
        return self.obs_seq[t + 1], reward, self.done_seq[t], {"TimeLimit.truncated": self.trunc_seq[t]}


ACTOR_CASE = dict(seed=31, E=24, O=9, A=3, warm_up=5, calls=[1, 1, 1, 3, 1, 1], nstep=3, tracker_len=7)


def actor_case_params(seed, O, A):
    from oracle.learner import init_mlp
    return init_mlp(O, A, torch.Generator().manual_seed(seed))
