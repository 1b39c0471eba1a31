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
