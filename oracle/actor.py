"""TEST INFRASTRUCTURE - CPU restatement (torch-CPU fp32) of the actor-side env-step path
(SURVEY f1).  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import it.

Follows, line by line:
    RunningMeanStd            pql/utils/torch_util.py:69-114
    add_normal_noise          pql/utils/noise.py:19-27
    add_mixed_normal_noise    pql/utils/noise.py:30-41
    handle_timeout            pql/utils/common.py:195-202
    Tracker                   pql/utils/common.py:103-126
    PQLActor.get_actions      pql/algo/pql_actor.py:70-85
    PQLActor.explore_env      pql/algo/pql_actor.py:87-127
    PQLActor.update_tracker   pql/algo/pql_actor.py:129-136
Pinned against the reference itself: tests/golden/make_golden.py runs the unmodified PQLActor on a
scripted env (tests/golden/inputs.py: ScriptedEnv) and stores its outputs in
tests/golden/actor_small.npz; tests/test_oracle_golden.py compares this file against them.
The random draws (torch.normal / torch.rand) are made here exactly as the reference makes them, so
on the CPU the same torch seed gives the same values; the GPU tests inject the draws instead.
"""
from collections import deque

import numpy as np
import torch

from . import learner as L
from .replay import NStepOracle


class RunningMeanStdOracle:
    """torch_util.py:69-114 (parallel-variance merge of batch moments)."""

    def __init__(self, epsilon=1e-4, shape=()):
        self.mean = torch.zeros(shape)
        self.var = torch.ones(shape)
        self.epsilon = epsilon
        self.count = epsilon

    def update(self, x):
        batch_mean = x.mean(dim=0)
        batch_var = x.var(dim=0)
        batch_count = x.shape[0]
        delta = batch_mean - self.mean
        tot_count = self.count + batch_count
        new_mean = self.mean + delta * batch_count / tot_count
        m_a = self.var * self.count
        m_b = batch_var * batch_count
        m_2 = m_a + m_b + delta ** 2 * self.count * batch_count / tot_count
        self.mean, self.var, self.count = new_mean, m_2 / tot_count, tot_count

    def normalize(self, x):
        return (x - self.mean) / torch.sqrt(self.var + self.epsilon)

    def get_states(self):
        return self.mean, self.var, self.epsilon


def normal_noise(shape, std):
    """noise.py:20-21: torch.normal(zeros, full(std))."""
    return torch.normal(torch.zeros(shape), torch.full(shape, std))


def mixed_noise(shape, std_max, std_min):
    """noise.py:31-35: one std per env, linspace(std_min, std_max, E) down the rows."""
    std_seq = torch.linspace(std_min, std_max, shape[0]).unsqueeze(-1).expand(shape)
    return torch.normal(torch.zeros(shape), std_seq)


def apply_noise(x, noise, noise_bounds=None, out_bounds=(-1.0, 1.0)):
    """noise.py:22-27 / 36-41 on a given draw."""
    if noise_bounds is not None:
        noise = noise.clamp(noise_bounds[0], noise_bounds[1])
    out = x + noise
    if out_bounds is not None:
        out = out.clamp(out_bounds[0], out_bounds[1])
    return out


def handle_timeout(dones, info):
    timeout_envs = info.get('TimeLimit.truncated') if hasattr(info, "get") else None
    if timeout_envs is not None:
        dones = dones * (~timeout_envs)
    return dones


class TrackerOracle:
    def __init__(self, max_len):
        self.moving_average = deque([0 for _ in range(max_len)], maxlen=max_len)

    def update(self, value):
        self.moving_average.extend(value.tolist())

    def mean(self):
        return np.mean(self.moving_average)


class ActorOracle:
    """PQLActor on the CPU with the policy given as a parameter list (oracle.learner.init_mlp layout).
    ``noise_type`` in {'mixed', 'fixed'}; ``draws`` (optional) is a list that receives every noise /
    random-action draw in order, or - with ``replay_draws`` - supplies them."""

    def __init__(self, env, num_envs, obs_dim, act_dim, actor_params, nstep=3, gamma=0.99, obs_norm=True,
                 noise_type="mixed", std_max=0.8, std_min=0.05, reward_scale=0.01, do_handle_timeout=True,
                 tracker_len=100, replay_draws=None):
        self.env, self.E, self.O, self.A = env, num_envs, obs_dim, act_dim
        self.actor = actor_params
        self.obs_norm, self.noise_type = obs_norm, noise_type
        self.std_max, self.std_min = std_max, std_min
        self.reward_scale, self.do_handle_timeout = reward_scale, do_handle_timeout
        self.obs_rms = RunningMeanStdOracle(shape=(obs_dim,)) if obs_norm else None
        self.n_step_buffer = NStepOracle(obs_dim, act_dim, num_envs, nstep, gamma)
        self.return_tracker, self.step_tracker = TrackerOracle(tracker_len), TrackerOracle(tracker_len)
        self.current_returns = torch.zeros(num_envs)
        self.current_lengths = torch.zeros(num_envs)
        self.obs = None
        self.draws = []
        self._replay = list(replay_draws) if replay_draws is not None else None

    def reset_agent(self):
        self.obs = self.env.reset()

    def _draw(self, make):
        d = self._replay.pop(0) if self._replay is not None else make()
        self.draws.append(d)
        return d

    def get_actions(self, obs, sample=True):
        if self.obs_norm:
            obs = self.obs_rms.normalize(obs)
        actions = L.actor_forward(obs, self.actor)
        if sample:
            shape = tuple(actions.shape)
            if self.noise_type == "fixed":
                noise = self._draw(lambda: normal_noise(shape, self.std_max))
            elif self.noise_type == "mixed":
                noise = self._draw(lambda: mixed_noise(shape, self.std_max, self.std_min))
            else:
                raise NotImplementedError
            actions = apply_noise(actions, noise)
        return actions

    def update_tracker(self, reward, done):
        self.current_returns += reward
        self.current_lengths += 1
        env_done_indices = torch.where(done)[0]
        self.return_tracker.update(self.current_returns[env_done_indices])
        self.step_tracker.update(self.current_lengths[env_done_indices])
        self.current_returns[env_done_indices] = 0
        self.current_lengths[env_done_indices] = 0

    @torch.no_grad()
    def explore_env(self, timesteps, random):
        E, O, A = self.E, self.O, self.A
        ts = torch.empty((E, timesteps, O)); ta = torch.empty((E, timesteps, A))
        tr = torch.empty((E, timesteps)); tn = torch.empty((E, timesteps, O)); td = torch.empty((E, timesteps))
        obs = self.obs
        for i in range(timesteps):
            if self.obs_norm:
                self.obs_rms.update(obs)
            if random:
                action = self._draw(lambda: torch.rand((E, A))) * 2.0 - 1.0
            else:
                action = self.get_actions(obs, sample=True)
            next_obs, reward, done, info = self.env.step(action)
            self.update_tracker(reward, done)
            if self.do_handle_timeout:
                done = handle_timeout(done, info)
            ts[:, i] = obs; ta[:, i] = action; td[:, i] = done; tr[:, i] = reward; tn[:, i] = next_obs
            obs = next_obs
        self.obs = obs
        tr = self.reward_scale * tr.reshape(E, timesteps, 1)
        td = td.reshape(E, timesteps, 1)
        out = self.n_step_buffer.push(ts.numpy(), ta.numpy(), tr.numpy(), tn.numpy(), td.numpy())
        obs_o, act_o, rew_o, nxt_o, done_o = (torch.from_numpy(np.asarray(x)) for x in out)
        return obs_o.clone(), (obs_o, act_o, rew_o, nxt_o, done_o), timesteps * E
