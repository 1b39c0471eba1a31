"""n-step return assembly on the GPU: drop-in for pql/replay/nstep_replay.py."""
import ctypes as C

import torch

from .. import _lib
from .simple_replay import _as_dev_f32, create_buffer


class NStepReplay:
    """nstep_replay.py:6-71.  The per-env window of ``nstep`` transitions is a circular array of
    records; one launch (K1') shifts it, evaluates compute_nstep_return (:74-92) for every
    emitting step and writes the emitted transitions time-major."""

    def __init__(self, obs_dim: int, action_dim: int, num_envs: int = 1, nstep: int = 3,
                 device: str = 'cuda', gamma: float = 0.99, left_agent: bool = False):
        if left_agent:
            raise NotImplementedError("left_agent is a fork feature PQL never enables")
        if not isinstance(obs_dim, int):
            if len(obs_dim) != 1:
                raise NotImplementedError("only flat observations are on the PQL path")
            obs_dim = int(obs_dim[0])
        self.obs_dim, self.action_dim = int(obs_dim), int(action_dim)
        self.num_envs, self.nstep = int(num_envs), int(nstep)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("pql_b200.NStepReplay lives in GPU memory; there is no CPU path")
        if self.nstep > 32:
            raise NotImplementedError("nstep > 32")
        self.gamma = gamma
        self.nstep_count = 0
        self.left_agent = left_agent
        if self.nstep > 1:
            self.window = create_buffer((self.num_envs, self.nstep), self.obs_dim, self.action_dim,
                                        device=self.device)
        # float32 of the python doubles gamma**i, like torch.tensor([...]) (:23)
        self._gammas = (C.c_float * 32)(*[float(self.gamma ** i) for i in range(self.nstep)])
        self.gamma_array = torch.tensor([self.gamma ** i for i in range(self.nstep)],
                                        device=self.device).view(-1, 1)

    def _push(self, obs, actions, rewards, next_obs, dones, T):
        E, n = self.num_envs, self.nstep
        t0 = min(T, max(0, n - 1 - self.nstep_count))
        rows = (T - t0) * E
        dev = self.device
        out = (torch.empty((rows, self.obs_dim), dtype=torch.float32, device=dev),
               torch.empty((rows, self.action_dim), dtype=torch.float32, device=dev),
               torch.empty((rows, 1), dtype=torch.float32, device=dev),
               torch.empty((rows, self.obs_dim), dtype=torch.float32, device=dev),
               torch.empty((rows, 1), dtype=torch.float32, device=dev))
        with torch.cuda.device(dev):
            _lib.call("pqlb_nstep_push", _lib.ptr(self.window), E, n, self.obs_dim, self.action_dim,
                      _lib.ptr(obs), _lib.ptr(actions), _lib.ptr(rewards), _lib.ptr(next_obs),
                      _lib.ptr(dones), T, self.nstep_count, self._gammas,
                      *(_lib.ptr(t) for t in out))
        self.nstep_count += T
        return out if rows else None

    @torch.no_grad()
    def add_to_buffer(self, obs, actions, rewards, next_obs, dones, reward_left=None):
        if self.nstep <= 1:                        # :66-67 passthrough
            return obs, actions, rewards, next_obs, dones
        E = self.num_envs
        T = obs.shape[1]
        dev = self.device
        obs = _as_dev_f32(obs, dev, (E, T, self.obs_dim))
        actions = _as_dev_f32(actions, dev, (E, T, self.action_dim))
        rewards = _as_dev_f32(rewards, dev, (E, T))
        next_obs = _as_dev_f32(next_obs, dev, (E, T, self.obs_dim))
        dones = _as_dev_f32(dones, dev, (E, T))
        res = self._push(obs, actions, rewards, next_obs, dones, T)
        if res is None:
            raise ValueError("NStepReplay.add_to_buffer: no n-step transition emitted yet "
                             "(the reference fails in torch.cat(): expected a non-empty list of Tensors)")
        return res
