"""GPU ring-buffer replay: drop-in for pql/replay/simple_replay.py.

Storage is one array-of-records tensor ``ring[capacity, rec_ld]`` (layout in
include/pqlb200.h) instead of the reference's five column tensors: a uniformly sampled
transition is then one contiguous, sector-aligned 800-byte read (AllegroHand) instead of
five scattered ones.  ``buf_obs`` ... ``buf_done`` remain available as views.
"""
import torch

from .. import _lib


def _geom(obs_dim, action_dim):
    lib = _lib.load()
    obs_pad = lib.pqlb_obs_pad(obs_dim)
    act_pad = (action_dim + 3) // 4 * 4
    return dict(obs_pad=obs_pad, act_pad=act_pad, rec_ld=lib.pqlb_record_ld(obs_dim, action_dim),
                off_obs=0, off_next=obs_pad, off_act=2 * obs_pad, off_rew=2 * obs_pad + act_pad,
                off_done=2 * obs_pad + act_pad + 1)


def create_buffer(capacity, obs_dim, action_dim, device='cuda', reserve_space=False):
    """simple_replay.py:4-18.  Returns the record storage; shape (*capacity, rec_ld)."""
    if reserve_space:
        raise NotImplementedError("reserve_space (fp16 host-side obs) is a fork feature PQL never enables")
    if isinstance(capacity, int):
        capacity = (capacity,)
    if not isinstance(obs_dim, int):
        if len(obs_dim) != 1:
            raise NotImplementedError("only flat observations are on the PQL path")
        obs_dim = int(obs_dim[0])
    g = _geom(int(obs_dim), int(action_dim))
    return torch.empty((*capacity, g["rec_ld"]), dtype=torch.float32, device=device)


def _as_dev_f32(x, device, shape):
    x = x.reshape(shape)
    if x.device != device or x.dtype != torch.float32:
        x = x.to(device=device, dtype=torch.float32, non_blocking=True)
    return x.contiguous()


class ReplayBuffer:
    """simple_replay.py:21-104 with the same attributes (next_p, if_full, cur_capacity, capacity)."""

    def __init__(self, capacity: int, obs_dim: int, action_dim: int, device='cpu',
                 left_agent: bool = False, reserve_space: bool = False):
        if left_agent or reserve_space:
            raise NotImplementedError("left_agent / reserve_space are fork features PQL never enables")
        self.obs_dim = (obs_dim,) if isinstance(obs_dim, int) else tuple(obs_dim)
        if len(self.obs_dim) != 1:
            raise NotImplementedError("only flat observations are on the PQL path")
        self.action_dim = int(action_dim)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("pql_b200.ReplayBuffer lives in GPU memory (device must be a CUDA device); "
                               "there is no CPU path")
        self.next_p = 0
        self.if_full = False
        self.cur_capacity = 0
        self.capacity = int(capacity)
        self._O = int(self.obs_dim[0])
        self._g = _geom(self._O, self.action_dim)
        self.ring = create_buffer(self.capacity, self._O, self.action_dim, device=self.device)
        # device-resident copy of cur_capacity: the range of the index draw when the sampler RNG is
        # fused into the gather kernel (a CUDA-graph replay cannot take it as a launch argument)
        self.cur_capacity_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.left_agent = left_agent
        self.reserve_space = reserve_space

    # ---- column views (the reference's buf_* tensors) -----------------------------------
    def _col(self, off, width):
        return self.ring[:, off:off + width]

    @property
    def buf_obs(self):
        return self._col(self._g["off_obs"], self._O)

    @property
    def buf_next_obs(self):
        return self._col(self._g["off_next"], self._O)

    @property
    def buf_action(self):
        return self._col(self._g["off_act"], self.action_dim)

    @property
    def buf_reward(self):
        return self._col(self._g["off_rew"], 1)

    @property
    def buf_done(self):
        return self._col(self._g["off_done"], 1) != 0

    # ---- insert ----------------------------------------------------------------------------
    @torch.no_grad()
    def add_to_buffer(self, trajectory):
        """simple_replay.py:40-83: wrap-aware insert; one kernel launch (K1)."""
        obs, actions, rewards, next_obs, dones = trajectory
        dev = self.device
        obs = _as_dev_f32(obs, dev, (-1, self._O))
        actions = _as_dev_f32(actions, dev, (-1, self.action_dim))
        rewards = _as_dev_f32(rewards, dev, (-1,))
        next_obs = _as_dev_f32(next_obs, dev, (-1, self._O))
        dones = _as_dev_f32(dones, dev, (-1,))
        n = rewards.shape[0]
        if not (obs.shape[0] == actions.shape[0] == next_obs.shape[0] == dones.shape[0] == n):
            raise RuntimeError("add_to_buffer: the five trajectory tensors disagree on the number of rows")
        p = self.next_p + n
        if p > self.capacity and p - self.capacity > self.capacity:
            raise RuntimeError(f"add_to_buffer: {n} rows do not fit a ring of capacity {self.capacity} "
                               f"at pointer {self.next_p} (the reference fails in the slice assignment)")
        with torch.cuda.device(dev):
            _lib.call("pqlb_ring_insert", _lib.ptr(self.ring), self.capacity, self._O, self.action_dim,
                      _lib.ptr(obs), _lib.ptr(actions), _lib.ptr(rewards), _lib.ptr(next_obs),
                      _lib.ptr(dones), n, self.next_p)
        if p > self.capacity:              # strict: next_p may rest at capacity (:54)
            self.if_full = True
            p -= self.capacity
        self.next_p = p
        self.cur_capacity = self.capacity if self.if_full else self.next_p
        with torch.cuda.device(dev):
            _lib.call("pqlb_store_i64", _lib.ptr(self.cur_capacity_dev), self.cur_capacity)

    # ---- sample ----------------------------------------------------------------------------
    @torch.no_grad()
    def sample_indices(self, batch_size):
        # torch.randint keeps the reference's Philox stream for a given seed (simple_replay.py:87)
        return torch.randint(self.cur_capacity, size=(batch_size,), device=self.device)

    @torch.no_grad()
    def gather(self, indices):
        """The five ``buf[indices]`` gathers (+ done.float()) in one launch (K2)."""
        B = indices.shape[0]
        dev = self.device
        out = (torch.empty((B, self._O), dtype=torch.float32, device=dev),
               torch.empty((B, self.action_dim), dtype=torch.float32, device=dev),
               torch.empty((B, 1), dtype=torch.float32, device=dev),
               torch.empty((B, self._O), dtype=torch.float32, device=dev),
               torch.empty((B, 1), dtype=torch.float32, device=dev))
        with torch.cuda.device(dev):
            _lib.call("pqlb_sample_gather", _lib.ptr(self.ring), self.capacity, self._O, self.action_dim,
                      _lib.ptr(indices), B, *(_lib.ptr(t) for t in out))
        return out

    @torch.no_grad()
    def sample_batch(self, batch_size, device='cuda'):
        """simple_replay.py:85-104."""
        out = self.gather(self.sample_indices(batch_size))
        tgt = torch.device(device)
        if tgt.type == "cuda" and tgt.index is None:
            return out
        return tuple(t.to(tgt) for t in out)
