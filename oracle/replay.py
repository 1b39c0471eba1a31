"""numpy restatement of the reference replay path (test infrastructure only).

Follows
  /root/reference/pql/replay/simple_replay.py:4-18   (create_buffer)
  /root/reference/pql/replay/simple_replay.py:40-83  (ReplayBuffer.add_to_buffer)
  /root/reference/pql/replay/simple_replay.py:85-104 (ReplayBuffer.sample_batch)
  /root/reference/pql/replay/nstep_replay.py:29-71   (NStepReplay.add_to_buffer, fifo_shift)
  /root/reference/pql/replay/nstep_replay.py:74-92   (compute_nstep_return)
  /root/reference/pql/algo/pql_p_learner.py:66-85    (P-learner observation ring)

Everything here is bit-exact byte/index work on float32 / bool arrays.
"""
import numpy as np

F32 = np.float32


def ring_advance(next_p, if_full, n, capacity):
    """Pointer bookkeeping of one insert of ``n`` rows (simple_replay.py:52-54,67,82-83).

    Returns (head_rows, tail_rows, new_next_p, new_if_full, new_cur_capacity):
    ``head_rows`` rows go to [next_p, next_p+head_rows); when the insert wraps
    (strict ``>``), the *last* ``tail_rows`` rows of the input go to [0, tail_rows).
    """
    p = next_p + n
    if p > capacity:
        head = capacity - next_p
        p -= capacity
        return head, p, p, True, capacity
    return n, 0, p, if_full, (capacity if if_full else p)


class RingOracle:
    """Transition ring: obs/action/reward/next_obs f32, done stored as bool."""

    def __init__(self, capacity, obs_dim, action_dim):
        self.capacity = int(capacity)
        self.obs_dim = int(obs_dim)
        self.action_dim = int(action_dim)
        self.buf_obs = np.zeros((self.capacity, self.obs_dim), F32)
        self.buf_action = np.zeros((self.capacity, self.action_dim), F32)
        self.buf_reward = np.zeros((self.capacity, 1), F32)
        self.buf_next_obs = np.zeros((self.capacity, self.obs_dim), F32)
        self.buf_done = np.zeros((self.capacity, 1), np.bool_)
        self.next_p = 0
        self.if_full = False
        self.cur_capacity = 0

    def insert(self, obs, actions, rewards, next_obs, dones):
        # simple_replay.py:47-51 - flatten to rows; done -> bool is (x != 0)
        obs = np.asarray(obs, F32).reshape(-1, self.obs_dim)
        actions = np.asarray(actions, F32).reshape(-1, self.action_dim)
        rewards = np.asarray(rewards, F32).reshape(-1, 1)
        next_obs = np.asarray(next_obs, F32).reshape(-1, self.obs_dim)
        dones = (np.asarray(dones).reshape(-1, 1) != 0)
        n = rewards.shape[0]
        head, tail, p, full, cur = ring_advance(self.next_p, self.if_full, n, self.capacity)
        cols = ((self.buf_obs, obs), (self.buf_action, actions), (self.buf_reward, rewards),
                (self.buf_next_obs, next_obs), (self.buf_done, dones))
        for dst, src in cols:
            dst[self.next_p:self.next_p + head] = src[:head]      # :56-61 / :73-80
            if tail:
                dst[0:tail] = src[n - tail:]                      # :64-72, x[-p:]
        self.next_p, self.if_full, self.cur_capacity = p, full, cur

    def gather(self, idx):
        # simple_replay.py:98-104 with the indices supplied by the caller
        idx = np.asarray(idx, np.int64)
        return (self.buf_obs[idx], self.buf_action[idx], self.buf_reward[idx],
                self.buf_next_obs[idx], self.buf_done[idx].astype(F32))


class ObsRingOracle:
    """Observation-only ring owned by the P-learner (pql_p_learner.py:34-37,66-83)."""

    def __init__(self, capacity, obs_dim):
        self.capacity = int(capacity)
        self.obs_dim = int(obs_dim)
        self.memory = np.zeros((self.capacity, self.obs_dim), F32)
        self.next_p = 0
        self.if_full = False
        self.cur_capacity = 0

    def insert(self, obs):
        obs = np.asarray(obs, F32).reshape(-1, self.obs_dim)
        n = obs.shape[0]
        head, tail, p, full, cur = ring_advance(self.next_p, self.if_full, n, self.capacity)
        self.memory[self.next_p:self.next_p + head] = obs[:head]
        if tail:
            self.memory[0:tail] = obs[n - tail:]
        self.next_p, self.if_full, self.cur_capacity = p, full, cur

    def gather(self, idx):
        return self.memory[np.asarray(idx, np.int64)]


def nstep_return(win_next_obs, win_done, win_reward, gammas):
    """compute_nstep_return (nstep_replay.py:74-92) for windows [E, n, *].

    k* = argmax over the window of ``done`` (first maximum); an env "has a done"
    when any window entry is non-zero.  Reward is sum_k (r_k*g_k)*m_k evaluated
    left to right in f32 with each product rounded separately.
    """
    E, n = win_done.shape[0], win_done.shape[1]
    d = win_done.reshape(E, n)
    any_done = (d != 0).any(axis=1)
    kstar = d.argmax(axis=1)
    done_out = d[:, -1].copy()
    done_out[any_done] = F32(1.0)
    pick = np.where(any_done, kstar, n - 1)
    next_obs_out = win_next_obs[np.arange(E), pick]
    mask = np.ones((E, n), F32)
    steps = np.arange(n)[None, :]
    mask[any_done] = (steps <= kstar[any_done][:, None]).astype(F32)
    disc = (win_reward.reshape(E, n).astype(F32) * gammas.reshape(1, n).astype(F32)).astype(F32)
    disc = (disc * mask).astype(F32)
    acc = disc[:, 0].copy()
    for k in range(1, n):
        acc = (acc + disc[:, k]).astype(F32)
    return acc.reshape(E, 1), next_obs_out, done_out.reshape(E, 1)


class NStepOracle:
    """Per-env FIFO window of ``nstep`` transitions (nstep_replay.py:7-71)."""

    def __init__(self, obs_dim, action_dim, num_envs, nstep=3, gamma=0.99):
        self.E, self.n = int(num_envs), int(nstep)
        self.obs_dim, self.action_dim = int(obs_dim), int(action_dim)
        self.w_obs = np.zeros((self.E, self.n, self.obs_dim), F32)
        self.w_act = np.zeros((self.E, self.n, self.action_dim), F32)
        self.w_next = np.zeros((self.E, self.n, self.obs_dim), F32)
        self.w_rew = np.zeros((self.E, self.n, 1), F32)
        self.w_done = np.zeros((self.E, self.n, 1), F32)
        self.count = 0
        # torch.tensor([gamma**i ...]) is a float32 tensor of python doubles (:23)
        self.gammas = np.array([gamma ** i for i in range(self.n)], dtype=np.float64).astype(F32)

    @staticmethod
    def _shift(win, new):
        win[:, :-1] = win[:, 1:].copy()
        win[:, -1] = new

    def push(self, obs, actions, rewards, next_obs, dones):
        """Inputs [E, T, *]; returns the emitted transitions, time-major rows (w*E + e)."""
        obs = np.asarray(obs, F32); actions = np.asarray(actions, F32)
        rewards = np.asarray(rewards, F32); next_obs = np.asarray(next_obs, F32)
        dones = np.asarray(dones, F32)
        if self.n <= 1:                      # :66-67 passthrough
            return obs, actions, rewards, next_obs, dones
        out = ([], [], [], [], [])
        for t in range(obs.shape[1]):
            self._shift(self.w_obs, obs[:, t]); self._shift(self.w_next, next_obs[:, t])
            self._shift(self.w_done, dones[:, t].reshape(self.E, 1))
            self._shift(self.w_act, actions[:, t])
            self._shift(self.w_rew, rewards[:, t].reshape(self.E, 1))
            self.count += 1
            if self.count < self.n:
                continue
            r, no, dn = nstep_return(self.w_next, self.w_done, self.w_rew, self.gammas)
            out[0].append(self.w_obs[:, 0].copy()); out[1].append(self.w_act[:, 0].copy())
            out[2].append(r); out[3].append(no); out[4].append(dn)
        if not out[0]:
            raise ValueError("no n-step transition emitted (the reference fails in torch.cat of an empty list)")
        return tuple(np.concatenate(x, axis=0) for x in out)
