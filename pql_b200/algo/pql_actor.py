"""Actor-side env-step path: drop-in for pql/algo/pql_actor.py:11-150.

Same constructor, attributes (``actor``, ``obs``, ``obs_rms``, ``n_step_buffer``, ``return_tracker``,
``step_tracker``, ``current_returns``, ``current_lengths``) and methods (``reset_agent``,
``get_noise_std``, ``update_noise``, ``get_actions``, ``explore_env``, ``update_tracker``,
``add_info_tracker_log``).  Per env step the reference issues ~60 small torch launches around
``env.step``; here it is five kernels: RunningMeanStd.update (pqlb_rms_update), normalise + operand
packing + the exploration-noise draw (pqlb_actor_inputs), the layer-fused policy forward with the
tanh / noise / clamp head (pqlb_mlp_forward), trackers + timeout handling + reward scaling
(pqlb_env_post) and the n-step push (pqlb_nstep_push).  The policy writes the action straight into
the trajectory tensor; with ``horizon_len == 1`` nothing is staged or copied.

The exploration noise is torch.normal(zeros, std) of pql/utils/noise.py:19-41 drawn inside
pqlb_actor_inputs from this actor's ``generator`` (seed, offset) - bit-identical to the torch call
with that generator (csrc/rng.cuh).  The warm-up's uniform random actions (``random=True``,
pql_actor.py:101-103) are a ``torch.rand`` draw: once per run, not on the hot path.
"""
import torch

from .. import _kernels as K
from .. import _lib
from ..models.mlp import FUSED_H_MAX_IN, FUSED_MAX_IN, HIDDEN, NetAddrs, NetLayout, _ru, forward_calls, half_arena
from ..replay.nstep_replay import NStepReplay
from ..utils.common import DeviceTracker
from ..utils.schedule_util import ExponentialSchedule, LinearSchedule
from ..utils.torch_util import RunningMeanStd
from ._engine import _operand_copies
from .pql_v_learner import _PRODUCER_STREAM, module_flat, note_read


class EpisodeTracker(DeviceTracker):
    """Tracker(tracker_len) (pql/utils/common.py:103-126) whose deque is a ring on the GPU:
    pqlb_env_post appends the finished episodes in env order; ``mean()`` reads it back (logging)."""

    def __init__(self, max_len, device):
        super().__init__(max_len, device)

    def update(self, value):
        raise NotImplementedError("the window is filled by pqlb_env_post")


class PQLActor:
    def __init__(self, env, cfg):
        self.env = env
        self.cfg = cfg
        self.obs_dim = self.env.observation_space.shape
        self.action_dim = self.env.action_space.shape[0]
        self.sim_device = torch.device(f"{cfg.sim_device}")
        if self.sim_device.type != "cuda":
            raise RuntimeError("PQLActor needs a CUDA sim_device: pql_b200 has no CPU path")
        if self.sim_device.index is None:
            self.sim_device = torch.device("cuda", torch.cuda.current_device())
        _lib.load()
        self.v_learner_device = torch.device(f"cuda:{cfg.algo.v_learner_gpu}")
        self.p_learner_device = torch.device(f"cuda:{cfg.algo.p_learner_gpu}")
        self._actor = None
        self.obs = None
        obs_dim = (self.obs_dim,) if isinstance(self.obs_dim, int) else tuple(self.obs_dim)
        if len(obs_dim) != 1:
            raise NotImplementedError("only flat observations are on the PQL path")
        self._O, self._A, self._E = int(obs_dim[0]), int(self.action_dim), int(cfg.num_envs)
        dev = self.sim_device

        L = int(self.cfg.algo.tracker_len)
        self.return_tracker = EpisodeTracker(L, dev)
        self.step_tracker = EpisodeTracker(L, dev)
        self._pushed = torch.zeros(1, dtype=torch.int64, device=dev)
        self.current_returns = torch.zeros(self._E, dtype=torch.float32, device=dev)
        self.current_lengths = torch.zeros(self._E, dtype=torch.float32, device=dev)
        if self.cfg.info_track_keys is not None:
            raise NotImplementedError("info_track_keys (env-specific logging, pql_actor.py:28-32) is not on the path")

        if self.cfg.algo.obs_norm:
            # data parallel (cfg.data_parallel): one normaliser for the whole job - the batch moments are summed over
            # the ranks before every update (the reference has ONE actor)
            self.obs_rms = RunningMeanStd(shape=self.obs_dim, device=dev, data_parallel=bool(getattr(cfg, "data_parallel", False)))
            if self.cfg.artifact is not None:
                raise NotImplementedError("W&B artifact loading is out of scope: use obs_rms.load_state_dict()")
        else:
            self.obs_rms = None
        self.n_step_buffer = NStepReplay(self.obs_dim, self.action_dim, self.cfg.num_envs, self.cfg.algo.nstep, device=dev)

        noise = self.cfg.algo.noise
        if noise.decay == 'linear':
            self.noise_scheduler = LinearSchedule(start_val=noise.std_max, end_val=noise.std_min,
                                                  total_iters=noise.lin_decay_iters)
        elif noise.decay == 'exp':
            self.noise_scheduler = ExponentialSchedule(start_val=noise.std_max, gamma=self.cfg.algo.exp_decay_rate,
                                                       end_val=noise.std_min)
        else:
            self.noise_scheduler = None
        # this actor's random stream (the reference draws from the process's default CUDA generator)
        self.generator = torch.Generator(device=dev)
        self.generator.manual_seed((torch.cuda.default_generators[dev.index].initial_seed() + 2) % (1 << 63))
        # one std per env for the mixed noise: linspace(std_min, std_max, E) (noise.py:31-32), built once
        self._row_std = torch.linspace(noise.std_min, noise.std_max, self._E).to(dev)
        self._plan = None

    # ---- policy hand-off (train_pql.py:52,109: ``pql_actor.actor = deepcopy(actor).to(sim_device)``) ----
    @property
    def actor(self):
        return self._actor

    @actor.setter
    def actor(self, module):
        self._actor = module
        self._weights_stale = True

    def _build_plan(self):
        E, O, A, dev = self._E, self._O, self._A, self.sim_device
        La = NetLayout(O, A, 1)
        p = type("ActorStepPlan", (), {})()
        p.La = La
        p.x_ld, p.a_ld = La.ldw[0], _ru(A, 4)
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)       # noqa: E731
        p.a_flat, p.a_tf = z(La.total), z(La.total)
        p.x, p.noise, p.act_tf = z(E, p.x_ld), z(E, A), z(E, p.a_ld)
        p.h = [z(E, d) for d in HIDDEN]
        # observations wider than 128 (ShadowHand: 211): the layer-fused TF32 kernel does not take them, the wide-input
        # split-fp16 kernel does (one fp16 MMA per product = the accuracy of the TF32 launches it replaces); it reads
        # fp16 copies of the weights, refreshed with the TF32 twin whenever the driver hands over a new actor
        wide = FUSED_MAX_IN < O <= FUSED_H_MAX_IN and La.ldw[0] % 8 == 0 and A % 4 == 0 and A <= 32
        p.a_h = half_arena(La, dev) if wide else None
        net = NetAddrs(La, 0, p.a_tf, p.a_flat, p.a_h)

        def make(noisy):
            act = dict(out=K.addr(p.act_tf), ldo=p.a_ld, out2=K.addr(p.act_tf), ldo2=p.a_ld)
            if noisy:       # the draw is pre-scaled by its std (pqlb_actor_inputs): std 1, no noise clamp
                act.update(noise=K.addr(p.noise), ldnoise=A, noise_std=1.0, noise_bound=3.0e38)
            inst = dict(net=net, x=K.addr(p.x), x_ld=p.x_ld, k_in=O, h=[K.addr(t) for t in p.h],
                        store=(False, False, False), act=act, terms=1)
            return forward_calls(E, [inst], False)
        p.calls = {True: make(True), False: make(False)}
        self._plan = p

    def _head_desc(self, calls):
        """(descriptor group, field prefix) of the launch that writes the action."""
        last = calls[-1]
        if isinstance(last, (K.MlpForward, K.MlpForwardH)):
            return last.desc.g[0], "act_"
        return last.desc.g[0], ""

    def _sync_weights(self):
        if self._actor is None:
            raise RuntimeError("PQLActor.actor has not been set")
        if self._plan is None:
            self._build_plan()
        if self._weights_stale:
            p = self._plan
            src = _PRODUCER_STREAM.get(self._actor)
            if src is not None:
                torch.cuda.current_stream(self.sim_device).wait_stream(src)
            p.a_flat.copy_(module_flat(self._actor, p.La.total, self.sim_device), non_blocking=True)
            note_read(self._actor)           # the P-learner's next in-place update waits for this copy
            _operand_copies(p.a_flat, p.a_tf, p.a_h)
            self._weights_stale = False

    # ---- reference API -------------------------------------------------------------------------
    def reset_agent(self):
        self.obs = self.env.reset()

    def get_noise_std(self):
        if self.noise_scheduler is None:
            return self.cfg.algo.noise.std_max
        return self.noise_scheduler.val()

    def update_noise(self):
        if self.noise_scheduler is not None:
            self.noise_scheduler.step()

    def _as_obs(self, obs):
        obs = obs.reshape(self._E, self._O)
        if obs.device != self.sim_device or obs.dtype != torch.float32:
            obs = obs.to(device=self.sim_device, dtype=torch.float32)
        return obs if obs.stride(1) == 1 else obs.contiguous()

    @torch.no_grad()
    def get_actions(self, obs, sample=True, out=None):
        """pql_actor.py:70-85: normalise, policy, exploration noise, clamp to [-1, 1].  ``out``
        (optional) is an fp32 [E, A] view whose row stride is a multiple of 4 words."""
        E, O, A = self._E, self._O, self._A
        obs = self._as_obs(obs)
        if out is None:
            out = torch.empty((E, A), dtype=torch.float32, device=self.sim_device)
        with torch.cuda.device(self.sim_device):
            self._sync_weights()
            p = self._plan
            rms = self.obs_rms if self.cfg.algo.obs_norm else None
            noise_ptr, row_std, std, seed, offset = None, None, 0.0, 0, 0
            if sample:
                kind = self.cfg.algo.noise.type
                if kind == 'fixed':
                    std = float(self.get_noise_std())
                elif kind == 'mixed':
                    row_std = _lib.ptr(self._row_std)
                else:
                    raise NotImplementedError
                noise_ptr = _lib.ptr(p.noise)
                seed, offset = self._next_draw()
            _lib.call("pqlb_actor_inputs", _lib.ptr(obs), E, O, obs.stride(0),
                      _lib.ptr(rms.mean) if rms is not None else None, _lib.ptr(rms.var) if rms is not None else None,
                      float(rms.epsilon) if rms is not None else 0.0, 0, 1, _lib.ptr(p.x), p.x_ld,
                      noise_ptr, A, row_std, std, seed, offset)
            calls = p.calls[bool(sample)]
            g, pre = self._head_desc(calls)
            setattr(g, pre + "out2", out.data_ptr())
            setattr(g, pre + "ldo2", out.stride(0))
            for c in calls:
                c()
        return out

    def _next_draw(self):
        """(seed, offset) of the next 4-offset draw of this actor's generator; advances it."""
        gen = self.generator
        seed, off = int(gen.initial_seed()), int(gen.get_offset())
        gen.set_offset(off + 4)
        return (seed - (1 << 64) if seed >= 1 << 63 else seed), off

    @torch.no_grad()
    def explore_env(self, env, timesteps: int, random: bool) -> list:
        """pql_actor.py:87-127."""
        E, O, A, T, dev = self._E, self._O, self._A, int(timesteps), self.sim_device
        f32 = dict(dtype=torch.float32, device=dev)
        traj_actions = torch.empty((E, T, A), **f32)
        traj_rewards = torch.empty((E, T), **f32)
        traj_dones = torch.empty((E, T), **f32)
        staged = T > 1           # one step per call: the step's own tensors are the trajectory
        if staged:
            traj_states = torch.empty((E, T, O), **f32)
            traj_next_states = torch.empty((E, T, O), **f32)
        fused_head = (A * T) % 4 == 0
        obs = self._as_obs(self.obs)
        with torch.cuda.device(dev):
            for i in range(T):
                if self.cfg.algo.obs_norm:
                    self.obs_rms.update(obs)
                if random:
                    action = torch.rand((E, A), device=dev, generator=self.generator) * 2.0 - 1.0
                    traj_actions[:, i] = action
                elif fused_head:
                    action = self.get_actions(obs, sample=True, out=traj_actions[:, i])
                else:
                    action = self.get_actions(obs, sample=True)
                    traj_actions[:, i] = action
                next_obs, reward, done, info = env.step(action)
                next_obs = self._as_obs(next_obs)
                self._env_post(reward, done, info, traj_rewards[:, i], traj_dones[:, i], T)
                if staged:
                    traj_states[:, i] = obs
                    traj_next_states[:, i] = next_obs
                else:
                    traj_states, traj_next_states = obs.reshape(E, 1, O), next_obs.reshape(E, 1, O)
                obs = next_obs
        self.obs = obs
        obs, action, reward, next_obs, done = self.n_step_buffer.add_to_buffer(
            traj_states, traj_actions, traj_rewards.reshape(E, T, 1), traj_next_states, traj_dones.reshape(E, T, 1))
        act_data = obs if self.p_learner_device == dev else obs.to(self.p_learner_device)
        v = self.v_learner_device
        cri_data = (obs, action, reward, next_obs, done) if v == dev else tuple(t.to(v) for t in (obs, action, reward, next_obs, done))
        return act_data, cri_data, T * self.cfg.num_envs

    def _env_post(self, reward, done, info, reward_out, done_out, out_stride):
        """update_tracker (pql_actor.py:129-136) + handle_timeout (common.py:195-202) + reward scaling
        (pql_actor.py:117) in one launch; writes column i of the [E, T] trajectory tensors."""
        E, dev = self._E, self.sim_device
        reward = reward.reshape(E)
        if reward.device != dev or reward.dtype != torch.float32 or reward.stride(0) != 1:
            reward = reward.to(device=dev, dtype=torch.float32).contiguous()
        done = done.reshape(E)
        if done.device != dev or done.dtype != torch.float32 or done.stride(0) != 1:
            done = done.to(device=dev, dtype=torch.float32).contiguous()
        trunc = None
        if self.cfg.algo.handle_timeout and hasattr(info, "get"):
            trunc = info.get('TimeLimit.truncated')
            if trunc is not None:
                trunc = trunc.reshape(E)
                if trunc.device != dev or trunc.dtype not in (torch.bool, torch.uint8) or trunc.stride(0) != 1:
                    trunc = (trunc.to(dev) != 0).contiguous()
        if out_stride != 1:          # strided columns of [E, T]: stage through contiguous rows
            r_tmp, d_tmp = torch.empty(E, dtype=torch.float32, device=dev), torch.empty(E, dtype=torch.float32, device=dev)
        else:
            r_tmp, d_tmp = reward_out, done_out
        _lib.call("pqlb_env_post", _lib.ptr(reward), _lib.ptr(done), _lib.ptr(trunc), float(self.cfg.algo.reward_scale), E,
                  _lib.ptr(self.current_returns), _lib.ptr(self.current_lengths), _lib.ptr(self.return_tracker.window),
                  _lib.ptr(self.step_tracker.window), self.return_tracker.max_len, _lib.ptr(self._pushed),
                  _lib.ptr(r_tmp), _lib.ptr(d_tmp))
        if out_stride != 1:
            reward_out.copy_(r_tmp)
            done_out.copy_(d_tmp)

    def update_tracker(self, reward, done, info):
        """pql_actor.py:129-150 as a stand-alone call (explore_env fuses it with the timeout handling)."""
        with torch.cuda.device(self.sim_device):
            self._env_post(reward, done, None, None, None, 1)
        return done

    def add_info_tracker_log(self, log_info):
        pass          # info_track_keys is None (checked in the constructor)
