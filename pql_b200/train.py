"""Lock-step driver of the PQL loop: what scripts/train_pql.py:44-158 does with Ray
actors, ``ray.wait`` polling and a sleep-based rate balancer, as one process per GPU enqueueing a
fixed schedule on CUDA streams.

The reference keeps the three workers at the ratios ``critic_sample_ratio`` (critic updates per env
step, pql_algo.yaml:18) and ``critic_actor_ratio`` (critic updates per actor update, :17) by
measuring their speeds and sleeping (train_pql.py:118-142).  Here the ratios are exact by
construction: per env step one ``explore_env`` (actor-side kernels + ``env.step``), the two
``update()`` exchanges (ring inserts, weight / normaliser hand-off), then ``critic_sample_ratio``
critic updates interleaved with ``critic_sample_ratio / critic_actor_ratio`` actor updates, each a
CUDA-graph replay on its learner's stream.  Weights move between the workers as device-to-device
copies of flat arenas (no pickling), and the host never waits for the step it has just enqueued: the
loss read-back inside ``update()`` returns the previous step's value (DeviceTracker.mean_lagged).
"""
import os
import time

import torch

from .algo import PQLActor, PQLPLearner, PQLVLearner
from .algo.pql_v_learner import wait_readers


class LockStepTrainer:
    def __init__(self, env, cfg, process_group_v=None, process_group_p=None):
        self.env, self.cfg = env, cfg
        self.actor_worker = PQLActor(env, cfg)
        obs_shape, act_dim = env.observation_space.shape, env.action_space.shape[0]
        self.v_learner = PQLVLearner(obs_shape, act_dim, cfg, process_group=process_group_v)
        self.p_learner = PQLPLearner(obs_shape, act_dim, cfg, process_group=process_group_p)
        self.critic, self.critic_updates, self.critic_loss = self.v_learner.start()          # train_pql.py:50-51
        self.actor, self.actor_updates, self.actor_loss = self.p_learner.start()
        self.actor_worker.actor = self.actor
        self.v_per_step = int(cfg.algo.critic_sample_ratio)
        self.p_every = max(1, int(cfg.algo.critic_actor_ratio))
        self.global_steps = 0
        self.sim_count = 0
        self._t0 = None
        self.observer = None          # optional callable(kind) after every learn() ("v" / "p"): tests read the plans' draws
        self._block, self._block_streams, self._block_key = None, None, None     # the per-step graph of all updates (_learn_block)

    def _rms(self, device):
        rms = self.actor_worker.obs_rms
        return rms.get_states(device) if rms is not None else None

    def _exchange(self, p_data, v_data):
        """train_pql.py:60-68 / 112-119: hand the new transitions, the other worker's weights and the
        normaliser statistics to both learners; take their current weights and losses back."""
        a = self.actor_worker
        rms_v = self._rms(a.v_learner_device)             # one snapshot per env step (get_states clones)
        rms_p = rms_v if a.p_learner_device == a.v_learner_device else self._rms(a.p_learner_device)
        self.critic, self.critic_loss, self.critic_updates = self.v_learner.update(self.actor, v_data, rms_v, 0)
        self.actor, self.actor_loss, self.actor_updates = self.p_learner.update(self.critic, p_data, rms_p, 0)
        a.actor = self.actor

    def warm_up(self):
        """train_pql.py:57-68: ``warm_up`` env steps with uniform random actions fill the rings."""
        self.actor_worker.reset_agent()
        p_data, v_data, steps = self.actor_worker.explore_env(self.env, self.cfg.algo.warm_up, random=True)
        self.global_steps += steps
        self._exchange(p_data, v_data)
        self._t0 = time.time()

    def step(self):
        """One iteration of the main loop (train_pql.py:97-158) at the exact ratios."""
        if self._t0 is None:
            self.warm_up()
        p_data, v_data, steps = self.actor_worker.explore_env(self.env, self.cfg.algo.horizon_len, random=False)
        self.global_steps += steps
        self.sim_count += 1
        self._exchange(p_data, v_data)
        if not self._learn_block():
            for j in range(self.v_per_step):
                self.v_learner.learn()
                if self.observer is not None:
                    self.observer("v")
                if (j + 1) % self.p_every == 0:
                    self.p_learner.learn()
                    if self.observer is not None:
                        self.observer("p")
        self.actor_worker.update_noise()
        return {"train/critic_loss": self.critic_loss, "train/actor_loss": self.actor_loss,
                "train/critic_update_times": self.critic_updates, "train/actor_update_times": self.actor_updates,
                "train/global_steps": self.global_steps}

    # ---- all updates of an env step as ONE CUDA graph ------------------------------------------------------------
    def _learn_block(self):
        """The ``critic_sample_ratio`` critic and ``critic_sample_ratio / critic_actor_ratio`` actor updates of this env
        step as ONE graph launch: two branches (the learners are independent between exchanges; the V branch carries
        the look-ahead sampler's side branches), captured once - every pointer of both launch lists is fixed and the
        update indices, fill levels and RNG offsets live on the device.  Returns False (the caller then issues the
        ``learn()`` calls one by one: twelve graph launches) when a test observes individual updates, the learners sit on
        different devices or run without graphs / the fused sampler RNG, or the learners are data parallel."""
        v, p = self.v_learner, self.p_learner
        want = os.environ.get("PQLB_STEP_GRAPH")
        want = bool(getattr(self.cfg, "step_graph", False)) if want is None else want != "0"
        if self.observer is not None or not want:
            return False
        n_v, n_p = self.v_per_step, self.v_per_step // self.p_every
        if v.actor is None or p.critic is None or v.device != p.device or not (v.use_cuda_graph and p.use_cuda_graph):
            return False
        vp, pp = v._plan, p._plan
        if vp is None or pp is None or n_v % 2 or n_p < 1 or self.v_per_step % self.p_every:
            return False
        if not vp.block_ready() or pp.rng_state is None or pp.world_size != 1:
            return False
        if v.memory.cur_capacity <= 0 or p.cur_capacity <= 0:
            return False
        dev = v.device
        with torch.cuda.device(dev):
            if self._block is None or self._block_key != (id(vp), id(pp)):      # first step, or a learner rebuilt its plan
                self._block = self._capture_block(vp, pp, v._sample, p._sample, n_v, n_p, dev)
                self._block_key = (id(vp), id(pp))
            own = v.stream is not None and p.stream is not None
            # with per-learner streams the graph runs on a stream of its own, downstream of both (update() ran there) and
            # upstream of both (their next update() follows it): the caller's stream - the actor worker's env step - is
            # never ordered behind the updates, exactly as with individual learn() calls
            bs = self._block_streams[2] if own else torch.cuda.current_stream(dev)
            if own:
                bs.wait_stream(v.stream)
                bs.wait_stream(p.stream)
            wait_readers(v.critic, bs)
            wait_readers(p.actor, bs)
            with torch.cuda.stream(bs):
                self._block.replay()
            if own:
                v.stream.wait_stream(bs)
                p.stream.wait_stream(bs)
        vp.block_done(n_v)
        pp.graph_launches += n_p * (len(pp.calls) + 1 + 2)
        v.update_count += n_v
        p.update_count += n_p
        return True

    def _capture_block(self, vp, pp, v_sample, p_sample, n_v, n_p, dev):
        torch.cuda.synchronize(dev)
        if self._block_streams is None:
            self._block_streams = (torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev))
        sv, sp = self._block_streams[:2]
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            cur = torch.cuda.current_stream(dev)
            sv.wait_stream(cur)
            sp.wait_stream(cur)
            with torch.cuda.stream(sv):
                for j in range(n_v):
                    vp.emit_prefetching(j == 0, j % 2)
            with torch.cuda.stream(sp):
                for _ in range(n_p):
                    pp._segment_a(p_sample)
                    pp._segment_b()
            cur.wait_stream(sv)
            cur.wait_stream(sp)
        return g

    def run(self, max_env_steps=None, max_time=None, log_every=0, log=print):
        """Run until ``max_env_steps`` transitions or ``max_time`` seconds (train_pql.py:174: max_step /
        max_time); returns the last log dict extended with the episode trackers."""
        info = {}
        while True:
            info = self.step()
            if log_every and self.sim_count % log_every == 0:
                a = self.actor_worker
                info.update({"train/return": a.return_tracker.mean(), "train/episode_length": a.step_tracker.mean(),
                             "train/fps": self.global_steps / max(time.time() - self._t0, 1e-9)})
                log(info)
            if max_env_steps is not None and self.global_steps >= max_env_steps:
                break
            if max_time is not None and time.time() - self._t0 >= max_time:
                break
        torch.cuda.synchronize(self.actor_worker.sim_device)
        return info
