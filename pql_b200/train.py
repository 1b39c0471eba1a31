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
import time

import torch

from .algo import PQLActor, PQLPLearner, PQLVLearner


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
