"""P-learner (policy update) entry points: drop-in for pql/algo/pql_p_learner.py:16-96.

Owns the actor, its AdamW state and the observation-only ring (``memory``, ``next_p``,
``if_full``, ``cur_capacity`` as in the reference); ``learn()`` is a fixed list of sm_100a kernel
launches (DPG through the frozen critic) whose first kernel also draws the reference's
torch.randint indices from this learner's generator state (``cfg.fused_rng = False``: torch.randint)."""
import os

import torch

from .. import _lib
from ..models import load_class
from ..utils.common import DeviceTracker
from . import _dp
from ._engine import ActorUpdate
from .pql_v_learner import LearnerStream, carry_plan_state, make_generator, module_flat, note_read, wait_readers


class PQLPLearner:
    def __init__(self, obs_dim, action_dim, cfg, process_group=None):
        self.cfg = cfg
        self.obs_dim = obs_dim
        self.action_dim = action_dim
        self.device = torch.device(f"cuda:{self.cfg.algo.p_learner_gpu}")
        if not torch.cuda.is_available():
            raise RuntimeError("PQLPLearner needs a CUDA device: pql_b200 has no CPU path")
        _lib.load()
        act_class = load_class(self.cfg.algo.act_class)
        self.actor = act_class(self.obs_dim, self.action_dim).to(self.device)
        if self.cfg.artifact is not None:
            raise NotImplementedError("W&B artifact loading (pql/utils/model_util.py) is out of scope: "
                                      "use actor.load_state_dict()")
        self.critic = None
        obs_dim = (self.obs_dim,) if isinstance(self.obs_dim, int) else tuple(self.obs_dim)
        if len(obs_dim) != 1:
            raise NotImplementedError("only flat observations are on the PQL path")
        self._O = int(obs_dim[0])
        self.memory_size = int(self.cfg.algo.memory_size)
        self.memory = torch.empty((self.memory_size, self._O), dtype=torch.float32, device=self.device)
        self.next_p = 0
        self.if_full = False
        self.cur_capacity = 0
        self.loss_tracker = DeviceTracker(5, self.device)
        self.update_count = 0
        self.normalize_tuple = None
        self.sleep_time = 0.01
        self.process_group = process_group
        self.world_size = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()
                                         and getattr(cfg, "data_parallel", False)):
            self.world_size = torch.distributed.get_world_size(process_group)
        self._plan = None
        self._sample = None
        self.use_cuda_graph = bool(getattr(cfg, "use_cuda_graph", True)) and not os.environ.get("PQLB_NO_GRAPH")
        # capture the gradient all-reduce into the update's CUDA graph (needs a communicator of this
        # learner's own: pass process_group=dist.new_group(...) per learner)
        self.graph_allreduce = bool(getattr(cfg, "dp_graph_allreduce", False)) and process_group is not None
        # cfg.dp_fused: no NCCL on the path - the optimiser kernel all-reduces over symmetric memory
        self.dp_fused = bool(getattr(cfg, "dp_fused", False)) and self.world_size > 1
        self._sync_loss = bool(getattr(cfg, "sync_loss", False))
        self._ls = LearnerStream(cfg, self.device, priority=int(os.environ.get("PQLB_P_PRIORITY", getattr(cfg, "p_stream_priority", 0))))
        self._ls.tag(self.actor)
        self.generator, self.fused_rng = make_generator(cfg, self.device, salt=1)
        if self.world_size > 1:          # rank 0's initial actor becomes everyone's (see PQLVLearner)
            _dp.broadcast_(self.actor.arena.flat, 0, process_group)
        self.cur_capacity_dev = torch.zeros(1, dtype=torch.int64, device=self.device)

    @property
    def stream(self):
        return self._ls.stream

    def disable_graph(self):
        self.use_cuda_graph = False

    def enable_graph(self):
        self.use_cuda_graph = True

    def _build(self):
        a = self.cfg.algo
        distl = bool(a.distl)
        eps = 1e-4 if self.normalize_tuple is None else float(self.normalize_tuple[2])
        old = self._plan
        self._plan = ActorUpdate(self._O, self.action_dim, int(a.batch_size), self.device, self.actor.arena.flat,
                                 distl=distl, num_atoms=a.num_atoms, v_min=a.v_min, v_max=a.v_max, lr=a.actor_lr,
                                 max_grad_norm=a.max_grad_norm,
                                 obs_norm=bool(a.obs_norm) and self.normalize_tuple is not None, eps=eps,
                                 world_size=self.world_size, loss_ring=self.loss_tracker.window,
                                  process_group=self.process_group, dp_fused=self.dp_fused,
                                  fwd_mode=getattr(self.cfg, "forward_mode", None))
        if old is not None:              # rebuilt launch list, same learner state (see PQLVLearner._build)
            carry_plan_state(old, self._plan)
        if self.fused_rng and self.memory_size < (1 << 28):
            if old is not None and old.rng_state is not None:
                self.generator.set_offset(int(old.rng_state[1].item()) + int(old.rng_state[2].item()) * old.opt.step)
            self._plan.enable_fused_rng(self.generator, draws_per_update=1)      # randint
        self._sample = self._plan.sample_call(self.memory, self.memory_size, self.cur_capacity_dev)

    def start(self):
        return self.actor, self.update_count, self.loss_tracker.mean()

    def _allreduce(self, grad):
        _dp.allreduce_sum_(grad, self.process_group)

    @torch.no_grad()
    def learn(self):
        if self.critic is not None:
            p = self._plan
            with torch.cuda.device(self.device), self._ls.ctx():
                if p.rng_state is None:
                    torch.randint(self.cur_capacity, size=(p.B,), device=self.device, out=p.idx,
                                  generator=self.generator)                                          # :49
                elif self.cur_capacity <= 0:
                    raise RuntimeError("learn(): the observation ring is empty (random_ expects 'from' to be less than 'to')")
                wait_readers(self.actor, torch.cuda.current_stream(self.device))      # pending copies of the live actor
                p.run(self._sample, self._allreduce if self.world_size > 1 and p.dp is None else None, self.use_cuda_graph,
                      self.graph_allreduce)
            self.update_count += 1
        return self.sleep_time

    @torch.no_grad()
    def update(self, critic, obs, normalize_tuple, sleep_time):
        self.critic = critic
        self.sleep_time = sleep_time
        rebuild = self._plan is not None and ((normalize_tuple is None) != (self.normalize_tuple is None))
        self.normalize_tuple = normalize_tuple
        obs = obs.reshape(-1, self._O)
        with torch.cuda.device(self.device):
            self._ls.join(critic, (obs,) + (tuple(normalize_tuple[:2]) if normalize_tuple is not None else ()))
            with self._ls.ctx():
                if obs.device != self.device or obs.dtype != torch.float32:
                    obs = obs.to(device=self.device, dtype=torch.float32, non_blocking=True)
                obs = obs.contiguous()
                self.add_capacity = obs.shape[0]
                p = self.next_p + self.add_capacity
                if p > self.memory_size and p - self.memory_size > self.memory_size:
                    raise RuntimeError(f"update: {self.add_capacity} observations do not fit a ring of {self.memory_size}")
                _lib.call("pqlb_obsring_insert", _lib.ptr(self.memory), self.memory_size, self._O, _lib.ptr(obs),
                          self.add_capacity, self.next_p)
                if p > self.memory_size:                   # :73-77, strict
                    p = p - self.memory_size
                    self.if_full = True
                self.next_p = p
                self.cur_capacity = self.memory_size if self.if_full else self.next_p
                _lib.call("pqlb_store_i64", _lib.ptr(self.cur_capacity_dev), self.cur_capacity)
                if self._plan is None or rebuild:
                    self._build()
                self._plan.set_critic(module_flat(critic, self._plan.Lc.total, self.device))
                note_read(critic)
                self._plan.set_norm(normalize_tuple if self.cfg.algo.obs_norm else None)
                # cfg.sync_loss: block until this learner's stream has drained and return the current mean;
                # default: the mean as of the previous update() (non-blocking, DeviceTracker.mean_lagged)
                loss = self.loss_tracker.mean() if self._sync_loss else self.loss_tracker.mean_lagged()
        return self.actor, loss, self.update_count
