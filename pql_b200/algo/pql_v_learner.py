"""V-learner (critic update) entry points: drop-in for pql/algo/pql_v_learner.py:22-133.

Same constructor, ``start()``, ``learn()``, ``update(actor, trajectory, normalize_tuple, sleep_time)``
and the same attributes (critic, critic_target weights, memory, loss_tracker, update_count).
``learn()`` is a fixed list of sm_100a kernel launches (pql_b200/algo/_engine.py), replayed from a CUDA
graph; the reference's two random draws per update (torch.randint for the indices, torch.normal for
the target-policy noise) are made inside the gather kernel from this learner's generator state, bit
for bit what torch returns for it (``cfg.fused_rng = False``: the torch calls themselves).  The reference's
``@ray.remote`` decoration is the caller's business: these are plain classes,
one process per GPU.
"""
import contextlib
import os
import weakref

import torch

from .. import _lib
from ..models import load_class
from ..replay.simple_replay import ReplayBuffer
from ..utils.common import DeviceTracker
from . import _dp
from ._engine import CriticUpdate


_PRODUCER_STREAM = weakref.WeakKeyDictionary()     # module -> the learner stream that writes its weights
_READ_EVENTS = weakref.WeakKeyDictionary()         # module -> events recorded after other streams' copies of its live arena


def note_read(module, stream=None):
    """A consumer has just enqueued a copy of ``module``'s live arena on ``stream`` (default: the
    current one).  The reference hands over snapshots (Ray pickles the module at return time,
    SURVEY App. A 18); here the weights are read in place, so the producer's next in-place write
    (AdamW in its ``learn()``) must be ordered after this copy: ``wait_readers`` makes it so."""
    if module is None or _PRODUCER_STREAM.get(module) is None:
        return
    ev = torch.cuda.Event()
    ev.record(stream if stream is not None else torch.cuda.current_stream())
    _READ_EVENTS.setdefault(module, []).append(ev)


def wait_readers(module, stream):
    """Order ``stream``'s next writes to ``module``'s arena after every pending copy of it."""
    evs = _READ_EVENTS.pop(module, None) if module is not None else None
    if evs:
        for ev in evs:
            stream.wait_event(ev)


class LearnerStream:
    """Optional per-learner CUDA stream (``cfg.learner_streams = True``).

    In the reference the V-learner and the P-learner are separate Ray actors (scripts/train_pql.py:
    40-52), possibly on the same GPU, so their updates overlap.  A single process hosting both gets
    the same concurrency by giving each learner its own stream: ``learn()`` enqueues there and
    returns, ``update()`` is the exchange point and orders this learner's stream after the caller's
    current stream (which produced the trajectory) and after the stream that produced the module it
    is handed.  Off by default: then everything runs on the caller's current stream."""

    def __init__(self, cfg, device, priority=0):
        self.device = device
        # priority < 0: this learner's kernels are dispatched ahead of the other learner's whenever SMs free up
        # (the V-learner is the longer chain of a step, 8 updates against 4: see DESIGN.md section 7)
        self.stream = torch.cuda.Stream(device, priority=int(priority)) if getattr(cfg, "learner_streams", False) else None

    def ctx(self):
        return torch.cuda.stream(self.stream) if self.stream is not None else contextlib.nullcontext()

    def tag(self, module):
        """Modules handed out by update()/start() remember which stream writes them."""
        if self.stream is not None:
            _PRODUCER_STREAM[module] = self.stream
        return module

    def join(self, module, tensors=()):
        if self.stream is None:
            return
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        src = _PRODUCER_STREAM.get(module) if module is not None else None
        if src is not None and src is not self.stream:
            self.stream.wait_stream(src)
        for t in tensors:
            if torch.is_tensor(t) and t.is_cuda:
                t.record_stream(self.stream)


def make_generator(cfg, device, salt):
    """The learner's own random stream.  In the reference every learner is a process of its own with
    its own default CUDA generator (scripts/train_pql.py:40-52); here a learner owns a
    ``torch.Generator`` seeded from the device's default generator (so ``torch.manual_seed`` governs it;
    ``salt`` keeps the two learners of one process from drawing the same indices).  With
    ``cfg.fused_rng`` (default) the draws are made inside the gather kernel from this generator's
    (seed, offset) - bit-identical to calling torch.randint / normal_ with it (csrc/rng.cuh) - and two
    torch launches per update disappear; ``cfg.fused_rng = False`` makes those torch calls instead."""
    gen = torch.Generator(device=device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    gen.manual_seed((torch.cuda.default_generators[idx].initial_seed() + salt) % (1 << 63))
    fused = bool(getattr(cfg, "fused_rng", True)) and not os.environ.get("PQLB_NO_FUSED_RNG")
    return gen, fused


def module_flat(module, layout_total, device):
    """Flat fp32 arena of an actor / critic module in the kernel layout: our modules expose it
    directly, a reference ``nn.Module`` with the same state_dict keys is repacked."""
    if hasattr(module, "arena"):
        flat = module.arena.flat
        if flat.numel() != layout_total:
            raise ValueError("module shape does not match this learner")
        return flat.to(device, non_blocking=True)
    return repack_reference_module(module, layout_total, device)


def repack_reference_module(module, layout_total, device):
    """Flat kernel-layout arena of a *reference* module (pql/models/mlp.py: ``TanhMLPPolicy`` /
    ``MLPNet`` with ``net.{0,2,4,6}.{weight,bias}``, ``DoubleQ`` / ``DistributionalDoubleQ`` with
    ``net_q1.net.*`` / ``net_q2.net.*``), read through its ``state_dict`` - what ``update()`` is handed
    when the caller still constructs the reference's own classes (pql_v_learner.py:117-122)."""
    from ..models.mlp import HIDDEN, NetLayout, ParamArena
    if not hasattr(module, "state_dict"):
        raise TypeError("expected an actor / critic nn.Module")
    sd = module.state_dict()
    prefixes = ["net_q1.net.", "net_q2.net."] if any(k.startswith("net_q1.") for k in sd) else ["net."]
    try:
        w0 = sd[prefixes[0] + "0.weight"]
        w_last = sd[prefixes[0] + "6.weight"]
        hidden = tuple(sd[prefixes[0] + f"{k}.weight"].shape[0] for k in (0, 2, 4))
    except KeyError as e:
        raise TypeError(f"module has no reference MLP state_dict keys ({e})") from None
    if hidden != tuple(HIDDEN):
        raise NotImplementedError(f"the kernels are specialised to hidden_layers={list(HIDDEN)}, module has {list(hidden)}")
    layout = NetLayout(w0.shape[1], w_last.shape[0], len(prefixes))
    if layout.total != layout_total:
        raise ValueError("module shape does not match this learner")
    arena = ParamArena(layout, "cpu")
    for i, pre in enumerate(prefixes):
        for l, k in enumerate((0, 2, 4, 6)):
            arena.weight(i, l).copy_(sd[f"{pre}{k}.weight"].detach().float().cpu())
            arena.bias(i, l).copy_(sd[f"{pre}{k}.bias"].detach().float().cpu())
    return arena.flat.to(device)


def carry_plan_state(old, new):
    """Optimiser moments, step count and Polyak target of ``old`` into the rebuilt plan ``new``."""
    new.opt.m.copy_(old.opt.m)
    new.opt.v.copy_(old.opt.v)
    new.opt.count.copy_(old.opt.count)
    if getattr(old, "t_flat", None) is not None and getattr(new, "t_flat", None) is not None and hasattr(new, "tau"):
        new.t_flat.copy_(old.t_flat)
        new.round_weights()


class PQLVLearner:
    def __init__(self, obs_dim, action_dim, cfg, process_group=None):
        self.cfg = cfg
        self.obs_dim = obs_dim
        self.action_dim = action_dim
        self.device = torch.device(f"cuda:{self.cfg.algo.v_learner_gpu}")
        if not torch.cuda.is_available():
            raise RuntimeError("PQLVLearner needs a CUDA device: pql_b200 has no CPU path")
        _lib.load()
        if self.cfg.algo.distl and "Distributional" not in self.cfg.algo.cri_class:
            self.cfg.algo.cri_class = "Distributional" + self.cfg.algo.cri_class       # :30-31
        cri_class = load_class(self.cfg.algo.cri_class)
        if self.cfg.algo.distl:
            self.critic = cri_class(self.obs_dim, self.action_dim, v_min=self.cfg.algo.v_min, v_max=self.cfg.algo.v_max,
                                    num_atoms=self.cfg.algo.num_atoms, device=self.device).to(self.device)
        else:
            self.critic = cri_class(self.obs_dim, self.action_dim).to(self.device)
        if self.cfg.artifact is not None:
            raise NotImplementedError("W&B artifact loading (pql/utils/model_util.py) is out of scope: "
                                      "use critic.load_state_dict()")
        self.actor = None
        self.memory = ReplayBuffer(capacity=int(cfg.algo.memory_size), obs_dim=self.obs_dim,
                                   action_dim=self.action_dim, device=self.device)
        self.loss_tracker = DeviceTracker(5, self.device)
        self.update_count = 0
        self.normalize_tuple = None
        self.sleep_time = 0
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
        self._ls = LearnerStream(cfg, self.device, priority=int(os.environ.get("PQLB_V_PRIORITY", getattr(cfg, "v_stream_priority", 0))))
        self._ls.tag(self.critic)
        self.generator, self.fused_rng = make_generator(cfg, self.device, salt=0)
        if self.world_size > 1:
            # ranks are seeded differently (distinct envs / replay shards), so their default-initialised
            # critics differ: rank 0's weights become everyone's BEFORE the Polyak target is cloned and the
            # tensor-core twins are rounded (_build), or the replicas would share gradients but never agree
            _dp.broadcast_(self.critic.arena.flat, 0, process_group)

    @property
    def stream(self):
        return self._ls.stream

    def disable_graph(self):
        self.use_cuda_graph = False

    def enable_graph(self):
        self.use_cuda_graph = True

    # ---- plan ------------------------------------------------------------------------------
    def _obs_dim_int(self):
        return int(self.obs_dim if isinstance(self.obs_dim, int) else self.obs_dim[0])

    def _build(self):
        a = self.cfg.algo
        eps = 1e-4 if self.normalize_tuple is None else float(self.normalize_tuple[2])
        old = self._plan
        self._plan = CriticUpdate(self._obs_dim_int(), self.action_dim, int(a.batch_size), self.device,
                                  self.critic.arena.flat, distl=bool(a.distl), num_atoms=a.num_atoms, v_min=a.v_min,
                                  v_max=a.v_max, gamma_n=a.gamma ** a.nstep, lr=a.critic_lr, tau=a.tau,
                                  max_grad_norm=a.max_grad_norm, noise_bound=a.noise.tgt_pol_noise_bound,
                                  noise_std=a.noise.tgt_pol_std,
                                  obs_norm=bool(a.obs_norm) and self.normalize_tuple is not None, eps=eps,
                                  world_size=self.world_size, loss_ring=self.loss_tracker.window,
                                  process_group=self.process_group, dp_fused=self.dp_fused,
                                  fwd_mode=getattr(self.cfg, "forward_mode", None))
        if old is not None:
            # normalize_tuple switched between None and a tuple: the launch list is rebuilt, the learner's
            # state is not - AdamW moments, step count and the Polyak target carry over, and the fused
            # sampler's stream continues where it was (enable_fused_rng rebases on the restored count)
            carry_plan_state(old, self._plan)
        if self.fused_rng and self.memory.capacity < (1 << 28):
            if old is not None and old.rng_state is not None:
                self.generator.set_offset(int(old.rng_state[1].item()) + int(old.rng_state[2].item()) * old.opt.step)
            self._plan.enable_fused_rng(self.generator, draws_per_update=2)      # randint, then normal
        self._sample = self._plan.bind_replay(self.memory.ring, self.memory.capacity, self.memory.cur_capacity_dev)

    @property
    def critic_target(self):
        """Snapshot of the Polyak target as a module (the live weights are the plan's flat arena)."""
        import copy
        tgt = copy.deepcopy(self.critic)
        if self._plan is not None:
            tgt.arena.flat.copy_(self._plan.t_flat)
        return tgt

    def start(self):
        return self.critic, self.update_count, self.loss_tracker.mean()

    def _allreduce(self, grad):
        _dp.allreduce_sum_(grad, self.process_group)

    @torch.no_grad()
    def learn(self):
        if self.actor is not None:
            if self._plan is None:
                self._build()
            p = self._plan
            with torch.cuda.device(self.device), self._ls.ctx():
                # same two draws, in the same order, as the reference: randint (simple_replay.py:87)
                # then torch.normal(zeros, full(std)) (noise.py:20-21), which ATen evaluates as
                # out.normal_(0, 1).mul_(std).add_(mean): we draw the N(0,1) part with the same
                # generator call and apply std inside the actor-head epilogue (this also avoids
                # torch.normal's std.min() >= 0 check, a device->host sync per update).
                if p.rng_state is None:
                    torch.randint(self.memory.cur_capacity, size=(p.B,), device=self.device, out=p.idx,
                                  generator=self.generator)
                    p.noise.normal_(generator=self.generator)
                elif self.memory.cur_capacity <= 0:       # what torch.randint(0, ...) raises in the reference
                    raise RuntimeError("learn(): the replay buffer is empty (random_ expects 'from' to be less than 'to')")
                wait_readers(self.critic, torch.cuda.current_stream(self.device))     # pending copies of the live critic
                p.run(self._sample, self._allreduce if self.world_size > 1 and p.dp is None else None, self.use_cuda_graph,
                      self.graph_allreduce)
            self.update_count += 1
        return self.sleep_time

    @torch.no_grad()
    def update(self, actor, trajectory, normalize_tuple, sleep_time):
        self.actor = actor
        with torch.cuda.device(self.device):
            self._ls.join(actor, tuple(trajectory) + (tuple(normalize_tuple[:2]) if normalize_tuple is not None else ()))
            with self._ls.ctx():
                self.memory.add_to_buffer(trajectory)
                rebuild = self._plan is not None and ((normalize_tuple is None) != (self.normalize_tuple is None))
                self.normalize_tuple = normalize_tuple
                self.sleep_time = sleep_time
                if self._plan is None or rebuild:
                    self._build()
                self._plan.set_actor(module_flat(actor, self._plan.La.total, self.device))
                note_read(actor)
                self._plan.set_norm(normalize_tuple if self.cfg.algo.obs_norm else None)
                self._plan.invalidate_prefetch()      # new transitions / actor / statistics: the batch drawn ahead is stale
                # cfg.sync_loss: block until this learner's stream has drained and return the current mean;
                # default: the mean as of the previous update() (non-blocking, DeviceTracker.mean_lagged)
                loss = self.loss_tracker.mean() if self._sync_loss else self.loss_tracker.mean_lagged()
        return self.critic, loss, self.update_count
