"""Prepared kernel plans of the two PQL updates.

``CriticUpdate`` is one ``PQLVLearner.learn()`` (pql/algo/pql_v_learner.py:73-115) and
``ActorUpdate`` one ``PQLPLearner.learn()`` (pql/algo/pql_p_learner.py:47-64), each expressed as
a fixed list of C-ABI launches over preallocated buffers: gather+normalise, tcgen05 forward
layers with fused ELU / tanh / softmax / scalar-head epilogues, the loss kernel, dgrad and
split-K wgrad GEMMs, one deterministic gradient reduction, and the fused clip + AdamW (+ Polyak)
step.  No PyTorch op computes anything here; torch only owns the memory and the RNG draws the
learners pass in.
"""
import ctypes as C

import numpy as np
import torch

from .. import _kernels as K
from .. import _lib
from ..models.mlp import FUSED_H_MAX_IN, FUSED_MAX_IN, HIDDEN, NetAddrs, NetLayout, _ru, forward_calls, fwd_tile, half_arena

H1, H2, H3 = HIDDEN
import os as _os
SEG = int(_os.environ.get("PQLB_REDUCE_SEG", 256))      # elements per block of the gradient reduction
# every weight gradient of an update in one launch (pqlb_wgrad_multi); "0" = one grouped GEMM per layer (round 1)
WGRAD_MULTI = _os.environ.get("PQLB_WGRAD_MULTI", "1") != "0"
# the V-learner's sampler runs one update ahead, as a side branch of the update's CUDA graph ("0": every update draws its own batch first)
PREFETCH = _os.environ.get("PQLB_PREFETCH", "1") != "0"
PF_FORK = int(_os.environ.get("PQLB_PF_FORK", 1))      # the side branch forks before main launch PF_FORK (0 = the critics' forward)


def forward_mode(requested, obs_dim, action_dim):
    """Number format of the fused forward launches: 'f16x3' (default) - critics with split-fp16
    operands and three MMAs per product, policy nets with one fp16 MMA (pqlb_mlp_forward_h; keeps every
    gradient tensor within 1e-3 of the fp32 reference) - or 'tf32' (the round-1
    kernels: one TF32 MMA per product everywhere).  Shapes 'f16x3' takes: inputs up to 128 wide whose weight
    rows are 16-byte aligned as halves and a policy head the kernel fuses (A <= 16, multiples of 4); and
    critics 129..256 wide (the wide-input kernel: ShadowHand's 231 columns) - there the policy net decides
    for itself (models.mlp.split_f16_ok / fused_head_ok: ShadowHand's 211 observations and 20 actions ride in
    the wide-input kernel too; a policy net it does not take would run through the TF32 launches, which is the
    accuracy class of its one-term fp16 forward anyway).  Everything else runs as 'tf32'."""
    mode = _os.environ.get("PQLB_FWD_MODE") or requested or "f16x3"
    if mode not in ("f16x3", "tf32"):
        raise ValueError(f"forward mode {mode!r}: expected 'f16x3' or 'tf32'")
    O, A = int(obs_dim), int(action_dim)
    narrow = O + A <= FUSED_MAX_IN and _ru(O + A, 4) % 8 == 0 and _ru(O, 4) % 8 == 0 and O % 4 == 0 and A % 4 == 0 and A <= 16
    wide = FUSED_MAX_IN < O + A <= FUSED_H_MAX_IN and _ru(O + A, 4) % 8 == 0
    return mode if narrow or wide else "tf32"


class _Optim:
    """Flat AdamW state + split-K / partial-sum workspace bookkeeping for one parameter arena."""

    def __init__(self, layout, device):
        self.layout = layout
        n = layout.total
        self.grad = torch.zeros(n, device=device)
        self.m = torch.zeros(n, device=device)
        self.v = torch.zeros(n, device=device)
        self.count = torch.zeros(1, dtype=torch.int64, device=device)   # completed updates (device-resident)
        self._segs = []            # rows of the segment table
        self.frozen = False        # second build pass of the same launches (other input set): sources already registered

    @property
    def step(self):
        """Completed AdamW steps (host read = synchronisation; tests and checkpoints only)."""
        return int(self.count.item())

    @step.setter
    def step(self, t):
        self.count.fill_(int(t))

    def add_source(self, arena_off, count, ws_off, ws_stride, n_part):
        """grad[arena_off + i] = sum_{s < n_part} ws[ws_off + s * ws_stride + i], i < count."""
        if self.frozen:
            return
        for o in range(0, count, SEG):
            self._segs.append([arena_off + o, min(SEG, count - o), ws_off + o, ws_stride, n_part])

    def finish(self, device):
        self.seg_table = torch.tensor(self._segs, dtype=torch.int64, device=device)
        self.n_seg = len(self._segs)
        self.sumsq = torch.zeros(self.n_seg, device=device)


def _some(call):
    return [] if call is None else [call]


def _colsum_call(B, entries, keep):
    d = _lib.ColsumDesc()
    d.n, d.rows = len(entries), B
    for i, (dz, ld, n_cols, part) in enumerate(entries):
        d.dz[i], d.ld[i], d.n_cols[i], d.part[i] = dz, ld, n_cols, part
    return K.Call("pqlb_colsum_partial_multi", C.byref(d), keep=(d, keep))


class _UpdateBase:
    _pf_samples = None         # prefetching sampler (CriticUpdate only)
    _pf_valid = False

    def _bind_set(self, si):
        pass

    def __init__(self, obs_dim, action_dim, batch, device, distl, num_atoms, v_min, v_max, loss_ring=None):
        self.O, self.A, self.B = int(obs_dim), int(action_dim), int(batch)
        self.device = torch.device(device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().pqlb_init(), "pqlb_init")
        self.loss_ring = loss_ring if loss_ring is not None else torch.zeros(5, device=self.device)
        self.graphs = None
        self.graph_launches = 0
        self._bufs = []
        self.distl = bool(distl)
        self.N = int(num_atoms) if distl else 1
        self.v_min, self.v_max = float(v_min), float(v_max)
        self.x_ld = _ru(self.O + self.A, 4)
        self.a_ld = _ru(self.A, 4)
        self.nblk = (self.B + 127) // 128
        self.nblk_head = (self.B + 63) // 64        # head_loss_kernel: 64 rows of one net per block
        self.Lc = NetLayout(self.O + self.A, self.N, 2)
        self.La = NetLayout(self.O, self.A, 1)
        self.pd = 64 if distl else 0           # row stride of the probability / dlogit buffers
        self.calls = []
        self.rng_state = None       # fused sampler RNG: int64 [seed, base_offset, increment] on the device
        self.z = torch.linspace(v_min, v_max, self.N, device=self.device) if distl else None

    def _buf(self, *shape):
        """Zero-initialised fp32 device buffer that lives as long as the plan: the prepared
        launches hold raw addresses, so every buffer must stay referenced (a freed block would be
        handed back to the driver by the empty_cache() that precedes a CUDA-graph capture)."""
        t = torch.zeros(*shape, dtype=torch.float32, device=self.device)
        self._bufs.append(t)
        return t

    def enable_fused_rng(self, generator, draws_per_update):
        """Draw the update's random numbers inside the gather kernel (csrc/rng.cuh) from
        ``generator``'s Philox stream: update k uses offset = generator offset now + 4 * draws * (k -
        completed updates now), i.e. exactly the values torch.randint / normal_ would return if they
        were called with this generator once per update."""
        inc = 4 * int(draws_per_update)
        seed = int(generator.initial_seed())
        if seed >= 1 << 63:
            seed -= 1 << 64
        base = int(generator.get_offset()) - inc * self.opt.step
        self.rng_state = torch.tensor([seed, base, inc], dtype=torch.int64, device=self.device)

    def _ws_init(self, wgrads, n_nets, bias_cols, extra=0):
        """One workspace for every split-K / per-block partial sum of the update, sized up front
        so that all addresses are fixed while the launch list is being prepared."""
        total = extra + 64
        self._wg_multi = WGRAD_MULTI and len(wgrads) * n_nets <= _lib.MAX_WGRAD
        self._wg_pending = []
        if self._wg_multi:
            # one launch for every weight gradient of the update: split counts balanced over one wave of CTAs
            plan = K.wgrad_plan([(n_out, n_in) for n_out, n_in, _ in wgrads for _ in range(n_nets)], self.B)
            self._wg_plan = {}
            for j, (n_out, n_in, ldw) in enumerate(wgrads):
                self._wg_plan[(n_out, n_in)] = plan[j * n_nets: (j + 1) * n_nets]
                total += sum(_ru(s * n_out * ldw, 32) for _, s in plan[j * n_nets: (j + 1) * n_nets])
        else:
            for n_out, n_in, ldw in wgrads:
                _, splits = K.wgrad_tiling(n_out, n_in, self.B, n_nets)
                total += n_nets * _ru(splits * n_out * ldw, 32)
        total += sum(_ru(self.nblk * c, 32) for c in bias_cols)
        self.ws = self._buf(total)
        self._ws_used = 0
        self._ws_log, self._ws_replay = [], None

    def _ws_alloc(self, n):
        if self._ws_replay is not None:       # second build pass: the same launches on the other input set share every buffer
            off, n0 = self._ws_replay.pop(0)
            assert n0 == n, "build passes diverged"
            return off
        off = self._ws_used
        self._ws_used = _ru(off + n, 32)
        assert self._ws_used <= self.ws.numel(), "workspace under-sized"
        self._ws_log.append((off, n))
        return off

    # ---- pieces shared by both updates -----------------------------------------------------
    def _dgrad_chain(self, nets, dz3, dz2, dz1, h1, h2, bias=None):
        """dgrad through layers 3 and 2 of len(nets) networks in ONE launch (pqlb_mlp_backward): dz2
        stays in tensor memory, weights are read as [K][N] (no transposes).  ``bias`` = (opt, layout,
        net ids): the kernel also writes the per-128-row partial column sums of dz2 / dz1, i.e. the
        bias gradients of layers 1 and 0, registered here as reduction sources."""
        groups = []
        for j, n in enumerate(nets):
            g = dict(dz3=K.addr(dz3[j]), w3=n.W[2], w2=n.W[1], h2=K.addr(h2[j]), h1=K.addr(h1[j]), dz2=K.addr(dz2[j]),
                     dz1=K.addr(dz1[j]))
            if bias is not None:
                opt, layout, ids = bias
                for layer, key, n_cols in ((1, "bias_part2", H2), (0, "bias_part1", H1)):
                    off = self._ws_alloc(self.nblk * n_cols)
                    g[key] = K.addr(self.ws, off)
                    opt.add_source(layout.b_off[ids[j]][layer], n_cols, off, n_cols, self.nblk)
            groups.append(g)
        return [K.MlpBackward(self.B, groups)]

    def _head_backward_c51(self, nets, dl, h3, dz3):
        g = [dict(a=K.addr(dl[i]), lda=self.pd, b=nets[i].W[3], ldb=H3, aux=K.addr(h3[i]), ldaux=H3,
                  out=K.addr(dz3[i]), ldo=H3) for i in range(2)]
        return K.Gemm(self.B, H3, self.N, g, epilogue=K.EPI_MUL_ELUGRAD, tile_n=128, b_major=K.MN_MAJOR)

    def _wgrad(self, opt, layout, net_ids, layer, dz, ldz, n_out, h, ldh, n_in):
        """dW[layer] of len(net_ids) nets = dz^T . h (contraction over the batch, split-K partials
        in the workspace; grad_reduce sums them in a fixed order)."""
        B = self.B
        if self._wg_multi:
            # recorded now, launched by _wgrad_flush() once every dz of the update exists
            ldw = layout.ldw[layer]
            stride = n_out * ldw
            for j, i in enumerate(net_ids):
                tile_n, splits = self._wg_plan[(n_out, n_in)][j]
                off = self._ws_alloc(splits * stride)
                self._wg_pending.append(dict(dz=K.addr(dz[j]), lddz=ldz, h=K.addr(h[j]), ldh=ldh, part=K.addr(self.ws, off),
                                             ldo=ldw, split_stride=stride, M=n_out, N=n_in, tile_n=tile_n, splits=splits))
                opt.add_source(layout.w_off[i][layer], stride, off, stride, splits)
            return None
        tile_n, splits = K.wgrad_tiling(n_out, n_in, B, len(net_ids))
        cluster = K.wgrad_cluster(splits, tile_n)       # the splits of a cluster leave the SMs as one partial
        n_part = splits // cluster
        ldw = layout.ldw[layer]
        stride = n_out * ldw
        groups = []
        for j, i in enumerate(net_ids):
            off = self._ws_alloc(n_part * stride)
            groups.append(dict(a=K.addr(dz[j]), lda=ldz, b=K.addr(h[j]), ldb=ldh, out=K.addr(self.ws, off), ldo=ldw,
                               split_stride=stride))
            opt.add_source(layout.w_off[i][layer], stride, off, stride, n_part)
        return K.Gemm(n_out, n_in, B, groups, epilogue=K.EPI_STORE, tile_n=tile_n, a_major=K.MN_MAJOR,
                      b_major=K.MN_MAJOR, splits=splits, cluster=cluster)

    def _wgrad_flush(self):
        """The one pqlb_wgrad_multi launch for the problems recorded by _wgrad (or nothing)."""
        if not self._wg_pending:
            return []
        call = K.WgradMulti(self.B, self._wg_pending)
        self.wgrad_ctas = sum(-(-g["M"] // 128) * -(-g["N"] // g["tile_n"]) * g["splits"] for g in self._wg_pending)
        self.wgrad_partial_bytes = 4 * sum(g["splits"] * g["split_stride"] for g in self._wg_pending)
        self._wg_pending = []
        return [call]

    def _bias_grads(self, opt, layout, entries):
        """entries: (net, layer, dz tensor, ld, n_cols).  One launch for all bias gradients."""
        packed = []
        for i, layer, dz, ld, n_cols in entries:
            off = self._ws_alloc(self.nblk * n_cols)
            packed.append((K.addr(dz), ld, n_cols, K.addr(self.ws, off)))
            opt.add_source(layout.b_off[i][layer], n_cols, off, n_cols, self.nblk)
        return _colsum_call(self.B, packed, keep=None)


class CriticUpdate(_UpdateBase):
    """pql_v_learner.py:73-115 as a prepared launch list (twin-Q or C51)."""

    def __init__(self, obs_dim, action_dim, batch, device, critic_flat, *, distl=False, num_atoms=51,
                 v_min=-10.0, v_max=10.0, gamma_n=0.99 ** 3, lr=5e-4, tau=0.05, max_grad_norm=0.5,
                 noise_bound=0.2, noise_std=0.8, obs_norm=True, eps=1e-4, world_size=1, loss_ring=None,
                 process_group=None, dp_fused=False, fwd_mode=None):
        super().__init__(obs_dim, action_dim, batch, device, distl, num_atoms, v_min, v_max, loss_ring)
        self.process_group, self.dp_fused = process_group, bool(dp_fused)
        O, A, B, N, x_ld = self.O, self.A, self.B, self.N, self.x_ld
        self.fwd_mode = forward_mode(fwd_mode, O, A)
        split = self.fwd_mode == "f16x3"
        dev = self.device
        self.lr, self.tau, self.max_grad_norm = float(lr), float(tau), max_grad_norm
        self.gamma_n = float(np.float32(gamma_n))
        self.noise_bound, self.eps, self.obs_norm = float(noise_bound), float(eps), bool(obs_norm)
        self.noise_std = float(noise_std)
        self.world_size = int(world_size)

        # parameter arenas: critic (owned by the nn.Module), target, actor copy + TF32 twins
        assert critic_flat.numel() == self.Lc.total and critic_flat.is_cuda
        self.c_flat = critic_flat
        self.t_flat = critic_flat.clone()                       # deepcopy(critic), :47
        self.c_tf, self.t_tf = self._buf(self.Lc.total), self._buf(self.Lc.total)
        self.a_flat, self.a_tf = self._buf(self.La.total), self._buf(self.La.total)
        # split-fp16 operand copies (hi | lo) of the three arenas, kept current by the optimiser kernel / set_actor
        self.c_h = half_arena(self.Lc, dev) if split else None
        self.t_h = half_arena(self.Lc, dev) if split else None
        self.a_h = half_arena(self.La, dev) if split else None
        self.round_weights()
        self.opt = _Optim(self.Lc, dev)

        # batch buffers: TWO input sets when the sampler prefetches (set k % 2 feeds update k while a side branch of the
        # same CUDA graph draws batch k + 1 and runs its target-policy forward into the other set), else one
        self.n_sets = 2 if PREFETCH else 1
        self.sets = []
        for _ in range(self.n_sets):
            st = dict(idx=torch.zeros(B, dtype=torch.int64, device=dev), noise=self._buf(B, A),
                      x_cur=self._buf(B, x_ld), x_tgt=self._buf(B, x_ld),
                      # the same rows without the TF32 operand rounding: inputs of the split-fp16 forward
                      xf_cur=self._buf(B, x_ld) if split else None, xf_tgt=self._buf(B, x_ld) if split else None,
                      reward=self._buf(B), done=self._buf(B))
            self.sets.append(st)
        self._bind_set(0)
        self.cur_set = 0                  # input set of the next update on the prefetching path
        self._pf_valid = False            # set cur_set already holds the next update's batch and target action
        self._pf_graphs = {}
        self._pf_samples = None
        self._side = None
        self.pre_k = torch.zeros(1, dtype=torch.int64, device=dev)     # update index of the batch the side branch draws next
        self.mean, self.var = self._buf(O), torch.ones(O, device=dev)
        ha = [self._buf(B, d) for d in HIDDEN]
        h_t = [[self._buf(B, d) for d in HIDDEN] for _ in range(2)]
        h_c = [[self._buf(B, d) for d in HIDDEN] for _ in range(2)]
        self.h_c = h_c
        self.dz = [[self._buf(B, d) for d in HIDDEN] for _ in range(2)]      # dz[i][l]
        self.q = [self._buf(B) for _ in range(2)]
        self.tq = [self._buf(B) for _ in range(2)]
        self.y = self._buf(B)
        self.loss = self._buf(1)
        self.grad_norm = self._buf(1)
        if distl:
            self.p = [self._buf(B, self.pd) for _ in range(2)]
            self.tp = [self._buf(B, self.pd) for _ in range(2)]
            self.dl = [self._buf(B, self.pd) for _ in range(2)]
            self.target = self._buf(B, N)
            self.loss_part = self._buf((B + 7) // 8)
        else:
            self.loss_part = self._buf(2 * self.nblk_head)

        actor = NetAddrs(self.La, 0, self.a_tf, self.a_flat, self.a_h)
        cnet = [NetAddrs(self.Lc, i, self.c_tf, self.c_flat, self.c_h) for i in range(2)]
        tnet = [NetAddrs(self.Lc, i, self.t_tf, self.t_flat, self.t_h) for i in range(2)]
        wg = [(H3, H2, H2), (H2, H1, H1), (H1, O + A, self.Lc.ldw[0])] + ([(N, H3, H3)] if distl else [])
        self._ws_init(wg, 2, [d for d in HIDDEN] * 2 + ([N, N] if distl else []),
                      extra=2 * _ru(self.nblk_head * (H3 + 1), 32) + 2 * _ru(self.nblk_head * H3, 32))
        # the launch list once per input set: identical launches, every buffer but the inputs shared
        self.policy_calls, self.main_calls = [], []
        for si, st in enumerate(self.sets):
            self.opt.frozen = si > 0
            self._ws_replay = list(self._ws_log) if si > 0 else None
            pol, main = self._build_calls(st, actor, cnet, tnet, ha, h_t, h_c, split)
            self.policy_calls.append(pol)
            self.main_calls.append(main)
        self.opt.frozen, self._ws_replay = False, None
        self.calls = self.policy_calls[0] + self.main_calls[0]       # one update, serially, on input set 0
        # the target nets only run forward: in split-fp16 mode nobody reads their TF32 twin, so the optimiser does not write it
        _finish_plan(self, self.opt, self.c_flat, self.t_flat, self.c_tf, None if split else self.t_tf, self.c_h, self.t_h)

    def _bind_set(self, si):
        """``idx`` / ``noise`` / ``x_cur`` ... name the input set of the update that ran last (tests and the loop
        observer read the draws there)."""
        for k, v in self.sets[si].items():
            setattr(self, k, v)

    def _build_calls(self, st, actor, cnet, tnet, ha, h_t, h_c, split):
        """(target-policy launches, everything from the critics' forward to the last weight gradient) on input set ``st``."""
        O, A, B, N, x_ld, distl = self.O, self.A, self.B, self.N, self.x_ld, self.distl
        x_cur, x_tgt, xf_cur, xf_tgt = st["x_cur"], st["x_tgt"], st["xf_cur"], st["xf_tgt"]
        calls = []
        # -- target policy: a' = clamp(tanh(actor(next_obs)) + clamp(noise), +-1)   :62-71, noise.py:19-27
        a_inst = dict(net=actor, x=K.addr(x_tgt), xf=K.addr(xf_tgt), x_ld=x_ld, k_in=O, h=[K.addr(t) for t in ha],
                      store=(False, False, False), terms=1,
                      act=dict(out=K.addr(x_tgt, O), ldo=x_ld, noise=K.addr(st["noise"]), ldnoise=A,
                               noise_std=self.noise_std, noise_bound=self.noise_bound))
        if split:       # the un-rounded action goes next to the un-rounded next_obs
            a_inst["act"].update(out2=K.addr(xf_tgt, O), ldo2=x_ld)
        policy = forward_calls(B, [a_inst], False)
        # -- both target nets on (next_obs, a') and both current nets on (obs, action), one launch
        # C51 + split-fp16 forward: the 51-atom softmax head rides in the fused launch (three MMAs per product
        # like the trunk: tools/precision_study.py - the head's forward is the site the actor gradient needs)
        fused_sm = distl and split
        insts = [dict(net=tnet[i], x=K.addr(x_tgt), xf=K.addr(xf_tgt), x_ld=x_ld, k_in=O + A,
                      h=[K.addr(t) for t in h_t[i]], store=(False, False, distl and not fused_sm), q=K.addr(self.tq[i]), terms=3)
                 for i in range(2)]
        insts += [dict(net=cnet[i], x=K.addr(x_cur), xf=K.addr(xf_cur), x_ld=x_ld, k_in=O + A,
                       h=[K.addr(t) for t in h_c[i]], store=(True, True, True), q=K.addr(self.q[i]), terms=3)
                  for i in range(2)]
        if fused_sm:
            for j, it in enumerate(insts):
                it["softmax"] = dict(out=K.addr(self.tp[j] if j < 2 else self.p[j - 2]), ldp=self.pd)
        # the current critics (which store their activations: the longer tiles) are dispatched first
        calls += forward_calls(B, insts[2:] + insts[:2], not distl)
        if not distl:
            ws_head = [self._ws_alloc(self.nblk_head * (H3 + 1)) for _ in range(2)]
            for i in range(2):
                self.opt.add_source(self.Lc.w_off[i][3], H3, ws_head[i], H3 + 1, self.nblk_head)
                self.opt.add_source(self.Lc.b_off[i][3], 1, ws_head[i] + H3, H3 + 1, self.nblk_head)
            b3 = [self._ws_alloc(self.nblk_head * H3) for _ in range(2)]       # layer-3 bias gradient partials
            for i in range(2):
                self.opt.add_source(self.Lc.b_off[i][2], H3, b3[i], H3, self.nblk_head)
            calls.append(K.Call("pqlb_doubleq_td_loss_b3", _lib.ptr(self.q[0]), _lib.ptr(self.q[1]), _lib.ptr(self.tq[0]),
                                _lib.ptr(self.tq[1]), _lib.ptr(st["reward"]), _lib.ptr(st["done"]), self.gamma_n, B,
                                _lib.ptr(h_c[0][2]), _lib.ptr(h_c[1][2]), C.c_void_p(cnet[0].Wf[3]),
                                C.c_void_p(cnet[1].Wf[3]), _lib.ptr(self.dz[0][2]), _lib.ptr(self.dz[1][2]),
                                _lib.ptr(self.y), C.c_void_p(K.addr(self.ws, ws_head[0])),
                                C.c_void_p(K.addr(self.ws, ws_head[1])), _lib.ptr(self.loss_part),
                                C.c_void_p(K.addr(self.ws, b3[0])), C.c_void_p(K.addr(self.ws, b3[1]))))
            self.loss_scale, self.n_loss_part = 1.0 / B, 2 * self.nblk_head
        else:
            if not fused_sm:
                groups = [dict(a=it["h"][2], lda=H3, b=it["net"].W[3], ldb=H3, bias=it["net"].b[3],
                               out=K.addr(self.tp[j] if j < 2 else self.p[j - 2]), ldo=self.pd)
                          for j, it in enumerate(insts)]
                calls.append(K.Gemm(B, N, H3, groups, epilogue=K.EPI_BIAS_SOFTMAX, tile_n=64))
            calls.append(K.Call("pqlb_c51_td_loss", _lib.ptr(self.p[0]), _lib.ptr(self.p[1]), _lib.ptr(self.tp[0]),
                                _lib.ptr(self.tp[1]), self.pd, _lib.ptr(st["reward"]), _lib.ptr(st["done"]),
                                _lib.ptr(self.z), self.gamma_n, self.v_min, self.v_max, N, B, _lib.ptr(self.target),
                                _lib.ptr(self.dl[0]), _lib.ptr(self.dl[1]), self.pd, _lib.ptr(self.loss_part)))
            calls.append(self._head_backward_c51(cnet, self.dl, [h_c[i][2] for i in range(2)],
                                                 [self.dz[i][2] for i in range(2)]))
            calls += _some(self._wgrad(self.opt, self.Lc, [0, 1], 3, self.dl, self.pd, N,
                                       [h_c[i][2] for i in range(2)], H3, H3))
            self.loss_scale, self.n_loss_part = 1.0 / (B * N), (B + 7) // 8
        # -- backward of the current nets
        dz3, dz2, dz1 = ([self.dz[i][l] for i in range(2)] for l in (2, 1, 0))
        calls += self._dgrad_chain(cnet, dz3, dz2, dz1, [h_c[i][0] for i in range(2)], [h_c[i][1] for i in range(2)],
                                   bias=(self.opt, self.Lc, [0, 1]))
        calls += _some(self._wgrad(self.opt, self.Lc, [0, 1], 2, dz3, H3, H3, [h_c[i][1] for i in range(2)], H2, H2))
        calls += _some(self._wgrad(self.opt, self.Lc, [0, 1], 1, dz2, H2, H2, [h_c[i][0] for i in range(2)], H1, H1))
        calls += _some(self._wgrad(self.opt, self.Lc, [0, 1], 0, dz1, H1, H1, [x_cur, x_cur], x_ld, O + A))
        calls += self._wgrad_flush()
        # bias gradients: layers 0 / 1 come out of the dgrad chain, layer 2 out of the twin-Q loss kernel
        if distl:
            entries = [(i, 2, self.dz[i][2], H3, H3) for i in range(2)] + [(i, 3, self.dl[i], self.pd, N) for i in range(2)]
            calls.append(self._bias_grads(self.opt, self.Lc, entries))
        return policy, calls

    def round_weights(self):
        with torch.cuda.device(self.device):
            for src, dst, half in ((self.c_flat, self.c_tf, self.c_h), (self.t_flat, self.t_tf, self.t_h),
                                   (self.a_flat, self.a_tf, self.a_h)):
                _operand_copies(src, dst, half)

    def set_actor(self, flat):
        """Weights of the target policy (the actor module the driver sends with every update())."""
        self.a_flat.copy_(flat, non_blocking=True)
        with torch.cuda.device(self.device):
            _operand_copies(self.a_flat, self.a_tf, self.a_h)

    def set_norm(self, normalize_tuple):
        if normalize_tuple is None:
            if self.obs_norm:
                raise ValueError("obs_norm is on but no normalize_tuple was delivered")
            return
        mean, var, eps = normalize_tuple
        if abs(float(eps) - self.eps) > 0:
            raise ValueError(f"normaliser epsilon changed ({eps} vs {self.eps}): rebuild the plan")
        self.mean.copy_(mean.reshape(-1).float(), non_blocking=True)
        self.var.copy_(var.reshape(-1).float(), non_blocking=True)

    def sample_call(self, ring, capacity, cur_capacity_dev=None, si=0, counter=None):
        """The fused gather + normalise + cat launch for a given replay ring into input set ``si``; with the fused
        sampler RNG enabled the same launch also draws the indices and the target-policy noise (``counter``: the
        device word holding the update index the draw belongs to; default: the optimiser's completed-update count)."""
        on = self.obs_norm
        st = self.sets[si]
        args = (_lib.ptr(ring), int(capacity), self.O, self.A, _lib.ptr(st["idx"]),
                self.B, _lib.ptr(self.mean) if on else None, _lib.ptr(self.var) if on else None, self.eps,
                _lib.ptr(st["x_cur"]), _lib.ptr(st["x_tgt"]), self.x_ld, _lib.ptr(st["reward"]), _lib.ptr(st["done"]))
        xf = (_lib.ptr(st["xf_cur"]), _lib.ptr(st["xf_tgt"]))
        if self.rng_state is None:
            return K.Call("pqlb_sample_critic_batch", *args, *xf, keep=(ring,))
        cnt = self.opt.count if counter is None else counter
        return K.Call("pqlb_sample_critic_batch_rng", *args, _lib.ptr(self.rng_state), _lib.ptr(cnt),
                      _lib.ptr(cur_capacity_dev), _lib.ptr(st["noise"]), st["noise"].numel(), *xf, keep=(ring, cur_capacity_dev, cnt))

    def bind_replay(self, ring, capacity, cur_capacity_dev):
        """Prepare the sampler launches of the prefetching path: per input set one that draws the CURRENT update's
        batch (first update after an exchange) and one that draws the NEXT update's batch from the side branch."""
        self._pf_samples = None
        self._pf_graphs = {}
        self._pf_valid = False
        if self.n_sets == 2 and self.rng_state is not None:
            self._pf_samples = [(self.sample_call(ring, capacity, cur_capacity_dev, si),
                                 self.sample_call(ring, capacity, cur_capacity_dev, si, counter=self.pre_k)) for si in range(2)]
            self._prek_init = K.Call("pqlb_add_i64", _lib.ptr(self.pre_k), _lib.ptr(self.opt.count), 1)
            self._prek_bump = K.Call("pqlb_add_i64", _lib.ptr(self.pre_k), _lib.ptr(self.pre_k), 1)
        call = self.sample_call(ring, capacity, cur_capacity_dev, 0)
        call.prefetchable = True
        return call

    def invalidate_prefetch(self):
        """The ring, the actor or the normaliser changed (update()): the batch drawn ahead is stale; the next
        update draws its own, with the same generator offsets, so the random stream is the reference's."""
        self._pf_valid = False

    def run(self, sample=None, allreduce=None, use_graph=False, graph_allreduce=False):
        """One update: [sample] + forward/backward launches + gradient reduction, the gradient
        all-reduce when data parallel, then clip + AdamW (+ Polyak) and the loss.  With
        ``use_graph`` the launches are captured once into CUDA graphs and replayed (every pointer is
        fixed; the step count and the loss window live on the device): one graph per update, or two
        with the all-reduce issued eagerly between them; ``graph_allreduce`` captures the NCCL
        all-reduce as well (one graph per update also when data parallel - the learner must then own
        its communicator, because two learners replaying on two streams give NCCL no common order).

        Prefetching path (CUDA graphs + fused sampler RNG + one graph per update): the graph of update k
        has a side branch that draws batch k + 1 and runs its target-policy forward into the other input
        set, concurrently with update k's own launches; update k + 1 then starts at the critics' forward."""
        if (use_graph and sample is not None and self._pf_samples is not None and getattr(sample, "prefetchable", False)
                and (allreduce is None or graph_allreduce)):
            return self._run_prefetching(allreduce)
        self._pf_valid = False
        self._bind_set(0)
        if not use_graph:
            self._segment_a(sample)
            if allreduce is not None:
                allreduce(self.opt.grad)
                self.sumsq_call()
            self._segment_b()
            return
        if self.graphs is None or self.graphs[2] is not sample:
            self._capture(sample, allreduce, graph_allreduce)
        # kernels a replay launches without passing through the C ABI's own launch counter
        self.graph_launches += len(self.calls) + (1 if sample is not None else 0) + 2 + (1 if allreduce is not None else 0)
        self.graphs[0].replay()
        if allreduce is not None and not self.graphs[3]:
            allreduce(self.opt.grad)
        self.graphs[1].replay()

    def _run_prefetching(self, allreduce):
        si, fresh = self.cur_set, not self._pf_valid
        key = (fresh, si)
        g = self._pf_graphs.get(key)
        if g is None:
            g = self._pf_graphs[key] = self._capture_prefetching(fresh, si, allreduce)
        n_pol = len(self.policy_calls[si])
        self.graph_launches += ((2 + n_pol) if fresh else 0) + (2 + n_pol) + len(self.main_calls[si]) + 2 + (1 if allreduce is not None else 0)
        g.replay()
        self._bind_set(si)
        self.cur_set, self._pf_valid = 1 - si, True

    def _capture_prefetching(self, fresh, si, allreduce):
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.emit_prefetching(fresh, si, allreduce)
        return g

    def emit_prefetching(self, fresh, si, allreduce=None):
        """The launches of one update on the prefetching path, on torch's current stream, which must be capturing
        (the side branch is a fork / join of that capture).  Also used by the lock-step driver to capture the updates
        of a whole env step into ONE graph (pql_b200/train.py)."""
        if self._side is None:
            self._side = torch.cuda.Stream(self.device)
        main_s, side_s = self._pf_samples[si][0], self._pf_samples[1 - si][1]
        cur = torch.cuda.current_stream(self.device)
        if fresh:
            self._prek_init()                     # pre_k = completed updates + 1: the side branch draws the NEXT batch
            main_s()
            for c in self.policy_calls[si]:
                c()
        for j, c in enumerate(self.main_calls[si]):
            if j == PF_FORK:
                # the side branch: batch k + 1 and its target action, concurrent with the rest of update k
                self._side.wait_stream(cur)
                with torch.cuda.stream(self._side):
                    side_s()
                    for c2 in self.policy_calls[1 - si]:
                        c2()
                    self._prek_bump()
            c()
        self.reduce_call()
        if allreduce is not None:
            allreduce(self.opt.grad)
            self.sumsq_call()
        self._segment_b()
        cur.wait_stream(self._side)

    def block_ready(self):
        """Can the lock-step driver capture this learner's updates into its per-step graph?  (prefetching sampler bound,
        next update on input set 0; single GPU only - the data-parallel exchange inside a two-branch graph has not been
        run on a multi-GPU box, so it is not offered)"""
        return self._pf_samples is not None and self.cur_set == 0 and self.world_size == 1

    def block_done(self, n_updates):
        """Host-side bookkeeping after the driver replayed a graph holding ``n_updates`` updates of this plan (the first
        drew its own batch, the others were fed by the look-ahead)."""
        n_pol = len(self.policy_calls[0])
        self.graph_launches += n_updates * (2 + n_pol + len(self.main_calls[0]) + 2) + 2 + n_pol
        last = (n_updates - 1) % 2
        self._bind_set(last)
        self.cur_set, self._pf_valid = 1 - last, True

    def _segment_a(self, sample):
        if sample is not None:
            sample()
        for c in self.calls:
            c()
        self.reduce_call()

    def _segment_b(self):
        self.adamw_call()
        self.loss_call()

    def _capture(self, sample, allreduce, graph_allreduce):
        torch.cuda.synchronize(self.device)
        data_parallel = allreduce is not None
        fused = data_parallel and graph_allreduce
        ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(ga):
            self._segment_a(sample)
            if fused:
                allreduce(self.opt.grad)
                self.sumsq_call()
            if fused or not data_parallel:
                self._segment_b()
        if data_parallel and not fused:
            with torch.cuda.graph(gb):
                self.sumsq_call()
                self._segment_b()
        else:
            gb = _NoGraph()
        self.graphs = (ga, gb, sample, fused)


class _NoGraph:
    def replay(self):
        pass


class ActorUpdate(_UpdateBase):
    """pql_p_learner.py:47-64 as a prepared launch list: DPG through the frozen twin-Q / C51 critic."""

    def __init__(self, obs_dim, action_dim, batch, device, actor_flat, *, distl=False, num_atoms=51,
                 v_min=-10.0, v_max=10.0, lr=5e-4, max_grad_norm=0.5, obs_norm=True, eps=1e-4, world_size=1,
                 loss_ring=None, process_group=None, dp_fused=False, fwd_mode=None):
        super().__init__(obs_dim, action_dim, batch, device, distl, num_atoms, v_min, v_max, loss_ring)
        self.process_group, self.dp_fused = process_group, bool(dp_fused)
        O, A, B, N, x_ld, a_ld = self.O, self.A, self.B, self.N, self.x_ld, self.a_ld
        self.fwd_mode = forward_mode(fwd_mode, O, A)
        split = self.fwd_mode == "f16x3"
        dev = self.device
        self.lr, self.max_grad_norm = float(lr), max_grad_norm
        self.eps, self.obs_norm = float(eps), bool(obs_norm)
        self.world_size = int(world_size)
        assert actor_flat.numel() == self.La.total and actor_flat.is_cuda
        self.a_flat, self.a_tf = actor_flat, self._buf(self.La.total)
        self.c_flat, self.c_tf = self._buf(self.Lc.total), self._buf(self.Lc.total)
        self.a_h = half_arena(self.La, dev) if split else None
        self.c_h = half_arena(self.Lc, dev) if split else None
        self.round_weights()
        self.opt = _Optim(self.La, dev)

        self.idx = torch.zeros(B, dtype=torch.int64, device=dev)
        self.x = self._buf(B, x_ld)
        self.xf = self._buf(B, x_ld) if split else None      # un-rounded [norm(obs) | action] rows (split-fp16 forward)
        self.mean, self.var = self._buf(O), torch.ones(O, device=dev)
        # the action tanh(actor(obs)) before any operand rounding: its own buffer, or the action columns of xf
        self.act = self.xf[:, O:O + a_ld] if split else self._buf(B, a_ld)
        act_addr, act_ld = (K.addr(self.xf, O), x_ld) if split else (K.addr(self.act), a_ld)
        ha = [self._buf(B, d) for d in HIDDEN]
        h_c = [[self._buf(B, d) for d in HIDDEN] for _ in range(2)]
        self.dzc = [[self._buf(B, d) for d in HIDDEN] for _ in range(2)]
        self.dza = [self._buf(B, d) for d in HIDDEN]
        self.dz_act = self._buf(B, a_ld)
        self.q = [self._buf(B) for _ in range(2)]
        self.loss, self.grad_norm = self._buf(1), self._buf(1)
        if distl:
            self.p = [self._buf(B, self.pd) for _ in range(2)]
            self.dl = [self._buf(B, self.pd) for _ in range(2)]
            self.loss_part = self._buf((B + 7) // 8)
        else:
            self.loss_part = self._buf(2 * self.nblk_head)
        actor = NetAddrs(self.La, 0, self.a_tf, self.a_flat, self.a_h)
        cnet = [NetAddrs(self.Lc, i, self.c_tf, self.c_flat, self.c_h) for i in range(2)]
        calls = self.calls

        self._ws_init([(A, H3, H3), (H3, H2, H2), (H2, H1, H1), (H1, O, self.La.ldw[0])], 1, [*HIDDEN, A])
        # -- action = tanh(actor(obs)) written straight into the critic input rows      :55
        a_inst = dict(net=actor, x=K.addr(self.x), xf=K.addr(self.xf), x_ld=x_ld, k_in=O, h=[K.addr(t) for t in ha], terms=1,
                      act=dict(out=K.addr(self.x, O), ldo=x_ld, out2=act_addr, ldo2=act_ld))
        calls += forward_calls(B, [a_inst], False)
        # -- frozen critic forward                                                     :56
        insts = [dict(net=cnet[i], x=K.addr(self.x), xf=K.addr(self.xf), x_ld=x_ld, k_in=O + A,
                      h=[K.addr(t) for t in h_c[i]], q=K.addr(self.q[i]), terms=3) for i in range(2)]
        dz3 = [self.dzc[i][2] for i in range(2)]
        fused_sm = distl and split
        if fused_sm:
            for j, it in enumerate(insts):
                it["softmax"] = dict(out=K.addr(self.p[j]), ldp=self.pd)
        calls += forward_calls(B, insts, not distl)
        if not distl:
            calls.append(K.Call("pqlb_dpg_loss", _lib.ptr(self.q[0]), _lib.ptr(self.q[1]), B, _lib.ptr(h_c[0][2]),
                                _lib.ptr(h_c[1][2]), C.c_void_p(cnet[0].Wf[3]), C.c_void_p(cnet[1].Wf[3]),
                                _lib.ptr(dz3[0]), _lib.ptr(dz3[1]), _lib.ptr(self.loss_part)))
            self.n_loss_part = 2 * self.nblk_head
        else:
            if not fused_sm:
                groups = [dict(a=it["h"][2], lda=H3, b=it["net"].W[3], ldb=H3, bias=it["net"].b[3],
                               out=K.addr(self.p[j]), ldo=self.pd) for j, it in enumerate(insts)]
                calls.append(K.Gemm(B, N, H3, groups, epilogue=K.EPI_BIAS_SOFTMAX, tile_n=64))
            calls.append(K.Call("pqlb_c51_dpg_loss", _lib.ptr(self.p[0]), _lib.ptr(self.p[1]), self.pd, _lib.ptr(self.z),
                                N, B, _lib.ptr(self.dl[0]), _lib.ptr(self.dl[1]), self.pd, _lib.ptr(self.q[0]),
                                _lib.ptr(self.loss_part)))
            calls.append(self._head_backward_c51(cnet, self.dl, [h_c[i][2] for i in range(2)], dz3))
            self.n_loss_part = (B + 7) // 8
        self.loss_scale = -1.0 / B                                                     # :57  -Q.mean()
        dz2, dz1 = ([self.dzc[i][l] for i in range(2)] for l in (1, 0))
        calls += self._dgrad_chain(cnet, dz3, dz2, dz1, [h_c[i][0] for i in range(2)], [h_c[i][1] for i in range(2)])
        # -- d loss / d action: both critics' layer-1 dgrad summed in one contraction, tanh' fused
        ldw1 = self.Lc.ldw[0]               # row stride of the critics' first-layer weights (= x_ld for every shape so far)
        g = dict(a=K.addr(dz1[0]), lda=H1, a2=K.addr(dz1[1]), lda2=H1, ldb=ldw1, ldb2=ldw1, out=K.addr(self.dz_act),
                 ldo=a_ld)
        if O % 4 == 0:
            g.update(b=cnet[0].W[0] + 4 * O, b2=cnet[1].W[0] + 4 * O, aux=act_addr, ldaux=act_ld)
            calls.append(K.Gemm(B, A, H1, [g], epilogue=K.EPI_MUL_TANHGRAD, tile_n=K.pick_tile_n(A),
                                b_major=K.MN_MAJOR, K2=H1))
        else:   # action columns are not 16-byte aligned inside W1: contract all columns, store the window
            g.update(b=cnet[0].W[0], b2=cnet[1].W[0], aux=act_addr - 4 * O, ldaux=act_ld)
            calls.append(K.Gemm(B, O + A, H1, [g], epilogue=K.EPI_MUL_TANHGRAD, tile_n=256, b_major=K.MN_MAJOR,
                                K2=H1, col_lo=O, col_hi=O + A))
        # -- actor backward
        calls.append(K.Gemm(B, H3, A, [dict(a=K.addr(self.dz_act), lda=a_ld, b=actor.W[3], ldb=H3, aux=K.addr(ha[2]),
                                             ldaux=H3, out=K.addr(self.dza[2]), ldo=H3)],
                            epilogue=K.EPI_MUL_ELUGRAD, tile_n=fwd_tile(B, H3, 1), b_major=K.MN_MAJOR))
        calls += self._dgrad_chain([actor], [self.dza[2]], [self.dza[1]], [self.dza[0]], [ha[0]], [ha[1]],
                                   bias=(self.opt, self.La, [0]))
        calls += _some(self._wgrad(self.opt, self.La, [0], 3, [self.dz_act], a_ld, A, [ha[2]], H3, H3))
        calls += _some(self._wgrad(self.opt, self.La, [0], 2, [self.dza[2]], H3, H3, [ha[1]], H2, H2))
        calls += _some(self._wgrad(self.opt, self.La, [0], 1, [self.dza[1]], H2, H2, [ha[0]], H1, H1))
        calls += _some(self._wgrad(self.opt, self.La, [0], 0, [self.dza[0]], H1, H1, [self.x], x_ld, O))
        calls += self._wgrad_flush()
        entries = [(0, 2, self.dza[2], H3, H3), (0, 3, self.dz_act, a_ld, A)]     # layers 0 / 1: fused into the dgrad chain
        calls.append(self._bias_grads(self.opt, self.La, entries))
        _finish_plan(self, self.opt, self.a_flat, None, self.a_tf, None, self.a_h, None)

    def round_weights(self):
        with torch.cuda.device(self.device):
            for src, dst, half in ((self.a_flat, self.a_tf, self.a_h), (self.c_flat, self.c_tf, self.c_h)):
                _operand_copies(src, dst, half)

    def set_critic(self, flat):
        self.c_flat.copy_(flat, non_blocking=True)
        with torch.cuda.device(self.device):
            _operand_copies(self.c_flat, self.c_tf, self.c_h)

    set_norm = CriticUpdate.set_norm

    def sample_call(self, obsring, capacity, cur_capacity_dev=None):
        on = self.obs_norm
        args = (_lib.ptr(obsring), int(capacity), self.O, _lib.ptr(self.idx), self.B,
                _lib.ptr(self.mean) if on else None, _lib.ptr(self.var) if on else None, self.eps,
                _lib.ptr(self.x), self.x_ld, self.A)
        if self.rng_state is None:
            return K.Call("pqlb_sample_obs_batch", *args, _lib.ptr(self.xf), keep=(obsring,))
        return K.Call("pqlb_sample_obs_batch_rng", *args, _lib.ptr(self.rng_state), _lib.ptr(self.opt.count),
                      _lib.ptr(cur_capacity_dev), _lib.ptr(self.xf), keep=(obsring, cur_capacity_dev))

    run, _segment_a, _segment_b, _capture = (CriticUpdate.run, CriticUpdate._segment_a, CriticUpdate._segment_b,
                                             CriticUpdate._capture)


def _operand_copies(src, tf32, half):
    """Tensor-core operand copies of a parameter arena: the TF32-rounded twin and, when kept, the
    split-fp16 hi | lo copy (what the optimiser kernel maintains after every step)."""
    _lib.call("pqlb_round_tf32", _lib.ptr(src), _lib.ptr(tf32), src.numel())
    if half is not None:
        n = src.numel()
        _lib.call("pqlb_split_f16", _lib.ptr(src), C.c_void_p(half.data_ptr()), C.c_void_p(half.data_ptr() + 2 * n), n)


def _finish_plan(plan, opt, p_flat, t_flat, p_tf, t_tf, p_h=None, t_h=None):
    """Prepare the reduce / clip+AdamW(+Polyak) / loss launches that close an update."""
    opt.finish(plan.device)
    plan.dp = None
    if plan.world_size > 1 and getattr(plan, "dp_fused", False):
        # data parallel without NCCL on the path: the gradient arena lives in symmetric memory and the
        # optimiser kernel does the two-shot all-reduce over NVLink itself (csrc/optim.cu)
        from ._dp import FusedExchange
        try:
            plan.dp = FusedExchange(opt.layout.total, plan.device, getattr(plan, "process_group", None),
                                    split=bool(_os.environ.get("PQLB_DP_SPLIT")) or getattr(plan, "dp_split", False))
            opt.grad = plan.dp.grad
        except Exception as e:      # no peer-mapped memory on this box (every rank fails alike): NCCL all-reduce instead
            import warnings
            warnings.warn(f"fused gradient exchange unavailable ({type(e).__name__}: {e}); falling back to ncclAllReduce")
            plan.dp = None
    big = plan.ws
    max_norm = -1.0 if plan.max_grad_norm is None else float(plan.max_grad_norm)
    tau = getattr(plan, "tau", 0.0)
    plan.adam_scalars = torch.zeros(16, device=plan.device)
    # reduction + loss sum + AdamW bias corrections in one launch; clip + AdamW (+ Polyak) + step
    # count in the second
    plan.reduce_call = K.Call("pqlb_grad_reduce_finish", _lib.ptr(opt.seg_table), opt.n_seg, _lib.ptr(big), _lib.ptr(opt.grad),
                              _lib.ptr(opt.sumsq), _lib.ptr(plan.loss_part), plan.n_loss_part, plan.loss_scale,
                              _lib.ptr(plan.loss), _lib.ptr(opt.count), _lib.ptr(plan.loss_ring), plan.loss_ring.numel(),
                              plan.lr, 0.9, 0.999, 1e-8, 0.01, tau, _lib.ptr(plan.adam_scalars))
    plan.sumsq_call = K.Call("pqlb_grad_sumsq", _lib.ptr(opt.seg_table), opt.n_seg, _lib.ptr(opt.grad),
                             _lib.ptr(opt.sumsq))
    plan.adamw_call = K.Call("pqlb_adamw_polyak_pre", _lib.ptr(p_flat), _lib.ptr(opt.grad), _lib.ptr(opt.m), _lib.ptr(opt.v),
                             _lib.ptr(t_flat), _lib.ptr(p_tf), _lib.ptr(t_tf), _lib.ptr(p_h), _lib.ptr(t_h),
                             opt.layout.total, _lib.ptr(opt.sumsq),
                             opt.n_seg, 1.0 / plan.world_size, max_norm, _lib.ptr(plan.adam_scalars),
                             _lib.ptr(opt.count), _lib.ptr(plan.grad_norm))
    if plan.dp is not None and plan.dp.split:
        # narrow exchange launch, then the ordinary full-width optimiser on the received gradient
        desc = plan.dp.desc()
        exch = K.Call("pqlb_grad_exchange_dp", opt.layout.total, C.byref(desc), keep=(desc, plan.dp))
        sumsq_addr = C.c_void_p(plan.dp.ctl.data_ptr() + 4 * plan.dp.FLAG_WORDS)
        step = K.Call("pqlb_adamw_polyak_pre", _lib.ptr(p_flat), _lib.ptr(plan.dp.red), _lib.ptr(opt.m), _lib.ptr(opt.v),
                      _lib.ptr(t_flat), _lib.ptr(p_tf), _lib.ptr(t_tf), _lib.ptr(p_h), _lib.ptr(t_h), opt.layout.total, sumsq_addr,
                      plan.dp.world * plan.dp.GRID, 1.0 / plan.world_size, max_norm, _lib.ptr(plan.adam_scalars),
                      _lib.ptr(opt.count), _lib.ptr(plan.grad_norm))
        plan.adamw_call = lambda: (exch(), step())
    elif plan.dp is not None:
        desc = plan.dp.desc()
        plan.adamw_call = K.Call("pqlb_adamw_polyak_dp", _lib.ptr(p_flat), _lib.ptr(opt.m), _lib.ptr(opt.v), _lib.ptr(t_flat),
                                 _lib.ptr(p_tf), _lib.ptr(t_tf), _lib.ptr(p_h), _lib.ptr(t_h), opt.layout.total, C.byref(desc), max_norm,
                                 _lib.ptr(plan.adam_scalars), _lib.ptr(opt.count), _lib.ptr(plan.grad_norm),
                                 keep=(desc, plan.dp))
    plan.loss_call = lambda: None
