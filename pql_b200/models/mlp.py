"""MLP actor / twin-Q critics with the reference's class names, constructor signatures, method
names and ``state_dict`` keys (pql/models/mlp.py:15-40,177-203,244-267), executed by the
tcgen05 kernels of libpqlb200.so.

Parameters live in one flat fp32 *arena* per top-level module (layout in include/pqlb200.h:
each ``nn.Linear`` weight stored ``[out, round_up(in, 4)]`` then its bias, 32-word aligned); the
``nn.Parameter``s of the ``nn.Linear`` sub-modules are views of it, so the fused optimiser kernel
updates the module in place and ``state_dict()/load_state_dict()/deepcopy/pickle`` keep working.
``deepcopy``/``pickle``/``.to()`` detach the views; ``_ensure_tied()`` re-attaches them lazily.

Module-level ``forward``/``get_q*`` are inference entry points (no autograd graph): training
goes through ``PQLVLearner.learn`` / ``PQLPLearner.learn``, which compute gradients with their
own backward kernels.
"""
from collections.abc import Sequence

import torch
import torch.nn as nn

from .. import _kernels as K
from .. import _lib

HIDDEN = (512, 256, 128)            # mlp.py:33-34


def _ru(x, m):
    return (x + m - 1) // m * m


class NetLayout:
    """Offsets (fp32 words) of ``n_nets`` MLPs ``in -> 512 -> 256 -> 128 -> out`` in a flat arena,
    in ``nn.Module.parameters()`` order: per net, per layer, weight then bias."""

    def __init__(self, in_dim, out_dim, n_nets=1):
        self.in_dim, self.out_dim, self.n_nets = int(in_dim), int(out_dim), int(n_nets)
        self.dims = [self.in_dim, *HIDDEN, self.out_dim]
        self.n_layers = len(self.dims) - 1
        # row strides: 16-byte rows as fp32; first layers wider than 128 inputs also as fp16 (multiples of 8), which is
        # what the TMA maps of the wide-input split-fp16 forward need (ShadowHand policy: 211 -> 216, critics: 231 -> 232)
        self.ldw = [_ru(d, 8) if d > 128 else _ru(d, 4) for d in self.dims[:-1]]
        off = 0
        self.w_off, self.b_off = [], []
        for _ in range(self.n_nets):
            wo, bo = [], []
            for l in range(self.n_layers):
                off = _ru(off, 32)
                wo.append(off)
                off += self.dims[l + 1] * self.ldw[l]
                off = _ru(off, 32)
                bo.append(off)
                off += self.dims[l + 1]
            self.w_off.append(wo)
            self.b_off.append(bo)
        self.total = _ru(off, 32)

    def tensors(self):
        """(net, layer, kind, offset, count) of every parameter tensor, in parameters() order."""
        for i in range(self.n_nets):
            for l in range(self.n_layers):
                yield i, l, "w", self.w_off[i][l], self.dims[l + 1] * self.ldw[l]
                yield i, l, "b", self.b_off[i][l], self.dims[l + 1]

    def n_params(self):
        return sum(self.dims[l + 1] * (self.dims[l] + 1) for l in range(self.n_layers)) * self.n_nets


class ParamArena:
    def __init__(self, layout, device="cpu"):
        self.layout = layout
        self.flat = torch.zeros(layout.total, dtype=torch.float32, device=device)

    def weight(self, i, l):
        L = self.layout
        o = L.w_off[i][l]
        return self.flat[o:o + L.dims[l + 1] * L.ldw[l]].view(L.dims[l + 1], L.ldw[l])[:, :L.dims[l]]

    def bias(self, i, l):
        L = self.layout
        o = L.b_off[i][l]
        return self.flat[o:o + L.dims[l + 1]]


def half_arena(layout, device):
    """fp16 operand copies of a parameter arena for the split-fp16 forward: 2 * total halves, hi copy
    at [0, total), lo copy at [total, 2 total); half i of a copy <-> element i of the arena
    (pqlb_split_f16 / the optimiser kernel's param_h output)."""
    return torch.zeros(2 * layout.total, dtype=torch.float16, device=device)


class NetAddrs:
    """Device addresses of one net's tensors: ``W`` = TF32-rounded tensor-core operands,
    ``Wf`` / ``b`` = fp32 weights / biases, ``Wh`` / ``Wl`` = fp16 hi / lo operand copies (or None)."""

    def __init__(self, layout, i, flat_tf32, flat_fp32, half=None):
        self.dims, self.ldw = layout.dims, layout.ldw
        self.W = [K.addr(flat_tf32, layout.w_off[i][l]) if flat_tf32 is not None else 0 for l in range(layout.n_layers)]
        self.Wf = [K.addr(flat_fp32, layout.w_off[i][l]) for l in range(layout.n_layers)]
        self.b = [K.addr(flat_fp32, layout.b_off[i][l]) for l in range(layout.n_layers)]
        self.Wh = self.Wl = None
        if half is not None:
            assert half.dtype == torch.float16 and half.numel() == 2 * layout.total
            self.Wh = [half.data_ptr() + 2 * layout.w_off[i][l] for l in range(layout.n_layers)]
            self.Wl = [half.data_ptr() + 2 * (layout.total + layout.w_off[i][l]) for l in range(layout.n_layers)]


import os as _os
_FWD_CTAS = int(_os.environ.get("PQLB_FWD_CTAS", 250))


def fwd_tile(B, N, n_groups):
    """Widest output tile that still gives the grid enough CTAs to fill the SMs."""
    if N <= 64:
        return K.pick_tile_n(N)
    for t in (256, 128, 64):
        if t > N and t // 2 >= N:
            continue
        if ((B + 127) // 128) * ((N + t - 1) // t) * n_groups >= _FWD_CTAS:
            return t
    return 64


def trunk_calls(B, insts, n_hidden_out=3):
    """Prepared launches of the hidden layers (Linear + ELU, mlp.py:15-24) for up to four
    network instances of identical shape.  inst = dict(net=NetAddrs, x=addr, x_ld, k_in,
    h=[addr h1, addr h2, addr h3])."""
    calls = []
    dims = insts[0]["net"].dims
    for l in range(n_hidden_out):
        groups = []
        for it in insts:
            net = it["net"]
            a, lda, k = (it["x"], it["x_ld"], it["k_in"]) if l == 0 else (it["h"][l - 1], dims[l], dims[l])
            groups.append(dict(a=a, lda=lda, b=net.W[l], ldb=net.ldw[l], bias=net.b[l], out=it["h"][l],
                               ldo=dims[l + 1]))
        k_l = insts[0]["k_in"] if l == 0 else dims[l]
        calls.append(K.Gemm(B, dims[l + 1], k_l, groups, epilogue=K.EPI_BIAS_ELU,
                            tile_n=fwd_tile(B, dims[l + 1], len(groups))))
    return calls


FUSED_MAX_IN = 128        # pqlb_mlp_forward keeps a 128-row input tile of at most 128 columns in shared memory
FUSED_H_MAX_IN = 256      # pqlb_mlp_forward_h: up to 128 columns, or 129..256 with its wide-input kernel (ShadowHand critics: 231)


def split_f16_ok(inst):
    """Can this instance's trunk run on the split-fp16 fused forward (pqlb_mlp_forward_h)?  Needs fp16 weight
    copies and an input at most 256 wide whose weight rows are 16-byte aligned as halves."""
    n = inst["net"]
    return n.Wh is not None and inst["k_in"] <= FUSED_H_MAX_IN and n.ldw[0] % 8 == 0


_WIDE_HEAD = _os.environ.get("PQLB_WIDE_HEAD", "1") != "0"      # "0": behind a wide input the policy head is always its own launch


def fused_head_ok(inst):
    """Does the policy head of this instance ride in the fused launch?  A <= 16, A % 4 == 0 and 16-byte aligned
    output rows; behind an input wider than 128 (the wide-input kernel) A <= 32 and output rows of any alignment
    (ShadowHand: 20 actions written behind 211 observations).  Otherwise the trunk stores h3 and the head is its own
    small launch (policy_head_call)."""
    act = inst["act"]
    A = inst["net"].dims[4]
    if A % 4 or act.get("ldnoise", 0) % 4 or act.get("noise", 0) % 16:
        return False
    if inst["k_in"] > FUSED_MAX_IN:
        return A <= 32 and _WIDE_HEAD
    return not (A > 16 or act.get("ldo", 0) % 4 or act.get("ldo2", 0) % 4 or act.get("out", 0) % 16 or act.get("out2", 0) % 16)


def forward_calls_h(B, insts, scalar_head, tile_sync=None):
    """ONE split-fp16 fused launch for up to five network instances (input widths may differ).  inst as
    in ``forward_calls`` plus ``terms`` (3 = hi/lo split of both operands, 1 = hi only), ``xf`` (the
    un-rounded input rows; falls back to ``x``) and optional ``wait_flag`` / ``done_flag`` / ``epoch``."""
    groups, heads_left = [], []
    for it in insts:
        n = it["net"]
        st = it.get("store", (True, True, True))
        terms = int(it.get("terms", 3))
        act = it.get("act")
        if act is not None and not fused_head_ok(it):
            # a head the kernel does not fuse (ShadowHand: 20 actions): the trunk leaves h3, the head is its own launch
            if it.get("publish") or it.get("wait"):
                raise ValueError("tile dependencies need the fused policy head")
            st, act = (st[0], st[1], True), None
            heads_left.append(it)
        g = dict(x=it.get("xf") or it["x"], ldx=it["x_ld"], w1h=n.Wh[0], ldw1=n.ldw[0], w2h=n.Wh[1], w3h=n.Wh[2],
                 b1=n.b[0], b2=n.b[1], b3=n.b[2], terms=terms, k_in=it["k_in"],
                 h1=it["h"][0] if st[0] else 0, h2=it["h"][1] if st[1] else 0, h3=it["h"][2] if st[2] else 0)
        if terms == 3:
            g.update(w1l=n.Wl[0], w2l=n.Wl[1], w3l=n.Wl[2])
        if scalar_head and it.get("q"):
            g.update(head_w=n.Wf[3], head_b=n.b[3], q=it["q"])
        if act is not None:
            g.update(act_wh=n.Wh[3], act_b=n.b[3], act_n=n.dims[4], act_out=act.get("out", 0), act_ldo=act.get("ldo", 0),
                     act_out2=act.get("out2", 0), act_ldo2=act.get("ldo2", 0), act_noise=act.get("noise", 0),
                     act_ldnoise=act.get("ldnoise", 0), noise_std=act.get("noise_std", 0.0),
                     noise_bound=act.get("noise_bound", 0.0))
            if terms == 3:
                g["act_wl"] = n.Wl[3]
        sm = it.get("softmax")          # C51 head fused behind the trunk: dict(out, ldp)
        if sm is not None:
            g.update(sm_wh=n.Wh[3], sm_b=n.b[3], sm_out=sm["out"], sm_ldp=sm["ldp"], sm_n=n.dims[4])
            if terms == 3:
                g["sm_wl"] = n.Wl[3]
        for k in ("publish", "wait"):
            if it.get(k):
                g[k] = 1
        groups.append(g)
    return [K.MlpForwardH(B, max(it["k_in"] for it in insts), groups, tile_sync=tile_sync)] + [policy_head_call(B, it) for it in heads_left]


def forward_calls(B, insts, scalar_head):
    """Prepared launches of the trunk (three Linear+ELU layers) for up to four network instances.
    inst = dict(net, x, x_ld, k_in, h=[h1, h2, h3 addresses], store=(s1, s2, s3), q=addr or 0,
    act=dict(out, ldo, [out2, ldo2], [noise, ldnoise, noise_std, noise_bound]) for a policy net).
    ONE layer-fused launch (activations stay in tensor memory, only the flagged ones are written; the
    scalar twin-Q head or the tanh policy head ride in the same launch) where the shapes allow it: the
    split-fp16 kernel when every instance carries fp16 weight copies and passes ``split_f16_ok`` (inputs up
    to 256 wide: ShadowHand's critics and policy net included), else the TF32 kernel for inputs up to 128
    wide; what is left (inputs wider than 256, wide inputs without fp16 copies: the modules' own forward)
    runs layer by layer."""
    k_in = insts[0]["k_in"]
    calls = []
    if all(split_f16_ok(it) for it in insts):
        return forward_calls_h(B, insts, scalar_head)
    if k_in <= FUSED_MAX_IN:
        groups, heads_left = [], []
        for it in insts:
            n = it["net"]
            st = it.get("store", (True, True, True))
            g = dict(x=it["x"], ldx=it["x_ld"], w1=n.W[0], ldw1=n.ldw[0], w2=n.W[1], w3=n.W[2], b1=n.b[0], b2=n.b[1],
                     b3=n.b[2], h1=it["h"][0] if st[0] else 0, h2=it["h"][1] if st[1] else 0,
                     h3=it["h"][2] if st[2] else 0)
            if scalar_head:
                g.update(head_w=n.Wf[3], head_b=n.b[3], q=it["q"])
            act = it.get("act")
            if act is not None:
                A = n.dims[4]
                if A <= 16 and A % 4 == 0 and act["ldo"] % 4 == 0 and act.get("ldo2", 0) % 4 == 0:
                    g.update(act_w=n.W[3], act_b=n.b[3], act_n=A, act_out=act["out"], act_ldo=act["ldo"],
                             act_out2=act.get("out2", 0), act_ldo2=act.get("ldo2", 0), act_noise=act.get("noise", 0),
                             act_ldnoise=act.get("ldnoise", 0), noise_std=act.get("noise_std", 0.0),
                             noise_bound=act.get("noise_bound", 0.0))
                else:
                    g["h3"] = it["h"][2]          # the separate head launch reads h3
                    heads_left.append(it)
            groups.append(g)
        calls.append(K.MlpForward(B, k_in, groups))
        for it in heads_left:
            calls.append(policy_head_call(B, it))
        return calls
    if not scalar_head:
        calls = trunk_calls(B, insts, 3)
        for it in insts:
            if it.get("act") is not None:
                calls.append(policy_head_call(B, it))
        return calls
    calls = trunk_calls(B, insts, 2)
    groups = []
    for it in insts:
        n = it["net"]
        st = it.get("store", (True, True, True))
        groups.append(dict(a=it["h"][1], lda=HIDDEN[1], b=n.W[2], ldb=n.ldw[2], bias=n.b[2], head_w=n.Wf[3],
                           head_b=n.b[3], q=it["q"], out=it["h"][2] if st[2] else 0, ldo=HIDDEN[2]))
    calls.append(K.Gemm(B, HIDDEN[2], HIDDEN[1], groups, epilogue=K.EPI_BIAS_ELU_HEAD, tile_n=128))
    return calls


def policy_head_call(B, it):
    """tanh(Linear(128, A)) (+ clipped target-policy noise) as its own launch: the path for action
    widths the fused kernel does not take (A > 16 or not a multiple of 4) and for wide inputs."""
    n, act = it["net"], it["act"]
    A, H3 = n.dims[4], HIDDEN[2]
    g = dict(a=it["h"][2], lda=H3, b=n.W[3], ldb=H3, bias=n.b[3], out=act["out"], ldo=act["ldo"])
    if act.get("out2"):
        g.update(out2=act["out2"], ldo2=act["ldo2"])
    if act.get("noise"):
        g.update(aux=act["noise"], ldaux=act["ldnoise"])
        return K.Gemm(B, A, H3, [g], epilogue=K.EPI_BIAS_TANH_NOISE, tile_n=K.pick_tile_n(A),
                      noise_bound=act["noise_bound"], noise_std=act["noise_std"])
    return K.Gemm(B, A, H3, [g], epilogue=K.EPI_BIAS_TANH, tile_n=K.pick_tile_n(A))


class _ArenaModule(nn.Module):
    """Shared arena plumbing of the top-level modules."""

    def _init_arena(self, layout, nets):
        # plain attributes (not buffers): the arena must not appear in state_dict()
        self._layout = layout
        self._nets = nets                    # list of nn.Sequential, arena order
        self._arena = ParamArena(layout, "cpu")
        for i, seq in enumerate(nets):
            for l, lin in enumerate(m for m in seq if isinstance(m, nn.Linear)):
                self._arena.weight(i, l).copy_(lin.weight.data)
                self._arena.bias(i, l).copy_(lin.bias.data)
        self._tie()

    def _linears(self):
        for i, seq in enumerate(self._nets):
            for l, lin in enumerate(m for m in seq if isinstance(m, nn.Linear)):
                yield i, l, lin

    def _tie(self):
        for i, l, lin in self._linears():
            lin.weight.data = self._arena.weight(i, l)
            lin.bias.data = self._arena.bias(i, l)

    def _ensure_tied(self):
        """Re-attach the parameters to the arena after deepcopy / pickle / ``.to()`` (which clone
        every parameter separately).  When detached, the parameters hold the truth."""
        dev = next(self.parameters()).device
        if self._arena.flat.device != dev:
            self._arena.flat = torch.zeros(self._layout.total, dtype=torch.float32, device=dev)
        tied = all(lin.weight.data_ptr() == self._arena.weight(i, l).data_ptr() and
                   lin.bias.data_ptr() == self._arena.bias(i, l).data_ptr() for i, l, lin in self._linears())
        if not tied:
            with torch.no_grad():
                for i, l, lin in self._linears():
                    self._arena.weight(i, l).copy_(lin.weight.data)
                    self._arena.bias(i, l).copy_(lin.bias.data)
            self._tie()
        return self._arena

    def _apply(self, fn, recurse=True):
        super()._apply(fn, recurse)          # moves every parameter separately ...
        self._ensure_tied()                  # ... so gather them back into one arena
        return self

    @property
    def arena(self):
        return self._ensure_tied()

    # ---- inference through the kernels -----------------------------------------------------
    @torch.no_grad()
    def _forward(self, net_ids, a, b=None, head="linear"):
        """Outputs of nets ``net_ids`` on input rows ``cat(a, b)``: list of fp32 tensors."""
        arena = self._ensure_tied()
        L = self._layout
        dev = arena.flat.device
        if dev.type != "cuda":
            raise RuntimeError("pql_b200 models run on CUDA only (there is no CPU or PyTorch fallback): "
                               "move the module and its inputs to a CUDA device")
        a = a.to(device=dev, dtype=torch.float32)
        if a.dim() != 2 or a.stride(1) != 1:
            a = a.reshape(a.shape[0], -1).contiguous()
        na, nb = a.shape[1], 0
        if b is not None:
            b = b.to(device=dev, dtype=torch.float32)
            if b.dim() != 2 or b.stride(1) != 1:
                b = b.reshape(b.shape[0], -1).contiguous()
            nb = b.shape[1]
        if na + nb != L.in_dim:
            raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied ({a.shape[0]}x{na + nb} and "
                               f"{L.in_dim}x{HIDDEN[0]})")
        B = a.shape[0]
        x_ld = L.ldw[0]
        with torch.cuda.device(dev):
            w_tf = torch.empty_like(arena.flat)
            _lib.call("pqlb_round_tf32", _lib.ptr(arena.flat), _lib.ptr(w_tf), arena.flat.numel())
            x = torch.empty((B, x_ld), dtype=torch.float32, device=dev)
            _lib.call("pqlb_pack_x", _lib.ptr(a), a.stride(0), na, _lib.ptr(b), b.stride(0) if nb else 0, nb,
                      _lib.ptr(x), x_ld, B)
            nets = [NetAddrs(L, i, w_tf, arena.flat) for i in net_ids]
            hs = [[torch.empty((B, d), dtype=torch.float32, device=dev) for d in HIDDEN] for _ in net_ids]
            insts = [dict(net=n, x=K.addr(x), x_ld=x_ld, k_in=L.in_dim, h=[K.addr(t) for t in h])
                     for n, h in zip(nets, hs)]
            out_dim = L.out_dim
            outs = []
            if head == "linear" and out_dim == 1:
                for c in trunk_calls(B, insts, 2):
                    c()
                qs = [torch.empty((B, 1), dtype=torch.float32, device=dev) for _ in net_ids]
                groups = [dict(a=it["h"][1], lda=HIDDEN[1], b=n.W[2], ldb=n.ldw[2], bias=n.b[2], head_w=n.Wf[3],
                               head_b=n.b[3], q=K.addr(q)) for it, n, q in zip(insts, nets, qs)]
                K.Gemm(B, HIDDEN[2], HIDDEN[1], groups, epilogue=K.EPI_BIAS_ELU_HEAD, tile_n=128)()
                return qs
            for c in trunk_calls(B, insts, 3):
                c()
            if head == "softmax" and out_dim > 64:
                raise NotImplementedError("softmax head supports at most 64 atoms")
            if out_dim > 256:
                raise NotImplementedError("output layers wider than 256 are not on the PQL path")
            tile = K.pick_tile_n(out_dim) if head != "softmax" else max(32, K.pick_tile_n(out_dim))
            scratch = torch.empty((B, _ru(out_dim, 4)), dtype=torch.float32, device=dev) if head == "tanh" else None
            for it, n in zip(insts, nets):
                o = torch.empty((B, out_dim), dtype=torch.float32, device=dev)
                g = dict(a=it["h"][2], lda=HIDDEN[2], b=n.W[3], ldb=n.ldw[3], bias=n.b[3])
                if head == "tanh":
                    g.update(out=K.addr(scratch), ldo=scratch.shape[1], out2=K.addr(o), ldo2=out_dim)
                    epi = K.EPI_BIAS_TANH
                else:
                    g.update(out=K.addr(o), ldo=out_dim)
                    epi = K.EPI_BIAS_SOFTMAX if head == "softmax" else K.EPI_BIAS
                K.Gemm(B, out_dim, HIDDEN[2], [g], epilogue=epi, tile_n=tile)()
                outs.append(o)
            return outs


def _make_seq(in_dim, out_dim):
    """create_simple_mlp (mlp.py:15-24) with the default hidden sizes: same construction order,
    hence the same seeded default initialisation as the reference."""
    dims = [in_dim, *HIDDEN, out_dim]
    mods = []
    for idx, (i, o) in enumerate(zip(dims[:-1], dims[1:])):
        mods.append(nn.Linear(i, o))
        if idx < len(dims) - 2:
            mods.append(nn.ELU())
    return nn.Sequential(*mods)


def _check_mlp_args(hidden_layers, use_batchnorm):
    if use_batchnorm:
        raise NotImplementedError("use_batchnorm is not on the PQL path (DoubleQBatchNorm is a CrossQ baseline)")
    if hidden_layers is not None and tuple(hidden_layers) != HIDDEN:
        raise NotImplementedError(f"the kernels are specialised to hidden_layers={list(HIDDEN)} (mlp.py:33-34)")


class MLPNet(_ArenaModule):
    """mlp.py:27-40."""

    def __init__(self, in_dim, out_dim, hidden_layers=None, use_batchnorm=False, _own_arena=True):
        super().__init__()
        if isinstance(in_dim, Sequence):
            in_dim = in_dim[0]
        _check_mlp_args(hidden_layers, use_batchnorm)
        self.net = _make_seq(int(in_dim), int(out_dim))
        self._own = bool(_own_arena)
        if self._own:
            self._init_arena(NetLayout(in_dim, out_dim, 1), [self.net])

    def _apply(self, fn, recurse=True):
        if self._own:
            return super()._apply(fn, recurse)
        return nn.Module._apply(self, fn, recurse)

    def _ensure_tied(self):
        if not self._own:
            raise RuntimeError("this MLPNet is a sub-network of a twin-Q critic: call the critic's methods")
        return super()._ensure_tied()

    def forward(self, x):
        return self._forward([0], x)[0]


class TanhMLPPolicy(MLPNet):
    """mlp.py:177-179."""

    def forward(self, state):
        return self._forward([0], state, head="tanh")[0]


class DoubleQ(_ArenaModule):
    """mlp.py:186-203."""
    _OUT = 1

    def __init__(self, state_dim, act_dim):
        super().__init__()
        if isinstance(state_dim, Sequence):
            state_dim = state_dim[0]
        self._build(int(state_dim), int(act_dim), 1)

    def _build(self, state_dim, act_dim, out_dim):
        self.state_dim, self.act_dim = state_dim, act_dim
        self.net_q1 = MLPNet(in_dim=state_dim + act_dim, out_dim=out_dim, _own_arena=False)
        self.net_q2 = MLPNet(in_dim=state_dim + act_dim, out_dim=out_dim, _own_arena=False)
        self._init_arena(NetLayout(state_dim + act_dim, out_dim, 2), [self.net_q1.net, self.net_q2.net])

    def get_q_min(self, state, action):
        q1, q2 = self._forward([0, 1], state, action)
        return torch.min(q1, q2)

    def get_q1_q2(self, state, action):
        q1, q2 = self._forward([0, 1], state, action)
        return q1, q2

    def get_q1(self, state, action):
        return self._forward([0], state, action)[0]


class DistributionalDoubleQ(DoubleQ):
    """mlp.py:244-267: 51-atom softmax heads; ``z_atoms`` is a plain tensor attribute."""

    def __init__(self, state_dim, act_dim, v_min=-10, v_max=10, num_atoms=51, device="cuda"):
        nn.Module.__init__(self)
        if isinstance(state_dim, Sequence):
            state_dim = state_dim[0]
        if num_atoms > 64:
            raise NotImplementedError("the C51 kernels keep one distribution per warp: num_atoms <= 64")
        self.device = device
        self.v_min, self.v_max, self.num_atoms = v_min, v_max, int(num_atoms)
        self._build(int(state_dim), int(act_dim), int(num_atoms))
        self.z_atoms = torch.linspace(v_min, v_max, num_atoms, device=device)

    def get_q_min(self, state, action):
        p1, p2 = self._forward([0, 1], state, action, head="softmax")
        B = p1.shape[0]
        q = torch.empty(B, dtype=torch.float32, device=p1.device)
        z = self.z_atoms.to(device=p1.device, dtype=torch.float32).contiguous()
        with torch.cuda.device(p1.device):
            _lib.call("pqlb_c51_dpg_loss", _lib.ptr(p1), _lib.ptr(p2), self.num_atoms, _lib.ptr(z), self.num_atoms,
                      B, None, None, 0, _lib.ptr(q), None)
        return q

    def get_q1_q2(self, state, action):
        p1, p2 = self._forward([0, 1], state, action, head="softmax")
        return p1, p2

    def get_q1(self, state, action):
        return self._forward([0], state, action, head="softmax")[0]
