"""Prepared launches of the C-ABI kernels (include/pqlb200.h).

A learner builds its whole update once as a list of prepared calls over preallocated buffers
(all device pointers fixed), then ``learn()`` just replays the list - on torch's current
stream, so the same list can be captured into a CUDA graph.  Nothing here computes anything
on the host or in PyTorch: every object is an argument pack for one ``extern "C"`` entry point.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import (EPI_BIAS, EPI_BIAS_ELU, EPI_BIAS_ELU_HEAD, EPI_BIAS_SOFTMAX, EPI_BIAS_TANH,  # noqa: F401
                   EPI_BIAS_TANH_NOISE, EPI_MUL_ELUGRAD, EPI_MUL_TANHGRAD, EPI_STORE, K_MAJOR,
                   MN_MAJOR)


def addr(t, off=0):
    """Device address of element ``off`` (fp32 words) of tensor ``t`` (None -> 0)."""
    if t is None:
        return 0
    if isinstance(t, int):
        return t + 4 * off
    return t.data_ptr() + t.element_size() * off


def _p(a):
    return C.c_void_p(a if a else None)


import os as _os
TRACE = bool(_os.environ.get("PQLB_TRACE"))     # debug: synchronise after every launch and name the failing one
PROFILE = None      # bench.py sets this to {} to collect (start, end) CUDA events per entry point


class Call:
    """One prepared C-ABI call: ``Call(name, *args)`` -> ``call()`` launches on the current stream."""

    def __init__(self, name, *args, keep=()):
        self.name, self.args, self._keep = name, args, keep

    def __call__(self):
        if TRACE:
            _lib.call(self.name, *self.args)
            try:
                torch.cuda.synchronize()
            except Exception:
                d = getattr(self, "desc", None)
                if d is not None:
                    print(f"[pqlb trace] FAILED {self.name}: M={d.M} N={d.N} K={d.K} K2={d.K2} a_major={d.a_major} "
                          f"b_major={d.b_major} epi={d.epilogue} tile_n={d.tile_n} splits={d.splits} groups={d.n_groups} "
                          f"col=[{d.col_lo},{d.col_hi})", flush=True)
                else:
                    print(f"[pqlb trace] FAILED {self.name}", flush=True)
                raise
            return
        if PROFILE is None:
            _lib.call(self.name, *self.args)
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call(self.name, *self.args)
        e1.record()
        PROFILE.setdefault(self.name, []).append((e0, e1))


class Gemm(Call):
    """Grouped ``D[M,N] = epilogue(A . B)`` on tcgen05 (pqlb_gemm_tf32).

    ``groups``: list of dicts with keys a, lda, b, ldb, [a2, lda2, b2, ldb2], bias, aux, ldaux,
    head_w, head_b, q, out, ldo, out2, ldo2, split_stride; pointer values are device addresses
    (``addr(tensor, word_offset)``)."""

    PTRS = ("a", "b", "a2", "b2", "bias", "aux", "head_w", "head_b", "q", "out", "out2")
    INTS = ("lda", "ldb", "lda2", "ldb2", "ldaux", "ldo", "ldo2", "split_stride")

    def __init__(self, M, N, K, groups, *, epilogue, tile_n, a_major=K_MAJOR, b_major=K_MAJOR,
                 splits=1, K2=0, col_lo=0, col_hi=0, noise_bound=0.0, noise_std=1.0, cluster=1, keep=()):
        d = _lib.GemmDesc()
        d.M, d.N, d.K, d.K2 = int(M), int(N), int(K), int(K2)
        d.a_major, d.b_major, d.epilogue, d.tile_n = a_major, b_major, epilogue, int(tile_n)
        d.splits, d.n_groups, d.cluster = int(splits), len(groups), int(cluster)
        d.col_lo, d.col_hi, d.noise_bound, d.noise_std = int(col_lo), int(col_hi), float(noise_bound), float(noise_std)
        if not 1 <= len(groups) <= _lib.MAX_GROUPS:
            raise ValueError("1..%d groups per launch" % _lib.MAX_GROUPS)
        for i, g in enumerate(groups):
            unknown = set(g) - set(self.PTRS) - set(self.INTS)
            if unknown:
                raise KeyError(f"unknown gemm group fields {sorted(unknown)}")
            for k in self.PTRS:
                setattr(d.g[i], k, g.get(k, 0) or None)
            for k in self.INTS:
                setattr(d.g[i], k, int(g.get(k, 0)))
        self.desc = d
        super().__init__("pqlb_gemm_tf32", C.byref(d), keep=keep)


class MlpForward(Call):
    """Layer-fused trunk forward (pqlb_mlp_forward) for up to four network instances.
    ``groups``: dicts with x, ldx, w1, ldw1, w2, w3, b1, b2, b3, [head_w, head_b, q], [h1, h2, h3]."""

    FIELDS = ("x", "w1", "w2", "w3", "b1", "b2", "b3", "head_w", "head_b", "q", "h1", "h2", "h3",
              "act_w", "act_b", "act_noise", "act_out", "act_out2")
    INTS = ("act_ldo", "act_ldo2", "act_ldnoise", "act_n")
    FLOATS = ("noise_std", "noise_bound")

    def __init__(self, M, k_in, groups):
        d = _lib.MlpDesc()
        d.M, d.k_in, d.n_groups = int(M), int(k_in), len(groups)
        for i, g in enumerate(groups):
            unknown = set(g) - set(self.FIELDS) - set(self.INTS) - set(self.FLOATS) - {"ldx", "ldw1"}
            if unknown:
                raise KeyError(f"unknown mlp group fields {sorted(unknown)}")
            for k in self.FIELDS:
                setattr(d.g[i], k, g.get(k, 0) or None)
            for k in self.INTS:
                setattr(d.g[i], k, int(g.get(k, 0)))
            for k in self.FLOATS:
                setattr(d.g[i], k, float(g.get(k, 0.0)))
            d.g[i].ldx, d.g[i].ldw1 = int(g["ldx"]), int(g["ldw1"])
        self.desc = d
        super().__init__("pqlb_mlp_forward", C.byref(d))


class MlpForwardH(Call):
    """Layer-fused trunk forward with split-fp16 operands (pqlb_mlp_forward_h) for up to five network
    instances.  ``groups``: dicts with x (fp32, un-rounded), ldx, w1h/w1l, ldw1 (halves), w2h/w2l, w3h/w3l,
    b1..b3, terms (1 | 3), [head_w, head_b, q], [h1, h2, h3], [act_wh, act_wl, act_b, ...], [sm_*],
    [publish | wait] (tile dependencies through ``tile_sync``: int32 [2 + row tiles], zero-initialised)."""

    FIELDS = ("x", "w1h", "w1l", "w2h", "w2l", "w3h", "w3l", "b1", "b2", "b3", "head_w", "head_b", "q", "h1", "h2", "h3",
              "act_wh", "act_wl", "act_b", "act_noise", "act_out", "act_out2", "sm_wh", "sm_wl", "sm_b", "sm_out")
    INTS = ("ldx", "ldw1", "act_ldo", "act_ldo2", "act_ldnoise", "act_n", "terms", "k_in", "sm_ldp", "sm_n", "publish", "wait")
    FLOATS = ("noise_std", "noise_bound")

    def __init__(self, M, k_in, groups, tile_sync=None):
        d = _lib.MlpHDesc()
        d.M, d.k_in, d.n_groups = int(M), int(k_in), len(groups)
        d.tile_sync = tile_sync.data_ptr() if tile_sync is not None else None
        self._tile_sync = tile_sync
        if not 1 <= len(groups) <= _lib.MAX_FWD_GROUPS:
            raise ValueError("1..%d groups per launch" % _lib.MAX_FWD_GROUPS)
        for i, g in enumerate(groups):
            unknown = set(g) - set(self.FIELDS) - set(self.INTS) - set(self.FLOATS)
            if unknown:
                raise KeyError(f"unknown mlp group fields {sorted(unknown)}")
            for k in self.FIELDS:
                setattr(d.g[i], k, g.get(k, 0) or None)
            for k in self.INTS:
                setattr(d.g[i], k, int(g.get(k, 0)))
            for k in self.FLOATS:
                setattr(d.g[i], k, float(g.get(k, 0.0)))
        self.desc = d
        super().__init__("pqlb_mlp_forward_h", C.byref(d))


class MlpBackward(Call):
    """Layer-fused dgrad chain (pqlb_mlp_backward) for up to four network instances.
    ``groups``: dicts with dz3, w3, w2, h2, h1, dz2, dz1, [bias_part2, bias_part1] device addresses."""

    FIELDS = ("dz3", "w3", "w2", "h2", "h1", "dz2", "dz1", "bias_part2", "bias_part1")

    def __init__(self, M, groups):
        d = _lib.MlpBwdDesc()
        d.M, d.n_groups = int(M), len(groups)
        for i, g in enumerate(groups):
            unknown = set(g) - set(self.FIELDS)
            if unknown:
                raise KeyError(f"unknown mlp backward group fields {sorted(unknown)}")
            for k in self.FIELDS:
                setattr(d.g[i], k, g.get(k, 0) or None)
        self.desc = d
        super().__init__("pqlb_mlp_backward", C.byref(d))


def pick_tile_n(N):
    for t in (16, 32, 64, 128, 256):
        if N <= t:
            return t
    return 256


def wgrad_tiling(M, N, K, n_groups, target_ctas=int(_os.environ.get("PQLB_WGRAD_CTAS", 296))):
    """Tile width and split-K factor of a weight-gradient GEMM (contraction over the batch):
    the smallest split count that still fills the SMs, with kb_total % splits == 0."""
    kb = (K + 31) // 32
    best = None
    for tile_n in (256, 128, 64):
        if tile_n > 64 and N <= tile_n // 2:
            continue
        tiles = ((M + 127) // 128) * ((N + tile_n - 1) // tile_n) * n_groups
        for splits in (1, 2, 4, 8, 16, 32, 64):
            if kb % splits:
                continue
            ctas = tiles * splits
            if ctas >= target_ctas * 0.8 or splits == 64 or kb // splits <= 2:
                cand = (splits, -tile_n)
                if best is None or cand < best[0]:
                    best = (cand, tile_n, splits)
                break
    if best is None:
        return pick_tile_n(min(N, 256)), 1
    return best[1], best[2]


def wgrad_tile_n(N):
    return 256 if N > 128 else 128 if N > 64 else 64 if N > 32 else 32


def wgrad_plan(problems, K, target_ctas=None):
    """Split-K factors of the weight-gradient problems of ONE launch (pqlb_wgrad_multi).

    ``problems``: [(M, N)] = (output features, input features) of every dW = dz^T . h of the update.
    The launch is one wave of at most ``target_ctas`` CTAs (one per SM); these GEMMs are bound by the
    operand bytes a CTA streams in from L2, so the splits are handed out greedily to the problem whose
    CTAs currently stream the most bytes.  Returns [(tile_n, splits)]."""
    target = int(target_ctas or _os.environ.get("PQLB_WGRAD_CTAS", 148))
    kb = (K + 31) // 32
    max_split = max(1, kb // 2)
    info = []
    for M, N in problems:
        tn = wgrad_tile_n(N)
        tiles, cost = 0, 0.0
        for m0 in range(0, M, 128):
            for n0 in range(0, N, tn):
                tiles += 1
                cost = max(cost, 4096.0 * (-(-min(M - m0, 128) // 32) + -(-min(N - n0, tn) // 32)))
        info.append([tn, tiles, cost, 1])
    total = sum(t for _, t, _, _ in info)
    while True:
        order = sorted(range(len(info)), key=lambda i: -info[i][2] / info[i][3])
        for i in order:
            if info[i][3] < max_split and total + info[i][1] <= target:
                info[i][3] += 1
                total += info[i][1]
                break
        else:
            break
    return [(tn, s) for tn, _, _, s in info]


class WgradMulti(Call):
    """All weight gradients of an update in one launch (pqlb_wgrad_multi).  ``problems``: dicts with
    dz, lddz, h, ldh, part, ldo, split_stride, M, N, tile_n, splits (addresses from ``addr``)."""

    def __init__(self, K, problems):
        d = _lib.WgradDesc()
        d.K, d.n_problems = int(K), len(problems)
        if not 1 <= len(problems) <= _lib.MAX_WGRAD:
            raise ValueError("1..%d problems per launch" % _lib.MAX_WGRAD)
        for i, g in enumerate(problems):
            for k in ("dz", "h", "part"):
                setattr(d.p[i], k, g[k])
            for k in ("lddz", "ldh", "ldo", "split_stride", "M", "N", "tile_n", "splits"):
                setattr(d.p[i], k, int(g[k]))
        self.desc = d
        super().__init__("pqlb_wgrad_multi", C.byref(d))


WGRAD_CLUSTER = int(_os.environ.get("PQLB_WGRAD_CLUSTER", 1))


def wgrad_cluster(splits, tile_n):
    """Thread-block cluster size for the in-cluster (distributed shared memory) reduction of the
    split-K partials: the largest of 8 / 4 / 2 that divides the split count (tile_n >= 32: the
    receive area is counted in 32-column chunks)."""
    if tile_n < 32:
        return 1
    for c in (8, 4, 2):
        if c <= WGRAD_CLUSTER and splits % c == 0:
            return c
    return 1


class Workspace:
    """Bump allocator over one fp32 device buffer (so a learner's scratch is one allocation
    and every sub-buffer address is fixed for the lifetime of the prepared plan)."""

    def __init__(self, device):
        self.device = device
        self._specs = []
        self._total = 0
        self.buf = None

    def reserve(self, name, n_words, align=32):
        off = (self._total + align - 1) // align * align
        self._specs.append((name, off, int(n_words)))
        self._total = off + int(n_words)
        return off

    def commit(self):
        self.buf = torch.zeros(max(self._total, 1), dtype=torch.float32, device=self.device)
        self.views = {name: self.buf[off:off + n] for name, off, n in self._specs}
        return self

    def __getitem__(self, name):
        return self.views[name]
