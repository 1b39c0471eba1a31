"""torch-CPU fp32 restatement of the reference learner arithmetic (test infrastructure only).

Follows
  /root/reference/pql/utils/common.py:139-145        normalize
  /root/reference/pql/models/mlp.py:15-40            create_simple_mlp / MLPNet (Linear-ELU x3, Linear)
  /root/reference/pql/models/mlp.py:177-179          TanhMLPPolicy.forward
  /root/reference/pql/models/mlp.py:186-203          DoubleQ
  /root/reference/pql/models/mlp.py:244-267          DistributionalDoubleQ
  /root/reference/pql/utils/noise.py:19-27           add_normal_noise
  /root/reference/pql/utils/distl_util.py:4-20       projection
  /root/reference/pql/utils/torch_util.py:9-12       soft_update
  /root/reference/pql/algo/pql_v_learner.py:73-133   PQLVLearner.learn / optimizer_update
  /root/reference/pql/algo/pql_p_learner.py:47-96    PQLPLearner.learn / optimizer_update
and the arithmetic the reference inherits from torch (clip_grad_norm_, AdamW).

Random draws (replay indices, target-policy noise) are *inputs* here so that the
oracle, the reference and the CUDA path can be fed identical values.
Parameters are plain lists ``[(W0,b0),(W1,b1),(W2,b2),(W3,b3)]`` of fp32 tensors
with ``W`` shaped ``[out, in]`` exactly like ``nn.Linear.weight``.
"""
import contextlib
import math

import numpy as np
import torch
import torch.nn.functional as F

HIDDEN = (512, 256, 128)          # mlp.py:33-34
LAYER_KEYS = (0, 2, 4, 6)         # nn.Sequential positions of the Linear layers


# ----------------------------------------------------------------------------- parameters
def init_mlp(in_dim, out_dim, generator=None):
    """nn.Linear default init (kaiming_uniform(a=sqrt(5)) == U(+-1/sqrt(fan_in)) for W and b)."""
    dims = (in_dim, *HIDDEN, out_dim)
    params = []
    for i, o in zip(dims[:-1], dims[1:]):
        bound = 1.0 / math.sqrt(i)
        w = (torch.rand(o, i, generator=generator) * 2 - 1) * bound
        b = (torch.rand(o, generator=generator) * 2 - 1) * bound
        params.append((w.float(), b.float()))
    return params


def params_from_state_dict(sd, prefix):
    """``prefix`` e.g. 'net_q1.net.' or 'net.' (state_dict names, evaluator.py:115-116)."""
    return [(sd[f"{prefix}{k}.weight"].detach().clone().float(),
             sd[f"{prefix}{k}.bias"].detach().clone().float()) for k in LAYER_KEYS]


def clone_params(params, requires_grad=False):
    return [(w.detach().clone().requires_grad_(requires_grad),
             b.detach().clone().requires_grad_(requires_grad)) for w, b in params]


def flat(params_list):
    """Flatten in nn.Module.parameters() order: per net, per layer, weight then bias."""
    return [t for params in params_list for wb in params for t in wb]


# ----------------------------------------------------------------------------- TF32 operand model
# The CUDA path feeds the tensor cores TF32 operands (fp32 accumulate).  With ``tf32_operands()``
# active, ``mlp`` applies exactly the roundings the kernels apply - and nothing else - so that the
# kernels can be checked against "the reference arithmetic with TF32-rounded GEMM operands" to
# ~1e-5, separately from the cost of the number format itself (fp32 oracle vs TF32 oracle):
#   * every GEMM operand is rounded to nearest (ties away) to a 10-bit mantissa when it is
#     produced: inputs, weights, h = elu(z) and the back-propagated dz = g * elu'(h);
#   * products are exact and accumulated in fp32 or better; biases, ELU, tanh, softmax, losses,
#     the scalar Q head (a dot product with the un-rounded h3 and fp32 w4) stay fp32;
#   * elu'(h) is evaluated on the rounded h (that is what the forward pass stored).
#   * ``precise_critic`` (default): the critics' TRUNK FORWARD is computed by the split-fp16 kernel
#     (csrc/mlp_fwd_h.cu: every operand as hi + lo halves, three MMAs per product = 22 significand
#     bits), modelled here as un-rounded operands; what the forward STORES for the backward pass
#     (h1..h3) is still TF32-rounded, and dgrad / wgrad / the C51 head keep TF32 operands.
_TF32 = False
_PRECISE_CRITIC = True


@contextlib.contextmanager
def tf32_operands(on=True, precise_critic=True):
    global _TF32, _PRECISE_CRITIC
    prev, _TF32, _PRECISE_CRITIC = (_TF32, _PRECISE_CRITIC), on, precise_critic
    try:
        yield
    finally:
        _TF32, _PRECISE_CRITIC = prev


def rn_tf32(x):
    """cvt.rna.tf32.f32 on every element: round to nearest, ties away, 10-bit mantissa."""
    i = x.detach().contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def _mm(a, b):
    return (a.double() @ b.double()).float()


class _LinELU(torch.autograd.Function):
    """hr, h = rn(elu(z)), elu(z) (one hidden layer of mlp.py:15-24) with z = rn(x) rn(W)^T + b, or
    z = x W^T + b on un-rounded operands when ``precise`` (split-fp16 forward).  Either way the backward
    pass sees the TF32-rounded x, W and h."""

    @staticmethod
    def forward(ctx, x, w, b, precise):
        xr, wr = rn_tf32(x), rn_tf32(w)
        z = _mm(x.detach(), w.detach().t()) if precise else _mm(xr, wr.t())
        h = F.elu(z + b.detach())
        hr = rn_tf32(h)
        ctx.save_for_backward(xr, wr, hr)
        return hr, h

    @staticmethod
    def backward(ctx, g_hr, g_h):
        xr, wr, hr = ctx.saved_tensors
        g = g_hr + g_h
        dz = rn_tf32(g * torch.where(hr > 0, torch.ones_like(hr), hr + 1))
        return _mm(dz, wr), _mm(dz.t(), xr), dz.sum(0), None


class _LinOut(torch.autograd.Function):
    """Last layer with more than one output (actor / C51 logits): z = rn(x) rn(W)^T + b in fp32 (or
    on un-rounded operands for the C51 head of a precise critic); the incoming gradient is rounded
    (the loss / tanh' kernels emit TF32 operands), the backward contractions use the rounded x and W."""

    @staticmethod
    def forward(ctx, x, w, b, x_exact):
        xr, wr = rn_tf32(x), rn_tf32(w)
        ctx.save_for_backward(xr, wr)
        if x_exact is not None:          # split-fp16 head (C51 logits fused behind the precise trunk)
            return _mm(x_exact.detach(), w.detach().t()) + b.detach()
        return _mm(xr, wr.t()) + b.detach()

    @staticmethod
    def backward(ctx, g):
        xr, wr = ctx.saved_tensors
        dz = rn_tf32(g)
        return _mm(dz, wr), _mm(dz.t(), xr), dz.sum(0), None


class _ScalarHead(torch.autograd.Function):
    """q = h3 w4^T + b4 on CUDA cores: un-rounded h3 and fp32 w4 forward; the weight gradient
    uses the stored (rounded) h3."""

    @staticmethod
    def forward(ctx, h, hr, w, b):
        ctx.save_for_backward(hr, w.detach())
        return _mm(h.detach(), w.detach().t()) + b.detach()

    @staticmethod
    def backward(ctx, g):
        hr, w = ctx.saved_tensors
        return g * w, None, _mm(g.t(), hr), g.sum(0)


def _mlp_tf32(x, params, precise=False):
    hr, h = x, x
    for w, b in params[:-1]:
        # precise: the next layer consumes the un-rounded activation (hi + lo halves in tensor memory)
        hr, h = _LinELU.apply(h if precise else hr, w, b, precise)
    w, b = params[-1]
    if w.shape[0] == 1:
        return _ScalarHead.apply(h, hr, w, b)
    return _LinOut.apply(hr, w, b, h if precise else None)


# ----------------------------------------------------------------------------- forward pieces
def normalize(x, norm):
    """common.py:139-145: clamp((x-mean)/sqrt(var+eps), -5, 5); identity when norm is None."""
    if norm is None:
        return x
    mean, var, eps = norm
    return torch.clamp((x - mean.float()) / torch.sqrt(var.float() + eps), min=-5.0, max=5.0)


def mlp(x, params, critic=False):
    """mlp.py:15-24: ELU after every layer but the last."""
    if _TF32:
        return _mlp_tf32(x, params, precise=critic and _PRECISE_CRITIC)
    h = x
    last = len(params) - 1
    for i, (w, b) in enumerate(params):
        h = F.linear(h, w, b)
        if i < last:
            h = F.elu(h)
    return h


def actor_forward(obs, actor):
    """mlp.py:177-179."""
    return torch.tanh(mlp(obs, actor))


def target_policy_action(next_obs, actor, noise, noise_bound=0.2):
    """pql_v_learner.py:62-71 + noise.py:19-27; ``noise`` = the N(0, std^2) draw before clamping."""
    a = actor_forward(next_obs, actor)
    return torch.clamp(a + torch.clamp(noise, -noise_bound, noise_bound), -1.0, 1.0)


def q1_q2(obs, act, q1, q2, distl=False):
    """mlp.py:197-199 / 261-263."""
    x = torch.cat((obs, act), dim=1)
    o1, o2 = mlp(x, q1, critic=True), mlp(x, q2, critic=True)
    if distl:
        return torch.softmax(o1, dim=1), torch.softmax(o2, dim=1)
    return o1, o2


def q_min(obs, act, q1, q2, distl=False, z_atoms=None):
    """mlp.py:194-195 / 255-259 (C51: expectation under z_atoms, shape [B])."""
    a, b = q1_q2(obs, act, q1, q2, distl)
    if distl:
        a = torch.sum(a * z_atoms, dim=1)
        b = torch.sum(b * z_atoms, dim=1)
    return torch.min(a, b)


def projection(next_dist, reward, done, gamma, v_min=-10.0, v_max=10.0, num_atoms=51):
    """distl_util.py:4-20, restated with the CPU accumulation order of index_add_:
    all lower-atom terms in ascending source order, then all upper-atom terms."""
    B = reward.shape[0]
    delta_z = (v_max - v_min) / (num_atoms - 1)
    support = torch.linspace(v_min, v_max, num_atoms)
    tz = (reward + (1 - done) * gamma * support).clamp(min=v_min, max=v_max)
    b = (tz - v_min) / delta_z
    l = b.floor().long()
    u = b.ceil().long()
    l = torch.where((u > 0) & (l == u), l - 1, l)
    u = torch.where((l < (num_atoms - 1)) & (l == u), u + 1, u)
    wl = (next_dist * (u.float() - b)).numpy()
    wu = (next_dist * (b - l.float())).numpy()
    out = np.zeros((B, num_atoms), np.float32)
    rows = np.arange(B)[:, None].repeat(num_atoms, 1)
    np.add.at(out, (rows, l.numpy()), wl)
    np.add.at(out, (rows, u.numpy()), wu)
    return torch.from_numpy(out)


# ----------------------------------------------------------------------------- optimiser pieces
def clip_grad_norm(grads, max_norm):
    """torch/nn/utils/clip_grad.py: coef = min(1, max_norm/(||g||+1e-6)); g *= coef always."""
    norms = torch.stack([torch.linalg.vector_norm(g) for g in grads])
    total = torch.linalg.vector_norm(norms)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return [g * coef for g in grads], total


class AdamW:
    """torch.optim.AdamW(params, lr) with all other defaults (pql_v_learner.py:46):
    betas (0.9, 0.999), eps 1e-8, decoupled weight_decay 0.01; SURVEY App. E order."""

    def __init__(self, tensors, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01):
        self.lr, self.b1, self.b2, self.eps, self.wd = lr, betas[0], betas[1], eps, weight_decay
        self.m = [torch.zeros_like(t) for t in tensors]
        self.v = [torch.zeros_like(t) for t in tensors]
        self.t = 0

    @torch.no_grad()
    def step(self, tensors, grads):
        self.t += 1
        bc1 = 1 - self.b1 ** self.t
        bc2 = 1 - self.b2 ** self.t
        step_size = self.lr / bc1
        bc2_sqrt = math.sqrt(bc2)
        for p, g, m, v in zip(tensors, grads, self.m, self.v):
            p.mul_(1 - self.lr * self.wd)
            m.lerp_(g, 1 - self.b1)
            v.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            denom = (v.sqrt() / bc2_sqrt).add_(self.eps)
            p.addcdiv_(m, denom, value=-step_size)


@torch.no_grad()
def polyak(target, current, tau):
    """torch_util.py:9-12: theta' <- theta*tau + theta'*(1-tau), parameters only."""
    for t, c in zip(target, current):
        t.copy_(c * tau + t * (1.0 - tau))


# ----------------------------------------------------------------------------- learners
class VLearnerOracle:
    """One critic update == PQLVLearner.learn() (pql_v_learner.py:73-115) on an injected batch."""

    def __init__(self, q1, q2, lr=5e-4, tau=0.05, gamma=0.99, nstep=3, max_grad_norm=0.5,
                 distl=False, v_min=-10.0, v_max=10.0, num_atoms=51, noise_bound=0.2):
        self.q1 = clone_params(q1, True)
        self.q2 = clone_params(q2, True)
        self.tq1 = clone_params(q1)           # deepcopy(critic), :47
        self.tq2 = clone_params(q2)
        self.opt = AdamW(flat([self.q1, self.q2]), lr)
        self.tau, self.max_grad_norm, self.distl = tau, max_grad_norm, distl
        self.gamma_n = gamma ** nstep
        self.v_min, self.v_max, self.num_atoms = v_min, v_max, num_atoms
        self.noise_bound = noise_bound
        self.last = {}

    def learn(self, batch, noise, actor, norm):
        obs, action, reward, next_obs, done = batch
        obs = normalize(obs, norm)
        next_obs = normalize(next_obs, norm)
        with torch.no_grad():
            next_act = target_policy_action(next_obs, actor, noise, self.noise_bound)
            if self.distl:
                t1, t2 = q1_q2(next_obs, next_act, self.tq1, self.tq2, True)
                p1 = projection(t1, reward, done, self.gamma_n, self.v_min, self.v_max, self.num_atoms)
                p2 = projection(t2, reward, done, self.gamma_n, self.v_min, self.v_max, self.num_atoms)
                target = torch.min(p1, p2)
            else:
                target = q_min(next_obs, next_act, self.tq1, self.tq2)
                target = reward + (1 - done) * self.gamma_n * target
        c1, c2 = q1_q2(obs, action, self.q1, self.q2, self.distl)
        if self.distl:
            loss = F.binary_cross_entropy(c1, target) + F.binary_cross_entropy(c2, target)
        else:
            loss = F.mse_loss(c1, target) + F.mse_loss(c2, target)
        tensors = flat([self.q1, self.q2])
        grads = list(torch.autograd.grad(loss, tensors))
        raw = [g.clone() for g in grads]
        gnorm = None
        if self.max_grad_norm is not None:
            grads, gnorm = clip_grad_norm(grads, self.max_grad_norm)
        self.opt.step(tensors, grads)
        polyak(flat([self.tq1, self.tq2]), tensors, self.tau)
        self.last = dict(loss=loss.detach(), q1=c1.detach(), q2=c2.detach(), target=target,
                         grads=raw, grad_norm=gnorm, next_action=next_act)
        return float(loss.detach())


class PLearnerOracle:
    """One actor update == PQLPLearner.learn() (pql_p_learner.py:47-64) on an injected obs batch."""

    def __init__(self, actor, lr=5e-4, max_grad_norm=0.5, distl=False, v_min=-10.0, v_max=10.0,
                 num_atoms=51):
        self.actor = clone_params(actor, True)
        self.opt = AdamW(flat([self.actor]), lr)
        self.max_grad_norm, self.distl = max_grad_norm, distl
        self.z = torch.linspace(v_min, v_max, num_atoms) if distl else None
        self.last = {}

    def learn(self, obs, q1, q2, norm):
        obs = normalize(obs, norm)
        act = actor_forward(obs, self.actor)
        qm = q_min(obs, act, q1, q2, self.distl, self.z)
        loss = -qm.mean()
        tensors = flat([self.actor])
        grads = list(torch.autograd.grad(loss, tensors))
        raw = [g.clone() for g in grads]
        gnorm = None
        if self.max_grad_norm is not None:
            grads, gnorm = clip_grad_norm(grads, self.max_grad_norm)
        self.opt.step(tensors, grads)
        self.last = dict(loss=loss.detach(), action=act.detach(), q=qm.detach(), grads=raw,
                         grad_norm=gnorm)
        return float(loss.detach())
