"""CPU study (no CUDA): which GEMM sites of a V-/P-learner update need more than TF32 operands for
every gradient tensor to stay within 1e-3 of the fp32 reference arithmetic (oracle/learner.py).

Every contraction of the update is a *site* ``(role, op)``: role in {actor, q_tgt, q_cur}, op in
{fwd, dgrad, wgrad}.  A site runs in one of the modes
    tf32   both operands rounded to TF32 (what one tcgen05 kind::tf32 MMA computes)
    a3     A exact (hi + lo), W rounded        = hi.hi + lo.hi
    w3     A rounded, W exact                  = hi.hi + hi.lo
    x3     both split: hi.hi + lo.hi + hi.lo   (3xTF32; the lo.lo term is dropped: 2^-22)
    fp32   exact operands
Usage: python tools/precision_study.py [B]
"""
import copy
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from oracle import learner as L          # noqa: E402
from tests.golden import inputs         # noqa: E402

POLICY = {}


def rn(x):
    return L.rn_tf32(x)


def split(x, mode_exact):
    """(hi, lo): lo is the TF32-rounded residual (None when this operand is only rounded)."""
    hi = rn(x)
    return (hi, rn(x.detach() - hi)) if mode_exact else (hi, None)


def mm(a, b, mode):
    """a @ b under the site's operand model; a: [M,K], b: [K,N]."""
    if mode == "fp32":
        return (a.double() @ b.double()).float()
    ah, al = split(a, mode in ("a3", "x3"))
    bh, bl = split(b, mode in ("w3", "x3"))
    out = ah.double() @ bh.double()
    if al is not None:
        out = out + al.double() @ bh.double()
    if bl is not None:
        out = out + ah.double() @ bl.double()
    return out.float()


class Lin(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, role, act):
        z = mm(x.detach(), w.detach().t(), POLICY.get((role, "fwd"), "tf32")) + b.detach()
        h = F.elu(z) if act else z
        ctx.save_for_backward(x.detach(), w.detach(), h)
        ctx.role, ctx.act = role, act
        return h

    @staticmethod
    def backward(ctx, g):
        x, w, h = ctx.saved_tensors
        dz = g * torch.where(h > 0, torch.ones_like(h), h + 1) if ctx.act else g
        dx = mm(dz, w, POLICY.get((ctx.role, "dgrad"), "tf32"))
        dw = mm(dz.t(), x, POLICY.get((ctx.role, "wgrad"), "tf32"))
        return dx, dw, dz.sum(0), None, None


def mlp(x, params, role):
    h = x
    for i, (w, b) in enumerate(params):
        last = i == len(params) - 1
        if last and w.shape[0] == 1:
            h = F.linear(h, w, b)                 # scalar head: fp32 CUDA-core dot product
        else:
            h = Lin.apply(h, w, b, role, not last)
    return h


def v_grads(case, noise, B):
    obs, action, reward, next_obs, done = case["batch"]
    norm = case["norm"]
    obs, next_obs = L.normalize(obs, norm), L.normalize(next_obs, norm)
    q1, q2 = L.clone_params(case["q1"], True), L.clone_params(case["q2"], True)
    with torch.no_grad():
        a = torch.tanh(mlp(next_obs, case["actor"], "actor"))
        a = torch.clamp(a + torch.clamp(noise, -0.2, 0.2), -1, 1)
        xt = torch.cat((next_obs, a), 1)
        y = reward + (1 - done) * 0.99 ** 3 * torch.min(mlp(xt, case["q1"], "q_tgt"), mlp(xt, case["q2"], "q_tgt"))
    x = torch.cat((obs, action), 1)
    loss = F.mse_loss(mlp(x, q1, "q_cur"), y) + F.mse_loss(mlp(x, q2, "q_cur"), y)
    return list(torch.autograd.grad(loss, L.flat([q1, q2])))


def p_grads(case, B):
    obs = L.normalize(case["batch"][0], case["norm"])
    actor = L.clone_params(case["actor"], True)
    act = torch.tanh(mlp(obs, actor, "actor"))
    x = torch.cat((obs, act), 1)
    loss = -torch.min(mlp(x, case["q1"], "q_cur"), mlp(x, case["q2"], "q_cur")).mean()
    return list(torch.autograd.grad(loss, L.flat([actor])))


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def report(name, policy, ref_v, ref_p, case, B):
    global POLICY
    POLICY = policy
    gv = v_grads(case, case["noises"][0], B)
    gp = p_grads(case, B)
    ev = [rel(a, b) for a, b in zip(gv, ref_v)]
    ep = [rel(a, b) for a, b in zip(gp, ref_p)]
    # the 1-element head bias is judged against sum |dq| in tests/parity.py; skip it here
    ev_ = [e for e, g in zip(ev, ref_v) if g.numel() > 1]
    print(f"{name:58s} V max {max(ev_):.2e} (bias {max(ev_[1::2]):.2e} w {max(ev_[0::2]):.2e})   "
          f"P max {max(ep):.2e} (bias {max(ep[1::2]):.2e} w {max(ep[0::2]):.2e})", flush=True)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    torch.set_num_threads(8)
    case = inputs.learner_case(1234, B, 88, 16)
    sites = [(r, o) for r in ("actor", "q_tgt", "q_cur") for o in ("fwd", "dgrad", "wgrad")]
    fp32 = {s: "fp32" for s in sites}
    global POLICY
    POLICY = fp32
    ref_v, ref_p = v_grads(case, case["noises"][0], B), p_grads(case, B)
    report("all tf32", {}, ref_v, ref_p, case, B)
    for mode in ("x3", "w3", "a3"):
        report(f"all fwd {mode}", {(r, "fwd"): mode for r in ("actor", "q_tgt", "q_cur")}, ref_v, ref_p, case, B)
    report("q_cur fwd x3", {("q_cur", "fwd"): "x3"}, ref_v, ref_p, case, B)
    report("q_cur+q_tgt fwd x3", {("q_cur", "fwd"): "x3", ("q_tgt", "fwd"): "x3"}, ref_v, ref_p, case, B)
    report("q_cur+q_tgt fwd w3", {("q_cur", "fwd"): "w3", ("q_tgt", "fwd"): "w3"}, ref_v, ref_p, case, B)
    report("q_cur+actor fwd x3", {("q_cur", "fwd"): "x3", ("actor", "fwd"): "x3"}, ref_v, ref_p, case, B)
    report("q_cur+actor fwd x3, q_cur dgrad x3", {("q_cur", "fwd"): "x3", ("actor", "fwd"): "x3", ("q_cur", "dgrad"): "x3"}, ref_v, ref_p, case, B)
    report("q_cur+actor fwd w3, q_cur dgrad w3", {("q_cur", "fwd"): "w3", ("actor", "fwd"): "w3", ("q_cur", "dgrad"): "w3"}, ref_v, ref_p, case, B)
    report("all fwd+dgrad x3", {(r, o): "x3" for r in ("actor", "q_tgt", "q_cur") for o in ("fwd", "dgrad")}, ref_v, ref_p, case, B)
    report("all fwd+dgrad w3", {(r, o): "w3" for r in ("actor", "q_tgt", "q_cur") for o in ("fwd", "dgrad")}, ref_v, ref_p, case, B)
    report("all x3", {s: "x3" for s in sites}, ref_v, ref_p, case, B)


if __name__ == "__main__":
    main()
