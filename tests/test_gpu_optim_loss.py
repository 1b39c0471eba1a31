"""GPU tier: K4 (deterministic gradient reduction, global-norm clip, AdamW, Polyak) and the
loss-side kernels (twin-Q TD/MSE, DPG, C51 projection + BCE) against the torch-CPU oracle."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import learner as L

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rn_tf32(x):
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def seg_table(n, seg=1024, ws_off=0, ws_stride=0, n_part=0):
    rows = []
    for off in range(0, n, seg):
        rows.append([off, min(seg, n - off), ws_off + off, ws_stride, n_part])
    return torch.tensor(rows, dtype=torch.int64, device=DEV)


@pytest.mark.parametrize("n,max_norm,with_target", [(436226, 0.5, True), (211856, 0.5, False), (1000, -1.0, True),
                                                    (4099, 1e9, True)])
def test_adamw_polyak_matches_torch_order(n, max_norm, with_target):
    from pql_b200 import _lib
    g = torch.Generator().manual_seed(n)
    p0 = torch.randn(n, generator=g) * 0.1
    tgt0 = torch.randn(n, generator=g) * 0.1
    p = p0.clone().to(DEV); tgt = tgt0.clone().to(DEV)
    m = torch.zeros(n, device=DEV); v = torch.zeros(n, device=DEV)
    p_tf = torch.zeros(n, device=DEV); t_tf = torch.zeros(n, device=DEV)
    opt = L.AdamW([p0], 5e-4)
    ref_p, ref_t = p0, tgt0.clone()
    segs = seg_table(n)
    sumsq = torch.zeros(segs.shape[0], device=DEV)
    norm_out = torch.zeros(1, device=DEV)
    count = torch.zeros(1, dtype=torch.int64, device=DEV)     # device-resident step count (n == 4099 case)
    for step in range(1, 4):
        grad = torch.randn(n, generator=g) * (0.01 * step)
        gd = grad.to(DEV)
        _lib.call("pqlb_grad_sumsq", _lib.ptr(segs), segs.shape[0], _lib.ptr(gd), _lib.ptr(sumsq))
        _lib.call("pqlb_adamw_polyak", _lib.ptr(p), _lib.ptr(gd), _lib.ptr(m), _lib.ptr(v),
                  _lib.ptr(tgt) if with_target else None, _lib.ptr(p_tf), _lib.ptr(t_tf) if with_target else None,
                  None, None, n, _lib.ptr(sumsq), segs.shape[0], 1.0, max_norm, 5e-4, 0.9, 0.999, 1e-8, 0.01,
                  step if n != 4099 else 0, _lib.ptr(count) if n == 4099 else None, 0.05, _lib.ptr(norm_out))
        count += 1
        grads = [grad]
        if max_norm >= 0:
            grads, total = L.clip_grad_norm(grads, max_norm)
            assert norm_out.item() == pytest.approx(total.item(), rel=1e-5)
        opt.step([ref_p], grads)
        if with_target:
            L.polyak([ref_t], [ref_p], 0.05)
        np.testing.assert_allclose(p.cpu().numpy(), ref_p.numpy(), rtol=2e-6, atol=1e-8)
        if with_target:
            np.testing.assert_allclose(tgt.cpu().numpy(), ref_t.numpy(), rtol=2e-6, atol=1e-8)
            assert torch.equal(t_tf, rn_tf32(tgt))
        assert torch.equal(p_tf, rn_tf32(p))


def test_grad_reduce_is_deterministic_and_exact():
    from pql_b200 import _lib
    n, parts = 5000, 7
    g = torch.Generator().manual_seed(3)
    ws = torch.randn(parts, n, generator=g).to(DEV)
    segs = seg_table(n, ws_off=0, ws_stride=n, n_part=parts)
    grad = torch.zeros(n, device=DEV); sumsq = torch.zeros(segs.shape[0], device=DEV)
    outs = []
    for _ in range(2):
        _lib.call("pqlb_grad_reduce", _lib.ptr(segs), segs.shape[0], _lib.ptr(ws), _lib.ptr(grad), _lib.ptr(sumsq))
        outs.append((grad.clone(), sumsq.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    ref = torch.zeros(n)
    for k in range(parts):
        ref = ref + ws[k].cpu()          # same left-to-right association
    assert torch.equal(grad.cpu(), ref)
    assert sumsq.sum().item() == pytest.approx((ref.double() ** 2).sum().item(), rel=1e-5)


def test_colsum_and_sum_partials():
    from pql_b200 import _lib
    rows, cols, ld = 1000, 51, 64
    dz = torch.randn(rows, ld, device=DEV)
    nblk = (rows + 127) // 128
    part = torch.zeros(nblk, cols, device=DEV)
    _lib.call("pqlb_colsum_partial", _lib.ptr(dz), ld, rows, cols, _lib.ptr(part))
    np.testing.assert_allclose(part.sum(0).cpu().numpy(), dz[:, :cols].double().sum(0).cpu().numpy(), rtol=1e-5, atol=1e-5)
    out = torch.zeros(1, device=DEV)
    flat = part.reshape(-1)
    _lib.call("pqlb_sum_partials", _lib.ptr(flat), flat.numel(), 0.5, _lib.ptr(out), None, None, 0)
    assert out.item() == pytest.approx(0.5 * flat.double().sum().item(), rel=1e-5, abs=1e-5)
    counter = torch.full((1,), 7, dtype=torch.int64, device=DEV)
    ring = torch.zeros(5, device=DEV)
    _lib.call("pqlb_sum_partials", _lib.ptr(flat), flat.numel(), 2.0, _lib.ptr(out), _lib.ptr(counter), _lib.ptr(ring), 5)
    assert counter.item() == 8 and ring[2].item() == out.item() and torch.count_nonzero(ring).item() == 1


def test_doubleq_td_loss_and_head_backward():
    from pql_b200 import _lib
    B = 1000
    g = torch.Generator().manual_seed(11)
    q1, q2, tq1, tq2 = (torch.randn(B, generator=g) for _ in range(4))
    reward = torch.randn(B, generator=g) * 0.1
    done = (torch.rand(B, generator=g) < 0.2).float()
    h3 = [rn_tf32(F.elu(torch.randn(B, 128, generator=g))) for _ in range(2)]
    w4 = [torch.randn(128, generator=g) * 0.1 for _ in range(2)]
    gamma_n = float(np.float32(0.99 ** 3))
    d = lambda x: x.to(DEV).contiguous()
    nblk = (B + 63) // 64
    dz = [torch.zeros(B, 128, device=DEV) for _ in range(2)]
    ws = [torch.zeros(nblk, 129, device=DEV) for _ in range(2)]
    y = torch.zeros(B, device=DEV); lp = torch.zeros(2 * nblk, device=DEV)
    dev_in = [d(x) for x in (q1, q2, tq1, tq2, reward, done)]
    dh3 = [d(x) for x in h3]; dw4 = [d(x) for x in w4]
    _lib.call("pqlb_doubleq_td_loss", *(_lib.ptr(x) for x in dev_in), gamma_n, B, _lib.ptr(dh3[0]), _lib.ptr(dh3[1]),
              _lib.ptr(dw4[0]), _lib.ptr(dw4[1]), _lib.ptr(dz[0]), _lib.ptr(dz[1]), _lib.ptr(y), _lib.ptr(ws[0]),
              _lib.ptr(ws[1]), _lib.ptr(lp))
    yref = reward + (1 - done) * (0.99 ** 3) * torch.min(tq1, tq2)
    np.testing.assert_allclose(y.cpu().numpy(), yref.numpy(), rtol=1e-6, atol=1e-7)
    loss = F.mse_loss(q1, yref) + F.mse_loss(q2, yref)
    assert lp.sum().item() / B == pytest.approx(loss.item(), rel=1e-5)
    for i, q in enumerate((q1, q2)):
        dq = 2 * (q - yref) / B
        dz_ref = dq[:, None] * w4[i][None, :] * torch.where(h3[i] > 0, torch.ones_like(h3[i]), h3[i] + 1)
        np.testing.assert_allclose(dz[i].cpu().numpy(), rn_tf32(dz_ref).numpy(), rtol=2e-3, atol=1e-9)
        gw = ws[i].sum(0).cpu()
        np.testing.assert_allclose(gw[:128].numpy(), (dq[:, None] * h3[i]).sum(0).numpy(), rtol=1e-4, atol=1e-6)
        assert gw[128].item() == pytest.approx(dq.sum().item(), rel=1e-4, abs=1e-7)


def test_dpg_loss_and_head_backward():
    from pql_b200 import _lib
    B = 700
    g = torch.Generator().manual_seed(12)
    q1, q2 = torch.randn(B, generator=g), torch.randn(B, generator=g)
    q2[:10] = q1[:10]                       # ties: torch.min backward splits the gradient evenly
    h3 = [rn_tf32(F.elu(torch.randn(B, 128, generator=g))) for _ in range(2)]
    w4 = [torch.randn(128, generator=g) * 0.1 for _ in range(2)]
    d = lambda x: x.to(DEV).contiguous()
    nblk = (B + 63) // 64
    dz = [torch.zeros(B, 128, device=DEV) for _ in range(2)]
    lp = torch.zeros(2 * nblk, device=DEV)
    dh3 = [d(x) for x in h3]; dw4 = [d(x) for x in w4]
    dq1, dq2 = d(q1), d(q2)
    _lib.call("pqlb_dpg_loss", _lib.ptr(dq1), _lib.ptr(dq2), B, _lib.ptr(dh3[0]), _lib.ptr(dh3[1]), _lib.ptr(dw4[0]),
              _lib.ptr(dw4[1]), _lib.ptr(dz[0]), _lib.ptr(dz[1]), _lib.ptr(lp))
    a, b = q1.clone().requires_grad_(), q2.clone().requires_grad_()
    loss = -torch.min(a, b).mean()
    loss.backward()
    assert -lp.sum().item() / B == pytest.approx(loss.item(), rel=1e-5)
    for i, gq in enumerate((a.grad, b.grad)):
        dz_ref = gq[:, None] * w4[i][None, :] * torch.where(h3[i] > 0, torch.ones_like(h3[i]), h3[i] + 1)
        np.testing.assert_allclose(dz[i].cpu().numpy(), rn_tf32(dz_ref).numpy(), rtol=2e-3, atol=1e-10)


def test_c51_projection_bce_and_softmax_backward(golden_dir):
    """Projection against the reference-generated known-answer fixture (bit-exact) and the
    BCE loss / logit gradients against torch autograd."""
    import os
    from pql_b200 import _lib
    kat = np.load(os.path.join(golden_dir, "projection_kat.npz"))
    B, N, ld = kat["dist"].shape[0], 51, 64
    g = torch.Generator().manual_seed(13)
    tp1 = torch.from_numpy(kat["dist"])
    tp2 = torch.softmax(torch.randn(B, N, generator=g) * 2, dim=1)
    reward, done = torch.from_numpy(kat["reward"]), torch.from_numpy(kat["done"])
    logits = [torch.randn(B, N, generator=g).requires_grad_() for _ in range(2)]
    p = [torch.softmax(x, dim=1) for x in logits]
    gamma_n = 0.99 ** 3
    pr1 = L.projection(tp1, reward, done, gamma_n)
    pr2 = L.projection(tp2, reward, done, gamma_n)
    assert np.array_equal(pr1.numpy().view(np.uint32), kat["out"].view(np.uint32))
    target = torch.min(pr1, pr2)
    loss = F.binary_cross_entropy(p[0], target) + F.binary_cross_entropy(p[1], target)
    loss.backward()

    def pad(x):
        o = torch.zeros(B, ld); o[:, :N] = x.detach(); return o.to(DEV)
    dp = [pad(x) for x in p]; dtp = [pad(tp1), pad(tp2)]
    z = torch.linspace(-10, 10, N).to(DEV)
    tout = torch.zeros(B, N, device=DEV)
    dl = [torch.full((B, ld), 9.0, device=DEV) for _ in range(2)]
    nblk = (B + 7) // 8
    lp = torch.zeros(nblk, device=DEV)
    r, dn = reward.reshape(-1).to(DEV), done.reshape(-1).to(DEV)
    _lib.call("pqlb_c51_td_loss", _lib.ptr(dp[0]), _lib.ptr(dp[1]), _lib.ptr(dtp[0]), _lib.ptr(dtp[1]), ld,
              _lib.ptr(r), _lib.ptr(dn), _lib.ptr(z), float(np.float32(gamma_n)), -10.0, 10.0, N, B, _lib.ptr(tout),
              _lib.ptr(dl[0]), _lib.ptr(dl[1]), ld, _lib.ptr(lp))
    # the kernel's own projection of tp1 must reproduce the reference fixture wherever it is the min
    np.testing.assert_array_equal(tout.cpu().numpy().view(np.uint32), target.numpy().view(np.uint32))
    assert lp.sum().item() / (B * N) == pytest.approx(loss.item(), rel=1e-5)
    for i in range(2):
        ref = logits[i].grad
        got = dl[i].cpu()
        assert torch.count_nonzero(got[:, N:]) == 0
        scale = ref.abs().max().item()
        assert (got[:, :N] - ref).abs().max().item() <= 1e-3 * scale
