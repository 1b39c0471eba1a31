"""GPU tier: the split-fp16 fused forward (pqlb_mlp_forward_h, csrc/mlp_fwd_h.cu) against an fp64
evaluation of the same three Linear+ELU layers (pql/models/mlp.py:15-24) and heads.

terms = 3 (hi/lo split of both operands, three MMAs per product) must reproduce the fp32 reference
arithmetic to ~1e-6; terms = 1 (hi halves only) has the accuracy of one TF32 MMA per product.  The
stored activations are the TF32 rounding of the (accurate) values; repeated launches must be
bit-identical (a race on the in-place TMEM / shared-memory conversions shows up as differing words)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def K():
    from pql_b200 import _kernels
    return _kernels


@pytest.fixture(params=[1, 2], ids=["cta-per-tile", "persistent"])
def mode(request):
    """Both schedules of the kernel: one CTA per (network, tile) and persistent CTAs pipelining tiles."""
    from pql_b200 import _lib
    _lib.load().pqlb_mlp_forward_h_mode(request.param)
    yield request.param
    _lib.load().pqlb_mlp_forward_h_mode(0)


def rn_tf32(x):
    return ((x.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def make_net(k_in, out_dim, g, ld=None):
    """fp32 weights in the arena layout of one net + their fp16 hi | lo copies (pqlb_split_f16)."""
    from pql_b200 import _lib
    ld = ld or (k_in + 3) // 4 * 4
    dims = [(512, k_in, ld), (256, 512, 512), (128, 256, 256), (out_dim, 128, 128)]
    ws, bs, hs, ls = [], [], [], []
    for o, i, l in dims:
        w = torch.zeros(o, l, device=DEV)
        w[:, :i] = (torch.rand(o, i, device=DEV, generator=g) * 2 - 1) / i ** 0.5
        b = (torch.rand(o, device=DEV, generator=g) * 2 - 1) / i ** 0.5
        hi = torch.zeros(o * l, dtype=torch.float16, device=DEV)
        lo = torch.zeros(o * l, dtype=torch.float16, device=DEV)
        _lib.call("pqlb_split_f16", _lib.ptr(w), _lib.ptr(hi), _lib.ptr(lo), w.numel())
        ws.append(w); bs.append(b); hs.append(hi); ls.append(lo)
    return ws, bs, hs, ls


def ref_trunk(x, ws, bs, k_in):
    h = x[:, :k_in].double()
    outs = []
    for l in range(3):
        w = ws[l][:, :h.shape[1]].double()
        z = h @ w.t() + bs[l].double()
        h = torch.where(z > 0, z, torch.expm1(z))
        outs.append(h)
    return outs


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def test_split_copies_reconstruct_the_weights():
    from pql_b200 import _lib
    g = torch.Generator(device=DEV).manual_seed(0)
    w = torch.randn(4099, device=DEV, generator=g) * 0.05
    hi = torch.zeros(4100, dtype=torch.float16, device=DEV)
    lo = torch.zeros(4100, dtype=torch.float16, device=DEV)
    _lib.call("pqlb_split_f16", _lib.ptr(w), _lib.ptr(hi), _lib.ptr(lo), w.numel())
    s = float(_lib.load().pqlb_f16_weight_scale())
    assert s == 256.0
    back = (hi[:4099].double() + lo[:4099].double()) / s
    assert (back - w.double()).abs().max().item() <= 2.0 ** -21 * w.abs().max().item()
    assert torch.equal(hi[:4099], (w * s).half())


@pytest.mark.parametrize("M,k_in,n_groups,terms", [(128, 104, 1, 3), (200, 88, 1, 1), (8192, 104, 4, 3), (300, 104, 2, 3),
                                                   (8192, 104, 5, 3), (1000, 16, 1, 3), (8192, 88, 2, 1), (40000, 104, 2, 3),
                                                   # wide inputs (129..256 columns: mlp_fwd_hw_kernel; ShadowHand critics = 231)
                                                   (128, 231, 1, 3), (8192, 231, 4, 3), (300, 232, 2, 3), (1000, 231, 1, 1),
                                                   (700, 160, 2, 3), (450, 256, 2, 3)])
def test_split_f16_trunk_and_q_head(K, mode, M, k_in, n_groups, terms):
    if k_in > 128 and mode == 2:
        pytest.skip("wide inputs always take the CTA-per-tile kernel")
    g = torch.Generator(device=DEV).manual_seed(11 * M + k_in + terms)
    ld = (k_in + 3) // 4 * 4
    groups, keep = [], []
    for i in range(n_groups):
        x = torch.zeros(M, ld, device=DEV)
        x[:, :k_in] = torch.randn(M, k_in, device=DEV, generator=g) * 1.5
        ws, bs, hs, ls = make_net(k_in, 1, g)
        h = [torch.zeros(M, n, device=DEV) for n in (512, 256, 128)]
        q = torch.zeros(M, device=DEV)
        store = i % 2 == 0
        grp = dict(x=K.addr(x), ldx=ld, w1h=hs[0].data_ptr(), ldw1=ld, w2h=hs[1].data_ptr(), w3h=hs[2].data_ptr(),
                   b1=K.addr(bs[0]), b2=K.addr(bs[1]), b3=K.addr(bs[2]), head_w=K.addr(ws[3]), head_b=K.addr(bs[3]),
                   q=K.addr(q), terms=terms)
        if terms == 3:
            grp.update(w1l=ls[0].data_ptr(), w2l=ls[1].data_ptr(), w3l=ls[2].data_ptr())
        if store:
            grp.update(h1=K.addr(h[0]), h2=K.addr(h[1]), h3=K.addr(h[2]))
        groups.append(grp)
        keep.append((x, ws, bs, hs, ls, h, q, store))
    call = K.MlpForwardH(M, k_in, groups)
    first = None
    for rep in range(3):
        call()
        torch.cuda.synchronize()
        sig = [t.clone() for (_, _, _, _, _, h, q, _) in keep for t in (*h, q)]
        if first is None:
            first = sig
        else:
            for a, b in zip(first, sig):
                assert torch.equal(a.view(torch.int32), b.view(torch.int32)), f"launch {rep} is not bit-identical to launch 0"
    tol_h, tol_q = (1e-5, 1e-5) if terms == 3 else (1.2e-3, 1.2e-3)
    for gi, (x, ws, bs, hs, ls, h, q, store) in enumerate(keep):
        r1, r2, r3 = ref_trunk(x, ws, bs, k_in)
        q_ref = r3 @ ws[3][0, :128].double() + bs[3].double()
        e = rel(q, q_ref)
        assert e <= tol_q, f"group {gi}: q error {e:.2e} (terms {terms})"
        if store:
            for name, got, ref in (("h1", h[0], r1), ("h2", h[1], r2), ("h3", h[2], r3)):
                if terms == 3:      # stored = rn_tf32(accurate value): one-ulp flips where the value sits on a rounding boundary
                    want = rn_tf32(ref.float())
                    bad = (got.view(torch.int32) != want.view(torch.int32)).float().mean().item()
                    assert bad <= 3e-2, f"group {gi} {name}: {bad:.2e} of the stored words differ from rn_tf32(reference)"
                    assert rel(got, ref) <= 3e-4
                else:
                    assert rel(got, ref) <= tol_h, f"group {gi} {name}: {rel(got, ref):.2e}"
        else:
            assert all(float(t.abs().max()) == 0.0 for t in h), "activations written although not requested"


@pytest.mark.parametrize("M,A,noisy,terms,k_in", [(8192, 16, True, 1, 88), (300, 16, False, 3, 88), (1000, 8, True, 1, 88),
                                                  (128, 4, False, 1, 88), (1000, 16, True, 1, 208), (300, 12, False, 3, 200),
                                                  # wide-input kernel: up to 32 actions, output rows of any alignment
                                                  (8192, 20, True, 1, 211), (300, 32, False, 3, 232), (1000, 20, False, 1, 160)])
def test_split_f16_policy_head(K, mode, M, A, noisy, terms, k_in):
    """tanh(Linear(128, A)) (+ clipped N(0, std^2) noise, clamp) fused behind the trunk: act_out2 holds the
    value, act_out its TF32 rounding (mlp.py:177-179, noise.py:19-27).  k_in > 128: the wide-input kernel."""
    if k_in > 128 and mode == 2:
        pytest.skip("wide inputs always take the CTA-per-tile kernel")
    g = torch.Generator(device=DEV).manual_seed(M + A)
    ld = (k_in + 7) // 8 * 8
    x = torch.zeros(M, ld, device=DEV)
    x[:, :k_in] = torch.randn(M, k_in, device=DEV, generator=g)
    ws, bs, hs, ls = make_net(k_in, A, g, ld)
    # the action columns sit at `off` of rows of W floats; 20 actions behind a wide input: off = 87 (rows not 16-byte aligned)
    W, off = 124, 87 if (k_in > 128 and A == 20) else 88
    out = torch.full((M, W), 7.0, device=DEV)
    out2 = torch.full((M, W), 7.0, device=DEV)
    noise = torch.randn(M, A, device=DEV, generator=g)
    grp = dict(x=K.addr(x), ldx=ld, w1h=hs[0].data_ptr(), ldw1=ld, w2h=hs[1].data_ptr(), w3h=hs[2].data_ptr(),
               b1=K.addr(bs[0]), b2=K.addr(bs[1]), b3=K.addr(bs[2]), terms=terms,
               act_wh=hs[3].data_ptr(), act_b=K.addr(bs[3]), act_n=A, act_out=K.addr(out, off), act_ldo=W,
               act_out2=K.addr(out2, off), act_ldo2=W)
    if terms == 3:
        grp.update(w1l=ls[0].data_ptr(), w2l=ls[1].data_ptr(), w3l=ls[2].data_ptr(), act_wl=ls[3].data_ptr())
    if noisy:
        grp.update(act_noise=K.addr(noise), act_ldnoise=A, noise_std=0.8, noise_bound=0.2)
    K.MlpForwardH(M, k_in, [grp])()
    torch.cuda.synchronize()
    h3 = ref_trunk(x, ws, bs, k_in)[2]
    a = torch.tanh(h3 @ ws[3][:, :128].double().t() + bs[3].double())
    if noisy:
        a = torch.clamp(a + torch.clamp(noise.double() * 0.8, -0.2, 0.2), -1.0, 1.0)
    got2, got = out2[:, off:off + A], out[:, off:off + A]
    tol = 1e-5 if terms == 3 else 1.2e-3
    assert rel(got2, a) <= tol, f"{rel(got2, a):.2e}"
    assert torch.equal(got, rn_tf32(got2))
    for o in (out, out2):       # nothing but the action columns is written
        assert float((o[:, :off] - 7.0).abs().max()) == 0.0 and float((o[:, off + A:] - 7.0).abs().max()) == 0.0


@pytest.mark.parametrize("M", [8192, 1000])
def test_policy_to_critic_dependency_in_one_launch(K, mode, M):
    """One launch: a policy net (terms 1) publishes its action rows tile by tile, a critic that does not depend
    on it runs meanwhile, two critics wait for the tile they read.  Must equal the two-launch sequence bit for
    bit, on every repetition (tile_sync carries over between launches without a reset)."""
    g = torch.Generator(device=DEV).manual_seed(M)
    O, A, ld = 88, 16, 104
    xa = torch.zeros(M, ld, device=DEV)
    xa[:, :O] = torch.randn(M, O, device=DEV, generator=g)
    xc = torch.randn(M, ld, device=DEV, generator=g)
    actor = make_net(O, A, g)
    crit = [make_net(O + A, 1, g, ld) for _ in range(3)]
    noise = torch.randn(M, A, device=DEV, generator=g)

    def groups(x_wait, q, with_flags):
        ws, bs, hs, ls = actor
        ga = dict(x=K.addr(x_wait), ldx=ld, k_in=O, w1h=hs[0].data_ptr(), ldw1=88, w2h=hs[1].data_ptr(), w3h=hs[2].data_ptr(),
                  b1=K.addr(bs[0]), b2=K.addr(bs[1]), b3=K.addr(bs[2]), terms=1, act_wh=hs[3].data_ptr(), act_b=K.addr(bs[3]),
                  act_n=A, act_out2=K.addr(x_wait, O), act_ldo2=ld, act_noise=K.addr(noise), act_ldnoise=A, noise_std=0.8,
                  noise_bound=0.2, publish=int(with_flags))
        gc = []
        for i, (ws, bs, hs, ls) in enumerate(crit):
            gc.append(dict(x=K.addr(xc if i == 0 else x_wait), ldx=ld, k_in=O + A, w1h=hs[0].data_ptr(), w1l=ls[0].data_ptr(), ldw1=ld,
                           w2h=hs[1].data_ptr(), w2l=ls[1].data_ptr(), w3h=hs[2].data_ptr(), w3l=ls[2].data_ptr(),
                           b1=K.addr(bs[0]), b2=K.addr(bs[1]), b3=K.addr(bs[2]), head_w=K.addr(ws[3]), head_b=K.addr(bs[3]),
                           q=K.addr(q[i]), terms=3, wait=int(with_flags and i > 0)))
        return ga, gc

    x_ref, x_one = xa.clone(), xa.clone()
    q_ref = [torch.zeros(M, device=DEV) for _ in range(3)]
    q_one = [torch.zeros(M, device=DEV) for _ in range(3)]
    ga, gc = groups(x_ref, q_ref, False)
    K.MlpForwardH(M, O, [ga])()
    K.MlpForwardH(M, O + A, gc)()
    sync = torch.zeros(2 + (M + 127) // 128, dtype=torch.int32, device=DEV)
    ga, gc = groups(x_one, q_one, True)
    fused = K.MlpForwardH(M, O + A, [ga] + gc, tile_sync=sync)
    for rep in range(4):
        x_one[:, O:].zero_()
        for t in q_one:
            t.zero_()
        fused()
        torch.cuda.synchronize()
        assert torch.equal(x_one, x_ref), f"rep {rep}: action rows differ"
        for i in range(3):
            assert torch.equal(q_one[i], q_ref[i]), f"rep {rep}: critic {i} differs"
        assert int(sync[1]) == rep + 1 and int(sync[0]) == 0
