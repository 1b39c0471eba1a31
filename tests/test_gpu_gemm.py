"""GPU tier: the tcgen05 TF32 grouped GEMM (pqlb_gemm_tf32) against an fp64 matmul of the
same TF32-rounded operands.  With pre-rounded operands every product is exact and only the
fp32 accumulation order differs, so the tolerance is 2e-5 of the output's largest magnitude."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def rn_tf32(x):
    """cvt.rna.tf32.f32: round to nearest (ties away) to a 10-bit mantissa."""
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def mk(shape, g, scale=1.0):
    return rn_tf32(torch.randn(*shape, generator=g, device=DEV) * scale)


def elu(x):
    return torch.where(x > 0, x, torch.expm1(x))


def check(out, ref, tol=2e-5, what=""):
    scale = ref.abs().max().item() + 1e-30
    err = (out.double() - ref.double()).abs().max().item() / scale
    assert err <= tol, f"{what}: max err {err:.3e} of max |ref|"


@pytest.fixture(scope="module")
def K():
    from pql_b200 import _kernels
    return _kernels


@pytest.mark.parametrize("M,N,Kdim,tile_n", [(128, 128, 32, 128), (300, 512, 104, 256), (256, 256, 512, 128),
                                             (1000, 128, 256, 64), (130, 512, 231, 256), (128, 16, 128, 16),
                                             (128, 48, 64, 64)])
def test_forward_bias_elu(K, M, N, Kdim, tile_n):
    g = torch.Generator(device=DEV).manual_seed(M + N + Kdim)
    ld = (Kdim + 3) // 4 * 4
    a = torch.zeros(M, ld, device=DEV); a[:, :Kdim] = mk((M, Kdim), g)
    w = torch.zeros(N, ld, device=DEV); w[:, :Kdim] = mk((N, Kdim), g, 0.1)
    bias = torch.randn(N, generator=g, device=DEV)
    out = torch.full((M, N), 7.0, device=DEV)
    K.Gemm(M, N, Kdim, [dict(a=K.addr(a), lda=ld, b=K.addr(w), ldb=ld, bias=K.addr(bias), out=K.addr(out), ldo=N)],
           epilogue=K.EPI_BIAS_ELU, tile_n=tile_n)()
    ref = rn_tf32(elu((a[:, :Kdim].double() @ w[:, :Kdim].double().t()).float() + bias))
    torch.cuda.synchronize()
    check(out, ref, 1e-3, "bias_elu")          # outputs are TF32-rounded: half-ulp = 2^-11 relative
    out2 = torch.zeros(M, N, device=DEV)
    K.Gemm(M, N, Kdim, [dict(a=K.addr(a), lda=ld, b=K.addr(w), ldb=ld, bias=K.addr(bias), out=K.addr(out2), ldo=N)],
           epilogue=K.EPI_BIAS, tile_n=tile_n)()
    check(out2, a[:, :Kdim].double() @ w[:, :Kdim].double().t() + bias.double(), 2e-5, "bias")


def test_grouped_head_epilogue(K):
    """Four groups in one launch, scalar Q head fused into the layer-3 epilogue."""
    M, N, Kd = 384, 128, 256
    g = torch.Generator(device=DEV).manual_seed(5)
    groups, refs, keep = [], [], []
    for i in range(4):
        a, w = mk((M, Kd), g), mk((N, Kd), g, 0.1)
        bias, hw, hb = torch.randn(N, generator=g, device=DEV), torch.randn(N, generator=g, device=DEV), torch.randn(1, generator=g, device=DEV)
        q = torch.zeros(M, device=DEV)
        out = torch.zeros(M, N, device=DEV) if i % 2 == 0 else None
        groups.append(dict(a=K.addr(a), lda=Kd, b=K.addr(w), ldb=Kd, bias=K.addr(bias), head_w=K.addr(hw),
                           head_b=K.addr(hb), q=K.addr(q), out=K.addr(out), ldo=N))
        h = elu((a.double() @ w.double().t()).float() + bias)
        refs.append((h, h.double() @ hw.double() + hb.double(), q, out))
        keep += [a, w, bias, hw, hb]
    K.Gemm(M, N, Kd, groups, epilogue=K.EPI_BIAS_ELU_HEAD, tile_n=128)()
    torch.cuda.synchronize()
    for h, qref, q, out in refs:
        check(q, qref, 2e-5, "head q")
        if out is not None:
            check(out, rn_tf32(h), 1e-3, "head h3")


def test_dgrad_mn_major_b_with_elugrad(K):
    """dz_prev = (dz . W) * elu'(h): A K-major, B = W[out,in] read as [K][N] (MN-major)."""
    M, n_out, n_in = 260, 128, 256
    g = torch.Generator(device=DEV).manual_seed(6)
    dz, w = mk((M, n_out), g), mk((n_out, n_in), g, 0.1)
    h = rn_tf32(elu(torch.randn(M, n_in, generator=g, device=DEV)))
    out = torch.zeros(M, n_in, device=DEV)
    K.Gemm(M, n_in, n_out, [dict(a=K.addr(dz), lda=n_out, b=K.addr(w), ldb=n_in, aux=K.addr(h), ldaux=n_in,
                                  out=K.addr(out), ldo=n_in)],
           epilogue=K.EPI_MUL_ELUGRAD, tile_n=256, b_major=K.MN_MAJOR)()
    ref = (dz.double() @ w.double()).float() * torch.where(h > 0, torch.ones_like(h), h + 1)
    check(out, rn_tf32(ref), 1e-3, "dgrad")


def test_dgrad_two_operand_pairs_and_column_window(K):
    """P-learner layer-1 dgrad: sum over both critics (a2/b2 continuation), only the action
    columns [O, O+A) stored, tanh' epilogue."""
    M, O, A, H = 200, 211, 20, 512
    ldw = 232
    g = torch.Generator(device=DEV).manual_seed(7)
    dz1, dz2 = mk((M, H), g), mk((M, H), g)
    w1 = torch.zeros(H, ldw, device=DEV); w1[:, :O + A] = mk((H, O + A), g, 0.1)
    w2 = torch.zeros(H, ldw, device=DEV); w2[:, :O + A] = mk((H, O + A), g, 0.1)
    act = torch.tanh(torch.randn(M, A, generator=g, device=DEV))
    out = torch.zeros(M, A, device=DEV)
    K.Gemm(M, O + A, H, [dict(a=K.addr(dz1), lda=H, b=K.addr(w1), ldb=ldw, a2=K.addr(dz2), lda2=H, b2=K.addr(w2),
                              ldb2=ldw, aux=K.addr(act, -O), ldaux=A, out=K.addr(out), ldo=A)],
           epilogue=K.EPI_MUL_TANHGRAD, tile_n=256, b_major=K.MN_MAJOR, K2=H, col_lo=O, col_hi=O + A)()
    ref = ((dz1.double() @ w1.double() + dz2.double() @ w2.double())[:, O:O + A]).float() * (1 - act * act)
    check(out, rn_tf32(ref), 1e-3, "dgrad2")


@pytest.mark.parametrize("n_out,n_in,B,tile_n,splits", [(512, 104, 1024, 64, 4), (128, 256, 2048, 64, 16),
                                                        (256, 512, 512, 128, 2), (16, 128, 512, 128, 1),
                                                        (51, 128, 256, 128, 2), (512, 231, 384, 128, 4)])
def test_wgrad_mn_mn_split_k(K, n_out, n_in, B, tile_n, splits):
    """dW[out,in] = dz^T . h with both operands MN-major and the contraction over the batch."""
    g = torch.Generator(device=DEV).manual_seed(n_out + n_in)
    ldz = (n_out + 3) // 4 * 4 if n_out != 51 else 64
    ldh = (n_in + 3) // 4 * 4
    dz = torch.zeros(B, ldz, device=DEV); dz[:, :n_out] = mk((B, n_out), g)
    h = torch.zeros(B, ldh, device=DEV); h[:, :n_in] = mk((B, n_in), g)
    stride = n_out * ldh
    part = torch.zeros(splits, n_out, ldh, device=DEV)
    K.Gemm(n_out, n_in, B, [dict(a=K.addr(dz), lda=ldz, b=K.addr(h), ldb=ldh, out=K.addr(part), ldo=ldh,
                                  split_stride=stride)],
           epilogue=K.EPI_STORE, tile_n=tile_n, a_major=K.MN_MAJOR, b_major=K.MN_MAJOR, splits=splits)()
    ref = dz[:, :n_out].double().t() @ h[:, :n_in].double()
    check(part.sum(0)[:, :n_in], ref, 2e-5, "wgrad")
    assert torch.count_nonzero(part[:, :, n_in:]) == 0


@pytest.mark.parametrize("shapes,B", [([(128, 256), (256, 512), (512, 104)] * 2, 8192),
                                      ([(16, 128), (128, 256), (256, 512), (512, 88)], 8192),
                                      ([(51, 128), (128, 256), (256, 512), (512, 104)] * 2, 1024),
                                      ([(128, 256), (256, 512), (512, 231)], 416),
                                      ([(20, 128), (512, 211)], 64)])
def test_wgrad_multi_one_launch(K, shapes, B):
    """pqlb_wgrad_multi: every dW = dz^T . h of an update in one launch, uneven split-K ranges
    (wgrad_plan), partial tiles summed in split order."""
    g = torch.Generator(device=DEV).manual_seed(B + len(shapes))
    plan = K.wgrad_plan(shapes, B)
    problems, checks = [], []
    for (n_out, n_in), (tile_n, splits) in zip(shapes, plan):
        ldz = 64 if n_out == 51 else (n_out + 3) // 4 * 4
        ldh = (n_in + 3) // 4 * 4
        dz = torch.zeros(B, ldz, device=DEV); dz[:, :n_out] = mk((B, n_out), g)
        h = torch.zeros(B, ldh, device=DEV); h[:, :n_in] = mk((B, n_in), g)
        part = torch.full((splits, n_out, ldh), 3.0, device=DEV)
        part[:, :, n_in:] = 0
        problems.append(dict(dz=K.addr(dz), lddz=ldz, h=K.addr(h), ldh=ldh, part=K.addr(part), ldo=ldh,
                             split_stride=n_out * ldh, M=n_out, N=n_in, tile_n=tile_n, splits=splits))
        checks.append((dz, h, part, n_out, n_in))
    assert sum(-(-m // 128) * -(-n // t) * s for (m, n), (t, s) in zip(shapes, plan)) <= 148
    K.WgradMulti(B, problems)()
    torch.cuda.synchronize()
    for dz, h, part, n_out, n_in in checks:
        ref = dz[:, :n_out].double().t() @ h[:, :n_in].double()
        check(part.sum(0)[:, :n_in], ref, 2e-5, f"wgrad_multi {n_out}x{n_in}")
        assert torch.count_nonzero(part[:, :, n_in:]) == 0


@pytest.mark.parametrize("n_out,n_in,B,tile_n,splits,cluster,groups", [
    (256, 512, 8192, 256, 32, 8, 2), (256, 512, 8192, 256, 32, 4, 2), (128, 256, 8192, 256, 64, 8, 2),
    (512, 104, 8192, 128, 32, 8, 2), (512, 104, 8192, 128, 32, 2, 1), (16, 128, 2048, 128, 8, 4, 1),
    (51, 128, 1024, 128, 4, 4, 2), (512, 231, 1024, 256, 8, 8, 1), (128, 256, 2048, 64, 16, 8, 1)])
def test_wgrad_cluster_reduction_equals_split_order_sum(K, n_out, n_in, B, tile_n, splits, cluster, groups):
    """Split-K partials summed through distributed shared memory inside a thread-block cluster
    (pqlb_gemm_desc.cluster): partial p must equal, bit for bit, the per-split partials
    [p * cluster, (p + 1) * cluster) of the un-clustered launch added in split order."""
    g = torch.Generator(device=DEV).manual_seed(n_out * 7 + n_in + cluster)
    ldz = (n_out + 3) // 4 * 4 if n_out != 51 else 64
    ldh = (n_in + 3) // 4 * 4
    stride = n_out * ldh
    dz, h, part, red = [], [], [], []
    for _ in range(groups):
        a = torch.zeros(B, ldz, device=DEV); a[:, :n_out] = mk((B, n_out), g)
        b = torch.zeros(B, ldh, device=DEV); b[:, :n_in] = mk((B, n_in), g)
        dz.append(a); h.append(b)
        part.append(torch.zeros(splits, n_out, ldh, device=DEV))
        red.append(torch.full((splits // cluster, n_out, ldh), 3.0, device=DEV))
    def launch(outs, c):
        K.Gemm(n_out, n_in, B, [dict(a=K.addr(dz[i]), lda=ldz, b=K.addr(h[i]), ldb=ldh, out=K.addr(outs[i]), ldo=ldh,
                                      split_stride=stride) for i in range(groups)],
               epilogue=K.EPI_STORE, tile_n=tile_n, a_major=K.MN_MAJOR, b_major=K.MN_MAJOR, splits=splits, cluster=c)()
    launch(part, 1)
    launch(red, cluster)
    torch.cuda.synchronize()
    for i in range(groups):
        exp = part[i][0::cluster].clone()
        for s in range(1, cluster):
            exp += part[i][s::cluster]
        assert torch.equal(red[i][:, :, :n_in], exp[:, :, :n_in]), f"group {i}"
        ref = dz[i][:, :n_out].double().t() @ h[i][:, :n_in].double()
        check(red[i].sum(0)[:, :n_in], ref, 2e-5, "wgrad cluster")


def test_actor_head_tanh_and_noise(K):
    M, Kd, A, O, x_ld = 300, 128, 16, 88, 104
    g = torch.Generator(device=DEV).manual_seed(9)
    h, w = mk((M, Kd), g), mk((A, Kd), g, 0.3)
    bias = torch.randn(A, generator=g, device=DEV) * 0.1
    z01 = torch.randn(M, A, generator=g, device=DEV)
    noise = z01 * 0.8
    x = torch.zeros(M, x_ld, device=DEV)
    K.Gemm(M, A, Kd, [dict(a=K.addr(h), lda=Kd, b=K.addr(w), ldb=Kd, bias=K.addr(bias), aux=K.addr(z01), ldaux=A,
                           out=K.addr(x, O), ldo=x_ld)],
           epilogue=K.EPI_BIAS_TANH_NOISE, tile_n=16, noise_bound=0.2, noise_std=0.8)()
    a = torch.tanh((h.double() @ w.double().t()).float() + bias)
    ref = torch.clamp(a + torch.clamp(noise, -0.2, 0.2), -1, 1)
    check(x[:, O:O + A], rn_tf32(ref), 1e-3, "tanh_noise")
    assert torch.count_nonzero(x[:, :O]) == 0
    act = torch.zeros(M, A, device=DEV)
    K.Gemm(M, A, Kd, [dict(a=K.addr(h), lda=Kd, b=K.addr(w), ldb=Kd, bias=K.addr(bias), out=K.addr(x, O), ldo=x_ld,
                           out2=K.addr(act), ldo2=A)], epilogue=K.EPI_BIAS_TANH, tile_n=16)()
    check(act, a, 2e-5, "tanh fp32 copy")
    check(x[:, O:O + A], rn_tf32(a), 1e-3, "tanh")


def test_softmax_head(K):
    M, Kd, N = 260, 128, 51
    g = torch.Generator(device=DEV).manual_seed(10)
    h, w = mk((M, Kd), g), mk((N, Kd), g, 0.3)
    bias = torch.randn(N, generator=g, device=DEV) * 0.1
    p = torch.zeros(M, 64, device=DEV)
    K.Gemm(M, N, Kd, [dict(a=K.addr(h), lda=Kd, b=K.addr(w), ldb=Kd, bias=K.addr(bias), out=K.addr(p), ldo=64)],
           epilogue=K.EPI_BIAS_SOFTMAX, tile_n=64)()
    ref = torch.softmax((h.double() @ w.double().t()).float() + bias, dim=1)
    check(p[:, :N], ref, 2e-5, "softmax")
    assert torch.count_nonzero(p[:, N:]) == 0


def test_round_tf32(K):
    from pql_b200 import _lib
    x = torch.randn(100003, device=DEV)
    y = torch.empty_like(x)
    _lib.call("pqlb_round_tf32", _lib.ptr(x), _lib.ptr(y), x.numel())
    assert torch.equal(y, rn_tf32(x))


@pytest.mark.parametrize("M,k_in,n_groups", [(128, 104, 1), (200, 88, 1), (8192, 104, 4), (300, 104, 2)])
def test_fused_trunk_forward(K, M, k_in, n_groups):
    """pqlb_mlp_forward: three Linear+ELU layers with activations kept in tensor memory, optional
    stores of h1/h2/h3 and the fused scalar head, against an fp64 evaluation with the same TF32
    roundings (h rounded after every ELU)."""
    g = torch.Generator(device=DEV).manual_seed(M + k_in)
    ld = (k_in + 3) // 4 * 4
    groups, checks, keep = [], [], []
    for i in range(n_groups):
        x = torch.zeros(M, ld, device=DEV); x[:, :k_in] = mk((M, k_in), g)
        w1 = torch.zeros(512, ld, device=DEV); w1[:, :k_in] = mk((512, k_in), g, 1.0 / k_in ** 0.5)
        w2, w3 = mk((256, 512), g, 1.0 / 512 ** 0.5), mk((128, 256), g, 1.0 / 16)
        b1, b2, b3 = (torch.randn(n, generator=g, device=DEV) * 0.1 for n in (512, 256, 128))
        w4, b4 = torch.randn(128, generator=g, device=DEV) * 0.1, torch.randn(1, generator=g, device=DEV)
        store = i % 2 == 0
        h = [torch.zeros(M, n, device=DEV) if store else None for n in (512, 256, 128)]
        q = torch.zeros(M, device=DEV) if (i < 2 or not store) else None
        groups.append(dict(x=K.addr(x), ldx=ld, w1=K.addr(w1), ldw1=ld, w2=K.addr(w2), w3=K.addr(w3), b1=K.addr(b1),
                           b2=K.addr(b2), b3=K.addr(b3), head_w=K.addr(w4), head_b=K.addr(b4), q=K.addr(q),
                           h1=K.addr(h[0]), h2=K.addr(h[1]), h3=K.addr(h[2])))
        r1 = rn_tf32(elu((x[:, :k_in].double() @ w1[:, :k_in].double().t()).float() + b1))
        r2 = rn_tf32(elu((r1.double() @ w2.double().t()).float() + b2))
        z3 = elu((r2.double() @ w3.double().t()).float() + b3)
        checks.append((h, q, (r1, r2, rn_tf32(z3)), z3.double() @ w4.double() + b4.double()))
        keep += [x, w1, w2, w3, b1, b2, b3, w4, b4]
    K.MlpForward(M, k_in, groups)()
    torch.cuda.synchronize()
    for h, q, refs, qref in checks:
        if q is not None:
            check(q, qref, 3e-4, "fused head q")     # rounding flips of h1/h2 (one TF32 ulp) propagate
        for got, ref, name in zip(h, refs, ("h1", "h2", "h3")):
            if got is not None:
                check(got, ref, 1e-3, "fused " + name)


@pytest.mark.parametrize("M,k_in,n_groups,cluster", [(128, 104, 1, 0), (8192, 104, 4, 0), (16384, 104, 4, 2), (8192, 88, 1, 1),
                                                     (300, 104, 2, 2), (700, 88, 3, 2), (8192, 104, 2, 1)])
def test_fused_trunk_equals_layerwise(K, M, k_in, n_groups, cluster):
    """The layer-fused trunk accumulates every output element over k in the same order as the
    per-layer GEMMs (k-blocks ascending into one fp32 TMEM accumulator), so h1/h2/h3 must be
    BIT-IDENTICAL between the two paths — a sharper check than any tolerance: a race on the
    in-place TMEM conversion or a staging-buffer hazard shows up as a handful of differing words."""
    g = torch.Generator(device=DEV).manual_seed(7 * M + k_in)
    ld = (k_in + 3) // 4 * 4
    groups, lw, outs = [], [[], [], []], []
    for i in range(n_groups):
        x = torch.zeros(M, ld, device=DEV); x[:, :k_in] = mk((M, k_in), g)
        w1 = torch.zeros(512, ld, device=DEV); w1[:, :k_in] = mk((512, k_in), g, 1.0 / k_in ** 0.5)
        w2, w3 = mk((256, 512), g, 1.0 / 512 ** 0.5), mk((128, 256), g, 1.0 / 16)
        b1, b2, b3 = (torch.randn(n, generator=g, device=DEV) * 0.1 for n in (512, 256, 128))
        hf = [torch.zeros(M, n, device=DEV) for n in (512, 256, 128)]
        hl = [torch.zeros(M, n, device=DEV) for n in (512, 256, 128)]
        groups.append(dict(x=K.addr(x), ldx=ld, w1=K.addr(w1), ldw1=ld, w2=K.addr(w2), w3=K.addr(w3), b1=K.addr(b1),
                           b2=K.addr(b2), b3=K.addr(b3), h1=K.addr(hf[0]), h2=K.addr(hf[1]), h3=K.addr(hf[2])))
        lw[0].append(dict(a=K.addr(x), lda=ld, b=K.addr(w1), ldb=ld, bias=K.addr(b1), out=K.addr(hl[0]), ldo=512))
        lw[1].append(dict(a=K.addr(hl[0]), lda=512, b=K.addr(w2), ldb=512, bias=K.addr(b2), out=K.addr(hl[1]), ldo=256))
        lw[2].append(dict(a=K.addr(hl[1]), lda=256, b=K.addr(w3), ldb=256, bias=K.addr(b3), out=K.addr(hl[2]), ldo=128))
        outs.append((hf, hl, (x, w1, w2, w3, b1, b2, b3)))
    from pql_b200 import _lib
    for rep in range(3):                       # repeated launches: a race need not fire every time
        _lib.load().pqlb_mlp_forward_cluster(cluster)     # 0 = default, 1 = one CTA per tile, 2 = CTA pairs (M = 300: odd tile count -> ghost CTA)
        K.MlpForward(M, k_in, groups)()
        _lib.load().pqlb_mlp_forward_cluster(0)
        K.Gemm(M, 512, k_in, lw[0], epilogue=K.EPI_BIAS_ELU, tile_n=256)()
        K.Gemm(M, 256, 512, lw[1], epilogue=K.EPI_BIAS_ELU, tile_n=128)()
        K.Gemm(M, 128, 256, lw[2], epilogue=K.EPI_BIAS_ELU, tile_n=128)()
        torch.cuda.synchronize()
        for gi, (hf, hl, _) in enumerate(outs):
            for name, a, b in zip(("h1", "h2", "h3"), hf, hl):
                bad = (a.view(torch.int32) != b.view(torch.int32))
                n_bad = int(bad.sum().item())
                if n_bad:
                    idx = bad.nonzero()[:8].tolist()
                    raise AssertionError(f"rep {rep} group {gi} {name}: {n_bad} words differ, first at {idx}, "
                                         f"fused {a[bad][:4].tolist()} layerwise {b[bad][:4].tolist()}")


@pytest.mark.parametrize("M,n_groups,with_bias", [(128, 1, True), (8192, 2, True), (16384, 2, False), (300, 3, True), (8192, 1, True)])
def test_fused_dgrad_chain_equals_layerwise(K, M, n_groups, with_bias):
    """pqlb_mlp_backward against the two per-layer dgrad GEMMs (EPI_MUL_ELUGRAD, weights read as
    [K][N]): same accumulation order, so dz2 / dz1 must be BIT-IDENTICAL; the fused bias-gradient
    partials against a float64 column sum of each 128-row tile."""
    g = torch.Generator(device=DEV).manual_seed(11 * M + n_groups)
    nblk = (M + 127) // 128
    fused, lw2, lw1, outs = [], [], [], []
    for i in range(n_groups):
        dz3 = mk((M, 128), g)
        w3, w2 = mk((128, 256), g, 1.0 / 16), mk((256, 512), g, 1.0 / 512 ** 0.5)
        h2 = rn_tf32(elu(torch.randn(M, 256, generator=g, device=DEV)))
        h1 = rn_tf32(elu(torch.randn(M, 512, generator=g, device=DEV)))
        f2, f1 = torch.zeros(M, 256, device=DEV), torch.zeros(M, 512, device=DEV)
        l2, l1 = torch.zeros(M, 256, device=DEV), torch.zeros(M, 512, device=DEV)
        p2 = torch.full((nblk, 256), 7.0, device=DEV) if with_bias else None
        p1 = torch.full((nblk, 512), 7.0, device=DEV) if with_bias else None
        fused.append(dict(dz3=K.addr(dz3), w3=K.addr(w3), w2=K.addr(w2), h2=K.addr(h2), h1=K.addr(h1), dz2=K.addr(f2),
                          dz1=K.addr(f1), bias_part2=K.addr(p2), bias_part1=K.addr(p1)))
        lw2.append(dict(a=K.addr(dz3), lda=128, b=K.addr(w3), ldb=256, aux=K.addr(h2), ldaux=256, out=K.addr(l2), ldo=256))
        lw1.append(dict(a=K.addr(l2), lda=256, b=K.addr(w2), ldb=512, aux=K.addr(h1), ldaux=512, out=K.addr(l1), ldo=512))
        outs.append((f2, f1, l2, l1, p2, p1, (dz3, w3, w2, h2, h1)))
    for rep in range(3):
        K.MlpBackward(M, fused)()
        K.Gemm(M, 256, 128, lw2, epilogue=K.EPI_MUL_ELUGRAD, tile_n=128, b_major=K.MN_MAJOR)()
        K.Gemm(M, 512, 256, lw1, epilogue=K.EPI_MUL_ELUGRAD, tile_n=256, b_major=K.MN_MAJOR)()
        torch.cuda.synchronize()
        for gi, (f2, f1, l2, l1, p2, p1, _) in enumerate(outs):
            for name, a, b in (("dz2", f2, l2), ("dz1", f1, l1)):
                bad = a.view(torch.int32) != b.view(torch.int32)
                n_bad = int(bad.sum().item())
                assert n_bad == 0, (f"rep {rep} group {gi} {name}: {n_bad} words differ, first at {bad.nonzero()[:6].tolist()}, "
                                    f"fused {a[bad][:4].tolist()} layerwise {b[bad][:4].tolist()}")
            if with_bias:
                for name, part, dz in (("bias2", p2, f2), ("bias1", p1, f1)):
                    pad = torch.zeros(nblk * 128, dz.shape[1], device=DEV, dtype=torch.float64)
                    pad[:M] = dz.double()
                    ref = pad.view(nblk, 128, -1).sum(1)
                    check(part, ref, 2e-6, f"rep {rep} group {gi} {name} partials")


@pytest.mark.parametrize("M,A,noisy", [(8192, 16, True), (300, 16, False), (1000, 8, True), (128, 4, False)])
def test_fused_policy_head_equals_separate_launch(K, M, A, noisy):
    """The tanh policy head as a fourth contraction inside pqlb_mlp_forward (h3 converted in place in
    TMEM, N = 16 MMA) against the separate head GEMM on the stored h3: same operands, same k order,
    so the actions (with and without clipped target-policy noise) and the fp32 tanh copy must be
    bit-identical."""
    g = torch.Generator(device=DEV).manual_seed(M + A)
    k_in, O, x_ld = 88, 88, 104
    x = torch.zeros(M, x_ld, device=DEV); x[:, :k_in] = mk((M, k_in), g)
    w1 = mk((512, k_in), g, 1.0 / k_in ** 0.5)
    w2, w3, w4 = mk((256, 512), g, 1.0 / 512 ** 0.5), mk((128, 256), g, 1.0 / 16), mk((A, 128), g, 0.3)
    b1, b2, b3, b4 = (torch.randn(n, generator=g, device=DEV) * 0.1 for n in (512, 256, 128, A))
    noise = torch.randn(M, A, generator=g, device=DEV)
    h3 = torch.zeros(M, 128, device=DEV)
    out_f, out_s = torch.zeros(M, x_ld, device=DEV), torch.zeros(M, x_ld, device=DEV)
    act_f, act_s = torch.zeros(M, A, device=DEV), torch.zeros(M, A, device=DEV)
    grp = dict(x=K.addr(x), ldx=x_ld, w1=K.addr(w1), ldw1=k_in, w2=K.addr(w2), w3=K.addr(w3), b1=K.addr(b1), b2=K.addr(b2),
               b3=K.addr(b3), h3=K.addr(h3), act_w=K.addr(w4), act_b=K.addr(b4), act_n=A, act_out=K.addr(out_f, O), act_ldo=x_ld)
    if noisy:
        grp.update(act_noise=K.addr(noise), act_ldnoise=A, noise_std=0.8, noise_bound=0.2)
    else:
        grp.update(act_out2=K.addr(act_f), act_ldo2=A)
    K.MlpForward(M, k_in, [grp])()
    hd = dict(a=K.addr(h3), lda=128, b=K.addr(w4), ldb=128, bias=K.addr(b4), out=K.addr(out_s, O), ldo=x_ld)
    if noisy:
        hd.update(aux=K.addr(noise), ldaux=A)
        K.Gemm(M, A, 128, [hd], epilogue=K.EPI_BIAS_TANH_NOISE, tile_n=K.pick_tile_n(A), noise_bound=0.2, noise_std=0.8)()
    else:
        hd.update(out2=K.addr(act_s), ldo2=A)
        K.Gemm(M, A, 128, [hd], epilogue=K.EPI_BIAS_TANH, tile_n=K.pick_tile_n(A))()
    torch.cuda.synchronize()
    assert torch.equal(out_f, out_s), f"{int((out_f != out_s).sum())} action words differ"
    assert torch.equal(act_f, act_s)
    assert out_f[:, O:O + A].abs().max() > 0.1 and torch.count_nonzero(out_f[:, :O]) == 0
