"""CPU tier: host-side logic above the C ABI that needs no GPU - parameter-arena geometry, launch
tiling heuristics, and the arithmetic of bench.py's roofline figures."""
import pytest
import torch

from pql_b200 import _kernels as K
from pql_b200.models import DistributionalDoubleQ, DoubleQ, TanhMLPPolicy
from pql_b200.models.mlp import HIDDEN, NetLayout


@pytest.mark.parametrize("in_dim,out_dim,n_nets,n_params", [(104, 1, 2, 436_226), (88, 16, 1, 211_856),
                                                            (231, 1, 2, 566_274), (211, 20, 1, 275_348),
                                                            (104, 51, 2, 449_126)])
def test_arena_layout_matches_reference_parameter_counts(in_dim, out_dim, n_nets, n_params):
    """Parameter counts of SURVEY §8(a8-a11); every tensor starts on a 32-word (128-byte) boundary,
    weight rows are padded to a multiple of 4 words (16-byte TMA rows), tensors do not overlap."""
    L = NetLayout(in_dim, out_dim, n_nets)
    assert L.n_params() == n_params
    spans = sorted((off, off + cnt) for _, _, _, off, cnt in L.tensors())
    assert all(off % 32 == 0 for off, _ in spans)
    assert all(a_end <= b_off for (_, a_end), (b_off, _) in zip(spans, spans[1:]))
    assert spans[-1][1] <= L.total and L.total % 32 == 0
    assert all(ld % 4 == 0 and ld >= d for ld, d in zip(L.ldw, L.dims[:-1]))


def test_module_parameters_are_views_of_the_arena():
    for m in (DoubleQ(9, 3), TanhMLPPolicy(9, 3), DistributionalDoubleQ(9, 3, device="cpu")):
        flat = m.arena.flat
        lo, hi = flat.data_ptr(), flat.data_ptr() + 4 * flat.numel()
        assert all(lo <= p.data_ptr() < hi for p in m.parameters())
        n = sum(p.numel() for p in m.parameters())
        assert n == m._layout.n_params()
        with torch.no_grad():
            next(m.parameters()).fill_(7.0)
        assert (flat == 7.0).sum() == next(m.parameters()).numel()


@pytest.mark.parametrize("M,N,B,groups", [(128, 256, 8192, 2), (256, 512, 8192, 2), (512, 104, 8192, 2), (51, 128, 16384, 2),
                                          (16, 128, 8192, 1), (512, 231, 8192, 2), (128, 256, 200, 1)])
def test_wgrad_tiling_invariants(M, N, B, groups):
    tile_n, splits = K.wgrad_tiling(M, N, B, groups)
    kb = (B + 31) // 32
    assert tile_n in (16, 32, 64, 128, 256) and splits >= 1 and kb % splits == 0
    c = K.wgrad_cluster(splits, tile_n)
    assert c in (1, 2, 4, 8) and splits % c == 0 and c <= max(1, K.WGRAD_CLUSTER)


def test_bench_flop_model_matches_survey():
    """bench.py's algorithmic FLOP constants against the layer shapes (3 684 352 FLOP per
    sample and critic update, 2 913 280 per actor update, AllegroHand)."""
    import bench
    O, A = bench.O, bench.A
    h1, h2, h3 = HIDDEN
    actor_fwd = O * h1 + h1 * h2 + h2 * h3 + h3 * A
    critic_fwd = (O + A) * h1 + h1 * h2 + h2 * h3 + h3
    critic_wgrad_dgrad = critic_fwd + (h1 * h2 + h2 * h3 + h3)        # wgrad of all layers + dgrad without layer 1
    assert 2 * (actor_fwd + 4 * critic_fwd + 2 * critic_wgrad_dgrad) == bench.FLOP_V
    assert bench.MAC_ACTOR_FWD == actor_fwd and bench.MAC_CRITIC_FWD == critic_fwd
    critic_dgrad_all = critic_fwd                                       # dgrad through every layer of the frozen critic
    actor_bwd = actor_fwd + (h1 * h2 + h2 * h3 + h3 * A)
    assert 2 * (actor_fwd + 2 * critic_fwd + 2 * critic_dgrad_all + actor_bwd) == bench.FLOP_P
    assert bench.BYTES_INSERT == 4 * (2 * O + A + 2) + 4 * (2 * O + A + 1) + 1
    assert bench.BYTES_SAMPLE == 8 + (4 * (2 * O + A + 1) + 1) + 4 * (2 * O + A + 2)


@pytest.mark.parametrize("shapes,B", [([(128, 256), (256, 512), (512, 104)] * 2, 8192),          # twin-Q critic, AllegroHand
                                      ([(16, 128), (128, 256), (256, 512), (512, 88)], 8192),     # actor
                                      ([(51, 128), (128, 256), (256, 512), (512, 104)] * 2, 16384),   # C51 critic
                                      ([(128, 256), (256, 512), (512, 231)] * 2, 8192),           # ShadowHand critic
                                      ([(128, 256), (256, 512), (512, 104)] * 2, 96)])            # tiny batch: few k-blocks
def test_wgrad_plan_fills_one_wave_with_balanced_ctas(shapes, B):
    """pqlb_wgrad_multi's split plan: at most one CTA per SM, every split owns >= 2 k-blocks, and the CTAs stream
    about the same number of operand bytes (these GEMMs are bound by L2 -> SM ingest)."""
    plan = K.wgrad_plan(shapes, B, target_ctas=148)
    kb = (B + 31) // 32
    assert len(plan) == len(shapes)
    ctas, per_cta = 0, []
    for (M, N), (tile_n, splits) in zip(shapes, plan):
        assert tile_n == K.wgrad_tile_n(N) and tile_n in (32, 64, 128, 256)
        assert 1 <= splits <= max(1, kb // 2)
        tiles = -(-M // 128) * -(-N // tile_n)
        ctas += tiles * splits
        per_cta.append(4096.0 * (-(-min(M, 128) // 32) + -(-min(N, tile_n) // 32)) * kb / splits)
    assert ctas <= 148
    if kb >= 64:                     # enough k-blocks to balance: no CTA streams more than 1.6x the lightest full-size one
        assert ctas >= 0.9 * 148
        full = [c for c, (M, N) in zip(per_cta, shapes) if M >= 128]
        assert max(full) <= 1.6 * min(full)


def test_forward_mode_selection():
    """Which shapes take the split-fp16 fused forward: inputs up to 128 wide with 16-byte
    aligned half rows and a policy head of at most 16 actions; critics of 129..256 columns (the wide-input kernel),
    whose policy net then decides for itself (split_f16_ok); everything else runs the TF32 kernels."""
    from pql_b200.algo._engine import forward_mode
    from pql_b200.models.mlp import NetLayout, split_f16_ok
    assert forward_mode(None, 88, 16) == "f16x3"
    assert forward_mode("tf32", 88, 16) == "tf32"
    assert forward_mode(None, 211, 20) == "f16x3"         # ShadowHand: 231 input columns, critics on the wide-input kernel
    assert forward_mode("tf32", 211, 20) == "tf32"
    assert forward_mode(None, 300, 20) == "tf32"          # more than 256 columns: per-layer TF32 launches
    assert forward_mode(None, 24, 4) in ("f16x3", "tf32")
    with pytest.raises(ValueError):
        forward_mode("bf16", 88, 16)

    class Net:                                             # what split_f16_ok / fused_head_ok read of a NetAddrs
        def __init__(self, in_dim, out_dim):
            L = NetLayout(in_dim, out_dim)
            self.dims, self.ldw, self.Wh = L.dims, L.ldw, [1]
    from pql_b200.models.mlp import fused_head_ok
    # ShadowHand: critics (231 -> rows of 232 halves) and the policy net (211 -> 216) take the fused kernel; behind a wide
    # input the head takes up to 32 actions and unaligned output rows (211 * 4 bytes), behind a narrow one 16 / aligned
    assert NetLayout(231, 1).ldw[0] == 232 and NetLayout(211, 20).ldw[0] == 216 and NetLayout(88, 16).ldw[0] == 88
    assert NetLayout(12, 4).ldw[0] == 12                   # narrow first layers keep 16-byte fp32 rows
    assert split_f16_ok(dict(net=Net(231, 1), k_in=231))
    shadow_policy = dict(net=Net(211, 20), k_in=211, act=dict(out=211 * 4, ldo=232, out2=211 * 4, ldo2=232))
    assert split_f16_ok(shadow_policy) and fused_head_ok(shadow_policy)
    assert not fused_head_ok(dict(net=Net(211, 36), k_in=211, act=dict(out=211 * 4, ldo=248)))     # more than 32 actions
    assert not fused_head_ok(dict(net=Net(88, 20), k_in=88, act=dict(out=88 * 4, ldo=108)))        # 20 actions, narrow input
    wide16 = dict(net=Net(208, 16), k_in=208, act=dict(out=208 * 4, ldo=224, out2=208 * 4, ldo2=224))
    assert split_f16_ok(wide16) and fused_head_ok(wide16)
    assert not fused_head_ok(dict(net=Net(88, 16), k_in=88, act=dict(out=88 * 4 + 4, ldo=104)))  # unaligned action rows
    assert not split_f16_ok(dict(net=Net(260, 1), k_in=260))
    assert not split_f16_ok(dict(net=Net(12, 4), k_in=12))  # 12 halves = 24-byte rows


def test_forward_launch_plans():
    """Which launches a forward becomes (models/mlp.py: forward_calls), built on fake addresses - no GPU needed:
    ShadowHand's critics, policy net and 20-action head are ONE fused split-fp16 launch each (no per-layer GEMM);
    a head the kernel does not fuse becomes one extra launch behind a trunk that stores h3; nets without fp16
    copies and inputs wider than 128 run layer by layer."""
    import torch
    from pql_b200 import _kernels as K
    from pql_b200.models.mlp import NetAddrs, NetLayout, forward_calls
    B, base = 8192, 1 << 20

    def net(in_dim, out_dim, halves=True):
        L = NetLayout(in_dim, out_dim)
        half = torch.zeros(2 * L.total, dtype=torch.float16) if halves else None
        return NetAddrs(L, 0, base, base + 4 * L.total, half), half

    def inst(n, k_in, x_ld, **kw):
        return dict(net=n, x=base, xf=base + 64, x_ld=x_ld, k_in=k_in, h=[base, base, base], **kw)

    keep = []
    # ShadowHand critics: four networks, scalar heads, one launch on the wide-input kernel
    crit = []
    for i in range(4):
        n, h = net(231, 1); keep.append(h)
        crit.append(inst(n, 231, 232, q=base + 128, terms=3, store=(i < 2,) * 3))
    calls = forward_calls(B, crit, True)
    assert len(calls) == 1 and isinstance(calls[0], K.MlpForwardH)
    d = calls[0].desc
    assert (d.M, d.k_in, d.n_groups) == (B, 231, 4) and d.g[0].ldw1 == 232 and d.g[0].terms == 3
    assert bool(d.g[0].h1) and not bool(d.g[2].h1) and bool(d.g[3].q)
    # ShadowHand policy: 211 observations (rows of 216 halves), 20 actions written at column 211 of the critics' rows
    n, h = net(211, 20); keep.append(h)
    act = dict(out=base + 4 * 211, ldo=232, out2=base + 64 + 4 * 211, ldo2=232, noise=base + 256, ldnoise=20, noise_std=0.8, noise_bound=0.2)
    calls = forward_calls(B, [inst(n, 211, 232, act=act, terms=1, store=(False,) * 3)], False)
    assert len(calls) == 1 and isinstance(calls[0], K.MlpForwardH)
    g = calls[0].desc.g[0]
    assert g.act_n == 20 and g.ldw1 == 216 and g.terms == 1 and not bool(g.h3) and g.act_ldo == 232
    # 36 actions: the trunk is still one fused launch (it stores h3), the head is its own GEMM
    n, h = net(211, 36); keep.append(h)
    calls = forward_calls(B, [inst(n, 211, 248, act=dict(out=base + 4 * 211, ldo=248), terms=1, store=(False,) * 3)], False)
    assert [type(c) for c in calls] == [K.MlpForwardH, K.Gemm]
    assert bool(calls[0].desc.g[0].h3) and calls[0].desc.g[0].act_n == 0
    # no fp16 copies (a module's own forward): 211 inputs run layer by layer - three trunk GEMMs + the head
    n, _ = net(211, 20, halves=False)
    calls = forward_calls(B, [inst(n, 211, 216, act=dict(out=base, ldo=20))], False)
    assert [type(c) for c in calls] == [K.Gemm] * 4
    # AllegroHand policy without fp16 copies: the TF32 layer-fused kernel
    n, _ = net(88, 16, halves=False)
    calls = forward_calls(B, [inst(n, 88, 88, act=dict(out=base, ldo=16))], False)
    assert len(calls) == 1 and isinstance(calls[0], K.MlpForward)
