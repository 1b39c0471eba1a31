"""GPU tier: the fused sampler RNG.  pqlb_sample_critic_batch_rng / pqlb_sample_obs_batch_rng
draw torch.randint(cur_capacity, (B,)) and the N(0,1) values behind torch.normal(zeros, full(std))
inside the gather kernel; for the same generator (seed, offset) the values must be the ones ATen
returns (simple_replay.py:87, pql_p_learner.py:49, noise.py:20-21) - bit for bit."""
import pytest
import torch

from pql_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ring(cap, O, A, gen):
    from pql_b200.replay import ReplayBuffer
    mem = ReplayBuffer(cap, O, A, device=DEV)
    rows = (torch.randn(cap, O, device=DEV, generator=gen), torch.rand(cap, A, device=DEV, generator=gen),
            torch.randn(cap, 1, device=DEV, generator=gen), torch.randn(cap, O, device=DEV, generator=gen),
            (torch.rand(cap, 1, device=DEV, generator=gen) < 0.1).float())
    mem.add_to_buffer(rows)
    return mem


@pytest.mark.parametrize("B,O,A,cap,fill,seed,offset,count", [
    (8192, 88, 16, 50_000, 50_000, 42, 0, 0),
    (16384, 88, 16, 100_000, 70_001, 7, 4096, 3),        # partly filled ring, later offset, count > 0
    (300, 211, 20, 4_000, 1_234, 2 ** 40 + 17, 8, 11),   # ShadowHand shape (scalar-item kernel), ragged batch
    (8192 * 4, 88, 16, 50_000, 50_000, 1, 0, 0),         # noise numel 524288 > one grid pass of ATen: components .y
])
def test_critic_batch_draws_equal_torch(B, O, A, cap, fill, seed, offset, count):
    g = torch.Generator(device=DEV).manual_seed(123)
    mem = _ring(cap, O, A, g)
    cur = torch.tensor([fill], dtype=torch.int64, device=DEV)
    x_ld = (O + A + 3) // 4 * 4
    inc = 8
    state = torch.tensor([seed, offset - inc * count, inc], dtype=torch.int64, device=DEV)
    counter = torch.tensor([count], dtype=torch.int64, device=DEV)
    idx = torch.full((B,), -1, dtype=torch.int64, device=DEV)
    noise = torch.zeros(B, A, device=DEV)
    mean, var = torch.randn(O, device=DEV, generator=g), torch.rand(O, device=DEV, generator=g) + 0.5
    outs = [[torch.zeros(B, x_ld, device=DEV), torch.zeros(B, x_ld, device=DEV), torch.zeros(B, device=DEV),
             torch.zeros(B, device=DEV)] for _ in range(2)]
    xf = [[torch.zeros(B, x_ld, device=DEV), torch.zeros(B, x_ld, device=DEV)] for _ in range(2)]     # un-rounded twins
    _lib.call("pqlb_sample_critic_batch_rng", _lib.ptr(mem.ring), cap, O, A, _lib.ptr(idx), B, _lib.ptr(mean), _lib.ptr(var),
              1e-4, *(_lib.ptr(t) for t in outs[0][:2]), x_ld, *(_lib.ptr(t) for t in outs[0][2:]), _lib.ptr(state),
              _lib.ptr(counter), _lib.ptr(cur), _lib.ptr(noise), noise.numel(), _lib.ptr(xf[0][0]), _lib.ptr(xf[0][1]))
    ref = torch.Generator(device=DEV).manual_seed(seed)
    ref.set_offset(offset)
    idx_ref = torch.randint(fill, size=(B,), device=DEV, generator=ref)
    noise_ref = torch.zeros(B, A, device=DEV).normal_(generator=ref)
    assert torch.equal(idx, idx_ref)
    assert torch.equal(noise, noise_ref)
    assert ref.get_offset() == offset + inc
    # and the gather itself equals the unfused entry point on those indices
    _lib.call("pqlb_sample_critic_batch", _lib.ptr(mem.ring), cap, O, A, _lib.ptr(idx_ref), B, _lib.ptr(mean), _lib.ptr(var),
              1e-4, *(_lib.ptr(t) for t in outs[1][:2]), x_ld, *(_lib.ptr(t) for t in outs[1][2:]),
              _lib.ptr(xf[1][0]), _lib.ptr(xf[1][1]))
    for a, b in zip(*outs):
        assert torch.equal(a[:, :O + A] if a.dim() == 2 else a, b[:, :O + A] if b.dim() == 2 else b)
    assert torch.equal(outs[0][0], outs[1][0])
    # the un-rounded rows: identical between the two entry points, and their TF32 rounding is the rounded row
    assert torch.equal(xf[0][0], xf[1][0]) and torch.equal(xf[0][1][:, :O], xf[1][1][:, :O])
    rn = lambda t: ((t.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)      # noqa: E731
    assert torch.equal(rn(xf[0][0]), outs[0][0]) and torch.equal(rn(xf[0][1][:, :O]), outs[0][1][:, :O])


@pytest.mark.parametrize("B,O,A,cap,fill,seed,offset", [(8192, 88, 16, 60_000, 60_000, 42, 0),
                                                        (1000, 211, 20, 5_000, 777, 9, 400)])
def test_obs_batch_draws_equal_torch(B, O, A, cap, fill, seed, offset):
    g = torch.Generator(device=DEV).manual_seed(5)
    ring = torch.randn(cap, O, device=DEV, generator=g)
    cur = torch.tensor([fill], dtype=torch.int64, device=DEV)
    x_ld = (O + A + 3) // 4 * 4
    state = torch.tensor([seed, offset, 4], dtype=torch.int64, device=DEV)
    counter = torch.zeros(1, dtype=torch.int64, device=DEV)
    idx = torch.full((B,), -1, dtype=torch.int64, device=DEV)
    mean, var = torch.randn(O, device=DEV, generator=g), torch.rand(O, device=DEV, generator=g) + 0.5
    x0, x1 = torch.full((B, x_ld), 7.0, device=DEV), torch.full((B, x_ld), 7.0, device=DEV)
    _lib.call("pqlb_sample_obs_batch_rng", _lib.ptr(ring), cap, O, _lib.ptr(idx), B, _lib.ptr(mean), _lib.ptr(var), 1e-4,
              _lib.ptr(x0), x_ld, A, _lib.ptr(state), _lib.ptr(counter), _lib.ptr(cur), None)
    ref = torch.Generator(device=DEV).manual_seed(seed)
    ref.set_offset(offset)
    idx_ref = torch.randint(fill, size=(B,), device=DEV, generator=ref)
    assert torch.equal(idx, idx_ref)
    _lib.call("pqlb_sample_obs_batch", _lib.ptr(ring), cap, O, _lib.ptr(idx_ref), B, _lib.ptr(mean), _lib.ptr(var), 1e-4,
              _lib.ptr(x1), x_ld, A, None)
    assert torch.equal(x0, x1)          # action columns untouched (7.0), padding zeroed, obs normalised + TF32-rounded


def test_learner_generator_stream():
    """A learner with cfg.fused_rng consumes its generator's stream exactly like the torch calls:
    after k updates plan.idx / plan.noise hold draw k of a twin generator."""
    from pql_b200.algo import PQLVLearner
    from pql_b200.models import TanhMLPPolicy
    from pql_b200.utils import default_pql_cfg
    B, O, A, E = 1024, 88, 16, 512
    torch.manual_seed(77)
    cfg = default_pql_cfg(batch_size=B, memory_size=3000, num_envs=E)
    v = PQLVLearner(O, A, cfg)
    assert v.fused_rng
    twin = torch.Generator(device=DEV).manual_seed(v.generator.initial_seed())
    g = torch.Generator(device=DEV).manual_seed(1)
    actor = TanhMLPPolicy(O, A).to(DEV)
    norm = (torch.zeros(O, device=DEV), torch.ones(O, device=DEV), 1e-4)
    for step in range(3):       # the ring fills up while we go: the range of the draw follows cur_capacity
        n = 1400
        tr = (torch.randn(n, O, device=DEV, generator=g), torch.rand(n, A, device=DEV, generator=g),
              torch.randn(n, 1, device=DEV, generator=g), torch.randn(n, O, device=DEV, generator=g),
              torch.zeros(n, 1, device=DEV))
        v.update(actor, tr, norm, 0)
        for _ in range(2):
            v.learn()
            idx_ref = torch.randint(v.memory.cur_capacity, size=(B,), device=DEV, generator=twin)
            noise_ref = torch.zeros(B, A, device=DEV).normal_(generator=twin)
            assert torch.equal(v._plan.idx, idx_ref) and torch.equal(v._plan.noise, noise_ref)
    assert v.memory.if_full and v.update_count == 6
