"""GPU tier: PQLVLearner.learn / PQLPLearner.learn (twin-Q, C51, ShadowHand shape) against the
torch-CPU oracle on identical weights, batches, indices and noise; module-level forward API;
RNG stream parity of the ``out=`` draws."""
import os

import numpy as np
import pytest
import torch

from oracle import learner as L
from tests import parity
from tests.golden import inputs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("tag,seed,B,O,A,distl,steps", [("doubleq", 1234, 512, 88, 16, False, 3),
                                                        ("c51", 4321, 256, 88, 16, True, 3),
                                                        ("shadow", 77, 128, 211, 20, False, 2),
                                                        ("ragged", 5, 200, 88, 16, False, 1)])
def test_learner_updates_match_oracle(tag, seed, B, O, A, distl, steps):
    res = parity.run_learner_parity(seed=seed, B=B, obs_dim=O, act_dim=A, distl=distl, steps=steps, device=DEV)
    print(tag, res)


def test_full_batch_c51():
    """BASELINE config 3 batch size (16384), 51 atoms."""
    res = parity.run_learner_parity(seed=12, B=16384, obs_dim=88, act_dim=16, distl=True, steps=1, device=DEV)
    print(res)


def test_learner_without_obs_norm():
    parity.run_learner_parity(seed=3, B=256, obs_dim=88, act_dim=16, distl=False, steps=1, device=DEV, obs_norm=False)


def test_full_batch_allegro_doubleq():
    """BASELINE config 1/2 batch size (8192): same tolerances (tests/parity.py) at the size the
    benchmark runs."""
    res = parity.run_learner_parity(seed=11, B=8192, obs_dim=88, act_dim=16, distl=False, steps=1, device=DEV)
    print(res)


def test_full_batch_shadowhand_doubleq():
    """BASELINE config 4's shapes (obs 211, act 20) at the benchmark's batch size: critics, policy net and heads on the
    wide-input split-fp16 kernel (csrc/mlp_fwd_h.cu, mlp_fwd_hw_kernel); same flat 1e-3 as every other case."""
    res = parity.run_learner_parity(seed=21, B=8192, obs_dim=211, act_dim=20, distl=False, steps=1, device=DEV)
    print(res)


def test_free_running_trace_vs_reference_fixture(golden_dir):
    """Three unsynchronised updates from the reference's initial weights: losses recorded from
    the reference run itself (tests/golden/learner_doubleq.npz).  TF32 rounding compounds
    through Adam's normalised step, so the trace tolerance is 1e-2 (step 1: 1e-3)."""
    from pql_b200.algo import PQLPLearner, PQLVLearner
    from pql_b200.models import TanhMLPPolicy
    g = np.load(os.path.join(golden_dir, "learner_doubleq.npz"))
    B, O, A = 512, 88, 16
    case = inputs.learner_case(1234, B, O, A, False)
    cfg = parity.make_cfg(B, False)
    dev = torch.device(DEV)
    v = PQLVLearner(O, A, cfg)
    parity.load_params(v.critic.net_q1, case["q1"]); parity.load_params(v.critic.net_q2, case["q2"])
    actor = TanhMLPPolicy(O, A).to(dev)
    parity.load_params(actor, case["actor"])
    norm = (case["norm"][0].to(dev), case["norm"][1].to(dev), case["norm"][2])
    v.update(actor, tuple(x.to(dev) for x in case["batch"]), norm, 0)
    idx = torch.from_numpy(g["idx"])
    losses = []
    for s in range(3):
        with parity.injected_draws(idx, case["noises"][s]):
            v.learn()
        losses.append(v._plan.loss.item())
    assert losses[0] == pytest.approx(g["v_losses"][0], rel=1e-3)
    np.testing.assert_allclose(losses, g["v_losses"], rtol=1e-2)
    assert v.update_count == 3
    crit, mean, cnt = v.update(actor, tuple(x[:7].to(dev) for x in case["batch"]), norm, 0)
    assert crit is v.critic and cnt == 3
    # Tracker(5) pre-filled with zeros (common.py:104-105): mean over [0, 0, l0, l1, l2]
    assert mean == pytest.approx(sum(losses) / 5, rel=1e-5)
    p = PQLPLearner(O, A, cfg)
    parity.load_params(p.actor, case["actor"])
    crit0 = type(v.critic)(O, A).to(dev)
    parity.load_params(crit0.net_q1, case["q1"]); parity.load_params(crit0.net_q2, case["q2"])
    p.update(crit0, case["batch"][0].to(dev), norm, 0)
    pl = []
    for s in range(3):
        with parity.injected_draws(idx):
            p.learn()
        pl.append(p._plan.loss.item())
    assert pl[0] == pytest.approx(g["p_losses"][0], rel=1e-3, abs=1e-6)
    np.testing.assert_allclose(pl, g["p_losses"], rtol=2e-2, atol=1e-5)


def test_module_forward_api_matches_oracle():
    from pql_b200.models import DistributionalDoubleQ, DoubleQ, TanhMLPPolicy
    B, O, A = 300, 88, 16
    dev = torch.device(DEV)
    for distl in (False, True):
        case = inputs.learner_case(9, B, O, A, distl)
        obs, act = case["batch"][0], case["batch"][1]
        crit = (DistributionalDoubleQ(O, A, device=dev) if distl else DoubleQ(O, A)).to(dev)
        parity.load_params(crit.net_q1, case["q1"]); parity.load_params(crit.net_q2, case["q2"])
        q1, q2 = crit.get_q1_q2(obs.to(dev), act.to(dev))
        r1, r2 = L.q1_q2(obs, act, case["q1"], case["q2"], distl)
        assert q1.shape == r1.shape
        assert parity.rel(q1, r1) <= 1e-3 and parity.rel(q2, r2) <= 1e-3
        assert parity.rel(crit.get_q1(obs.to(dev), act.to(dev)), r1) <= 1e-3
        qm = crit.get_q_min(obs.to(dev), act.to(dev))
        rm = L.q_min(obs, act, case["q1"], case["q2"], distl, torch.linspace(-10, 10, 51) if distl else None)
        assert qm.shape == rm.shape and parity.rel(qm, rm) <= 1e-3
    pol = TanhMLPPolicy(O, A).to(dev)
    parity.load_params(pol, case["actor"])
    a = pol(obs.to(dev))
    assert a.shape == (B, A) and parity.rel(a, L.actor_forward(obs, case["actor"])) <= 1e-3
    # deepcopy / state_dict round trip keeps the kernels and the parameters in sync
    import copy
    pol2 = copy.deepcopy(pol)
    assert torch.equal(pol2(obs.to(dev)), a)
    pol3 = TanhMLPPolicy(O, A).to(dev)
    pol3.load_state_dict(pol.state_dict())
    assert torch.equal(pol3(obs.to(dev)), a)


def test_rng_out_variants_follow_the_reference_stream():
    """learn() draws with out= buffers; the values must equal the reference's functional calls
    (simple_replay.py:87, noise.py:20-21) for the same seed."""
    dev = torch.device(DEV)
    torch.manual_seed(42)
    idx_ref = torch.randint(12345, size=(8192,), device=dev)
    noise_ref = torch.normal(torch.zeros(8192, 16, device=dev), torch.full((8192, 16), 0.8, device=dev))
    torch.manual_seed(42)
    idx = torch.zeros(8192, dtype=torch.int64, device=dev)
    noise = torch.zeros(8192, 16, device=dev)
    torch.randint(12345, size=(8192,), device=dev, out=idx)
    noise.normal_()          # the kernel applies std: normal(zeros, full(std)) == normal_(0,1).mul_(std).add_(0)
    assert torch.equal(idx, idx_ref) and torch.equal(noise * 0.8, noise_ref)


@pytest.mark.parametrize("B,distl,which", [(8192, False, "critic"), (8192, False, "actor"),
                                           (16384, True, "critic"), (16384, True, "actor")])
def test_update_launch_list_is_bit_reproducible(B, distl, which):
    """Every launch of the update must reproduce its outputs bit for bit (fixed-order
    reductions, no atomics).  Regression test for the dgrad-epilogue race found in round 1: the
    refill of the ELU'-operand staging buffer could overtake the shared-memory reads of the
    previous chunk when two CTAs shared an SM."""
    assert parity.plan_divergence(B, distl, which, repeats=4, device=DEV) is None


def _run_interleaved(streams, seed=9, B=512, steps=3, fused_rng=True, graph=True, sync_loss=True):
    """bench.py's loop in miniature: per env step one update() exchange, 4 critic + 2 actor updates."""
    from pql_b200.algo import PQLPLearner, PQLVLearner
    from pql_b200.replay import NStepReplay
    from pql_b200.utils import default_pql_cfg
    O, A, E = 88, 16, 256
    torch.manual_seed(seed)
    cfg = default_pql_cfg(batch_size=B, memory_size=4096, num_envs=E)
    cfg.learner_streams = streams
    cfg.fused_rng = fused_rng
    cfg.use_cuda_graph = graph
    cfg.sync_loss = sync_loss
    v, p = PQLVLearner(O, A, cfg), PQLPLearner(O, A, cfg)
    ns = NStepReplay(O, A, num_envs=E, nstep=3, device=DEV)
    g = torch.Generator(device=DEV).manual_seed(seed)
    norm = (torch.zeros(O, device=DEV), torch.ones(O, device=DEV), 1e-4)
    critic, actor = v.start()[0], p.start()[0]
    losses = []
    for k in range(steps):
        T = 8 if k == 0 else 1
        blk = (torch.randn(E, T, O, device=DEV, generator=g), torch.rand(E, T, A, device=DEV, generator=g) * 2 - 1,
               torch.randn(E, T, 1, device=DEV, generator=g) * 0.01, torch.randn(E, T, O, device=DEV, generator=g),
               (torch.rand(E, T, 1, device=DEV, generator=g) < 0.05).float())
        tr = ns.add_to_buffer(*blk)
        critic, vl, _ = v.update(actor, tr, norm, 0)
        actor, pl, _ = p.update(critic, tr[0], norm, 0)
        losses.append((vl, pl))
        for j in range(4):
            v.learn()
            if j % 2 == 1:
                p.learn()
    torch.cuda.synchronize()
    return v.critic.arena.flat.clone(), p.actor.arena.flat.clone(), losses


def test_learner_streams_match_single_stream():
    """cfg.learner_streams only changes WHERE the launches are enqueued: with the update() exchange
    as the join point, weights and reported losses must be bit-identical to the one-stream run."""
    c0, a0, l0 = _run_interleaved(False)
    c1, a1, l1 = _run_interleaved(True)
    assert torch.equal(c0, c1) and torch.equal(a0, a1)
    assert l0 == l1
    assert torch.isfinite(c0).all() and torch.isfinite(a0).all()


def test_fused_rng_update_equals_torch_draws():
    """cfg.fused_rng moves torch.randint / normal_ into the gather kernel (csrc/rng.cuh): same
    generator state => same indices and noise => bit-identical weights and losses, with and without
    CUDA-graph replay (the offsets advance on the device)."""
    c0, a0, l0 = _run_interleaved(False, fused_rng=False)
    c1, a1, l1 = _run_interleaved(False, fused_rng=True)
    c2, a2, l2 = _run_interleaved(True, fused_rng=True, graph=False)
    assert torch.equal(c0, c1) and torch.equal(a0, a1) and l0 == l1
    assert torch.equal(c0, c2) and torch.equal(a0, a2) and l0 == l2


def test_prefetching_sampler_is_the_path_that_runs_and_changes_nothing():
    """With CUDA graphs and the fused sampler RNG (the defaults) the V-learner draws batch k + 1 and runs its
    target-policy forward in a side branch of update k's graph (_engine.CriticUpdate._run_prefetching).  The graphs
    of both kinds (first update after an exchange / update fed by the look-ahead) must have been built for both input
    sets, the last update's draws must be readable under the usual names, and weights and losses must equal the
    run in which every update draws its own batch first (graphs off)."""
    from pql_b200.algo import PQLVLearner
    from pql_b200.utils import default_pql_cfg
    c0, a0, l0 = _run_interleaved(False, fused_rng=True, graph=False)
    c1, a1, l1 = _run_interleaved(True, fused_rng=True, graph=True)
    assert torch.equal(c0, c1) and torch.equal(a0, a1) and l0 == l1
    cfg = default_pql_cfg(batch_size=256, memory_size=2048, num_envs=64)
    v = PQLVLearner(24, 4, cfg)
    from pql_b200.models import TanhMLPPolicy
    actor = TanhMLPPolicy(24, 4).to(DEV)
    g = torch.Generator(device=DEV).manual_seed(0)
    tr = (torch.randn(512, 24, device=DEV, generator=g), torch.rand(512, 4, device=DEV, generator=g), torch.randn(512, 1, device=DEV, generator=g),
          torch.randn(512, 24, device=DEV, generator=g), torch.zeros(512, 1, device=DEV))
    norm = (torch.zeros(24, device=DEV), torch.ones(24, device=DEV), 1e-4)
    seen = []
    for _ in range(2):
        v.update(actor, tr, norm, 0)
        for _ in range(3):
            v.learn()
            seen.append((v._plan.idx.clone(), v.memory.cur_capacity))
    torch.cuda.synchronize()
    plan = v._plan
    assert plan.n_sets == 2 and sorted(plan._pf_graphs) == [(False, 0), (False, 1), (True, 0), (True, 1)]
    # the draws are torch.randint's for the learner's generator, one call per update, in order
    gen = torch.Generator(device=DEV)
    gen.manual_seed(v.generator.initial_seed())
    assert [cap for _, cap in seen] == [512] * 3 + [1024] * 3
    for got, cap in seen:
        want = torch.randint(cap, (256,), device=DEV, generator=gen)
        torch.empty(256, 4, device=DEV).normal_(generator=gen)
        assert torch.equal(got, want)


def test_lagged_loss_readback_is_the_previous_mean():
    """Default (cfg.sync_loss off): update() returns the loss mean as of the previous update() without
    waiting for the GPU; weights are unaffected."""
    c0, a0, l0 = _run_interleaved(True, sync_loss=True)
    c1, a1, l1 = _run_interleaved(True, sync_loss=False)
    assert torch.equal(c0, c1) and torch.equal(a0, a1)
    assert l1[0] == (0.0, 0.0)
    assert l1[1:] == l0[:-1]
