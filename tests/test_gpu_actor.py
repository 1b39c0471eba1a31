"""GPU tier: the actor-side env-step path - RunningMeanStd.update, the exploration-noise
draw, trackers / timeout handling / reward scaling, and PQLActor.explore_env end to end - against
oracle/actor.py (itself pinned to the reference's PQLActor, tests/test_oracle_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import actor as OA
from pql_b200 import _lib
from tests import parity
from tests.golden import inputs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("rows,cols", [(4096, 88), (130, 211), (16384, 211), (2, 5)])
def test_running_mean_std_matches_oracle(rows, cols):
    """torch_util.py:77-103.  Tolerance: 2e-6 relative (+1e-6 of the column scale): the batch moments
    are accumulated in fp64 here and in fp32 by torch, the merge is the same fp32 arithmetic."""
    from pql_b200.utils import RunningMeanStd
    g = torch.Generator().manual_seed(rows + cols)
    ref = OA.RunningMeanStdOracle(shape=(cols,))
    rms = RunningMeanStd(shape=(cols,), device=DEV)
    scale = torch.rand(cols, generator=g) * 3 + 0.1
    shift = torch.randn(cols, generator=g) * 2
    for k in range(4):
        x = torch.randn(rows, cols, generator=g) * scale + shift * (1 + 0.1 * k)
        ref.update(x)
        rms.update(x.to(DEV))
        assert rms.count == ref.count
        assert float(rms._count_dev.item()) == ref.count
        np.testing.assert_allclose(rms.mean.cpu().numpy(), ref.mean.numpy(), rtol=2e-6, atol=1e-6)
        np.testing.assert_allclose(rms.var.cpu().numpy(), ref.var.numpy(), rtol=4e-6, atol=1e-6)
    x = torch.randn(7, cols, generator=g)
    np.testing.assert_allclose(rms.normalize(x.to(DEV)).cpu().numpy(), ref.normalize(x).numpy(), rtol=1e-5, atol=1e-6)
    m, v, eps = rms.get_states()
    # snapshots (update() writes mean / var in place; the reference rebinds fresh tensors): same values, other storage
    assert torch.equal(m, rms.mean) and torch.equal(v, rms.var) and eps == 1e-4
    assert m.data_ptr() != rms.mean.data_ptr() and v.data_ptr() != rms.var.data_ptr()


@pytest.mark.parametrize("E,A,kind", [(4096, 16, "mixed"), (16384, 20, "mixed"), (300, 3, "fixed")])
def test_exploration_noise_draw_equals_torch(E, A, kind):
    """noise.py:19-41: torch.normal(zeros, std) with the actor's generator state - bit for bit."""
    seed, offset = 1234, 40
    noise = torch.zeros(E, A, device=DEV)
    row_std = torch.linspace(0.05, 0.8, E).to(DEV)
    _lib.call("pqlb_actor_inputs", None, E, 0, 0, None, None, 0.0, 0, 0, None, 0, _lib.ptr(noise), A,
              _lib.ptr(row_std) if kind == "mixed" else None, 0.3, seed, offset)
    gen = torch.Generator(device=DEV).manual_seed(seed)
    gen.set_offset(offset)
    std = row_std.unsqueeze(-1).expand(E, A) if kind == "mixed" else torch.full((E, A), 0.3, device=DEV)
    ref = torch.normal(torch.zeros(E, A, device=DEV), std, generator=gen)
    assert torch.equal(noise, ref)
    assert gen.get_offset() == offset + 4


def test_env_post_matches_tracker_semantics():
    """pql_actor.py:129-136 + common.py:195-202 + :117, including a step that finishes more episodes
    than the Tracker window holds (only the last max_len survive, in env order)."""
    E, L = 3000, 7
    g = torch.Generator().manual_seed(5)
    ret_t, len_t = OA.TrackerOracle(L), OA.TrackerOracle(L)
    cur_r, cur_l = torch.zeros(E), torch.zeros(E)
    d = dict(returns=torch.zeros(E, device=DEV), lengths=torch.zeros(E, device=DEV), rw=torch.zeros(L, device=DEV),
             lw=torch.zeros(L, device=DEV), pushed=torch.zeros(1, dtype=torch.int64, device=DEV))
    for step, p_done in enumerate([0.0, 0.001, 0.5, 0.0005, 0.01, 1.0, 0.002]):
        reward = torch.randn(E, generator=g)
        done = (torch.rand(E, generator=g) < p_done).float()
        trunc = (torch.rand(E, generator=g) < 0.5) & (done != 0)
        cur_r += reward; cur_l += 1
        idx = torch.where(done)[0]
        ret_t.update(cur_r[idx]); len_t.update(cur_l[idx])
        cur_r[idx] = 0; cur_l[idx] = 0
        r_out, d_out = torch.empty(E, device=DEV), torch.empty(E, device=DEV)
        rd, dd, td = reward.to(DEV), done.to(DEV), trunc.to(DEV)      # kept alive: ptr() of a temporary would dangle
        _lib.call("pqlb_env_post", _lib.ptr(rd), _lib.ptr(dd), _lib.ptr(td), 0.01, E,
                  _lib.ptr(d["returns"]), _lib.ptr(d["lengths"]), _lib.ptr(d["rw"]), _lib.ptr(d["lw"]), L,
                  _lib.ptr(d["pushed"]), _lib.ptr(r_out), _lib.ptr(d_out))
        assert torch.equal(d["returns"].cpu(), cur_r) and torch.equal(d["lengths"].cpu(), cur_l)
        assert torch.equal(r_out.cpu(), 0.01 * reward)
        assert torch.equal(d_out.cpu(), OA.handle_timeout(done, {"TimeLimit.truncated": trunc}))
        # the device ring holds the deque's contents (rotated by pushed % L)
        n = int(d["pushed"].item())
        ring_r, ring_l = d["rw"].cpu().tolist(), d["lw"].cpu().tolist()
        as_deque = lambda ring: [ring[(n + k) % L] for k in range(L)]           # noqa: E731
        assert as_deque(ring_r) == [float(np.float32(x)) for x in ret_t.moving_average]
        assert as_deque(ring_l) == [float(x) for x in len_t.moving_average]


def _twin_draws(seed, E, A, warm, steps, kind, std_max=0.8, std_min=0.05):
    """What a torch.Generator seeded like the actor's returns for the same sequence of calls."""
    gen = torch.Generator(device=DEV).manual_seed(seed)
    draws = [torch.rand((E, A), device=DEV, generator=gen).cpu() for _ in range(warm)]
    std = (torch.linspace(std_min, std_max, E).to(DEV).unsqueeze(-1).expand(E, A) if kind == "mixed"
           else torch.full((E, A), std_max, device=DEV))
    draws += [torch.normal(torch.zeros(E, A, device=DEV), std, generator=gen).cpu() for _ in range(steps)]
    return draws


@pytest.mark.parametrize("E,O,A,kind,calls", [(512, 88, 16, "mixed", [1, 1, 1, 3, 1]),
                                              (256, 211, 20, "mixed", [1, 2, 1]),
                                              (128, 88, 16, "fixed", [1, 1, 1])])
def test_explore_env_matches_oracle(E, O, A, kind, calls):
    """PQLActor.explore_env (warm-up with random actions, then policy steps) on the scripted env:
    observations / next observations / dones / trackers' lengths exact, obs_rms to 4e-6, actions,
    rewards and returns to 1e-3 norm-wise (TF32 policy forward against the fp32 reference arithmetic)."""
    from pql_b200.algo import PQLActor
    from pql_b200.models import TanhMLPPolicy
    from pql_b200.utils import default_pql_cfg
    warm, seed = 5, 3
    T = warm + sum(calls)
    torch.manual_seed(99)
    cfg = default_pql_cfg(num_envs=E, sim_device=DEV, tracker_len=11)
    cfg.algo.noise.type = kind
    env = inputs.ScriptedEnv(seed, E, O, A, T, device=DEV)
    act = PQLActor(env, cfg)
    params = inputs.actor_case_params(seed, O, A)
    pol = TanhMLPPolicy(O, A).to(DEV)
    parity.load_params(pol, params)
    act.actor = pol
    act.reset_agent()
    draws = _twin_draws(act.generator.initial_seed(), E, A, warm, sum(calls), kind)
    ref = OA.ActorOracle(inputs.ScriptedEnv(seed, E, O, A, T), E, O, A, params, nstep=3, noise_type=kind,
                         reward_scale=0.01, tracker_len=11, replay_draws=draws)
    ref.reset_agent()

    def check(tag, got, exp):
        p_g, v_g, n_g = got
        p_e, v_e, n_e = exp
        assert n_g == n_e
        assert torch.equal(p_g.cpu(), p_e), tag
        for k in (0, 3, 4):
            assert torch.equal(v_g[k].cpu(), v_e[k]), (tag, k)
        assert parity.rel(v_g[1], v_e[1]) <= 1e-3, (tag, "actions", parity.rel(v_g[1], v_e[1]))
        assert parity.rel(v_g[2], v_e[2]) <= 1e-3, (tag, "rewards", parity.rel(v_g[2], v_e[2]))
        np.testing.assert_allclose(act.obs_rms.mean.cpu().numpy(), ref.obs_rms.mean.numpy(), rtol=4e-6, atol=1e-6)
        np.testing.assert_allclose(act.obs_rms.var.cpu().numpy(), ref.obs_rms.var.numpy(), rtol=4e-6, atol=1e-6)
        assert act.obs_rms.count == ref.obs_rms.count
        assert torch.equal(act.current_lengths.cpu(), ref.current_lengths)
        assert parity.rel(act.current_returns, ref.current_returns) <= 1e-3
        assert act.step_tracker.mean() == pytest.approx(ref.step_tracker.mean(), rel=1e-12)
        assert act.return_tracker.mean() == pytest.approx(ref.return_tracker.mean(), rel=2e-3, abs=1e-4)

    check("warm", act.explore_env(env, warm, random=True), ref.explore_env(warm, random=True))
    for j, Tj in enumerate(calls):
        check(f"call{j}", act.explore_env(env, Tj, random=False), ref.explore_env(Tj, random=False))
    # the last exploration draw is the twin generator's, bit for bit
    assert torch.equal(act._plan.noise.cpu(), draws[-1])
    assert ref.step_tracker.mean() > 0
    a = act.get_actions(act.obs, sample=False)
    e = OA.L.actor_forward(ref.obs_rms.normalize(ref.obs), params)
    assert a.shape == (E, A) and parity.rel(a, e) <= 1e-3
