"""CPU tier: the oracle (oracle/) against fixtures produced by the reference itself
(tests/golden/make_golden.py).  Replay / n-step / projection are bit-exact; everything
downstream of a GEMM is compared with a tight tolerance because MKL blocking differs with
the host's thread count."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import learner as L
from oracle import replay as R
from tests.golden import inputs


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def test_ring_insert_and_gather_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "replay_small.npz"))
    rb = R.RingOracle(50, 5, 2)
    for i, n in enumerate(g["sizes"]):
        rb.insert(*inputs.flat_rows(100 + i, int(n), 5, 2))
        assert [rb.next_p, int(rb.if_full), rb.cur_capacity] == g["ptrs"][i].tolist()
    for name in ("buf_obs", "buf_action", "buf_reward", "buf_next_obs", "buf_done"):
        assert np.array_equal(getattr(rb, name), g[name]), name
    s = rb.gather(g["idx"])
    for got, name in zip(s, ("s_obs", "s_action", "s_reward", "s_next_obs", "s_done")):
        assert got.dtype == np.float32 and np.array_equal(got, g[name]), name


@pytest.mark.parametrize("tag,cfg", [("n3", (8, 3, 5, 2, [1, 1, 1, 5, 32, 1])),
                                     ("n5", (6, 5, 4, 3, [7, 1, 9])),
                                     ("n1", (4, 1, 3, 2, [2]))])
def test_nstep_bit_exact(golden_dir, tag, cfg):
    g = np.load(os.path.join(golden_dir, "nstep_small.npz"))
    E, n, O, A, Ts = cfg
    ns = R.NStepOracle(O, A, E, n, 0.99)
    for j, T in enumerate(Ts):
        blk = inputs.transition_stream(200 + j, E, T, O, A, p_done=0.3)
        if f"{tag}_push{j}_empty" in g:
            with pytest.raises(ValueError):
                ns.push(*blk)
            continue
        res = ns.push(*blk)
        for name, r in zip(("obs", "act", "rew", "next", "done"), res):
            ref = g[f"{tag}_push{j}_{name}"]
            assert r.shape == ref.shape, (name, r.shape, ref.shape)
            if name == "rew" and n > 4:
                # for n > 4 torch's CPU sum kernel interleaves 4 accumulators (an ATen detail that
                # differs again on CUDA); the restatement sums left to right -> last-ulp differences
                np.testing.assert_allclose(r, ref, rtol=3e-7, atol=1e-9)
                continue
            assert np.array_equal(r.view(np.uint32), ref.view(np.uint32)), (tag, j, name)


def test_allegro_stream_digests(golden_dir):
    meta = json.load(open(os.path.join(golden_dir, "replay_allegro_stream.json")))
    E, O, A, C = meta["E"], meta["O"], meta["A"], meta["C"]
    ns, rb = R.NStepOracle(O, A, E, 3, 0.99), R.RingOracle(C, O, A)
    for step, T in enumerate([32] + [1] * 40):
        traj = ns.push(*inputs.transition_stream(300 + step, E, T, O, A, p_done=0.02))
        rb.insert(*traj)
        assert [rb.next_p, int(rb.if_full), rb.cur_capacity] == meta["ptrs"][step]
        assert sha(*traj) == meta["push_digests"][step], step
    assert sha(rb.buf_obs, rb.buf_action, rb.buf_reward, rb.buf_next_obs, rb.buf_done) == meta["ring_digest"]
    s = rb.gather(inputs.indices(11, rb.cur_capacity, 8192))
    assert sha(*s) == meta["sample_digest"]


def test_obs_ring_matches_transition_ring_rule():
    a, b = R.ObsRingOracle(50, 5), R.RingOracle(50, 5, 2)
    for i, n in enumerate([16, 16, 16, 16, 7, 50, 1]):
        rows = inputs.flat_rows(100 + i, n, 5, 2)
        a.insert(rows[0]); b.insert(*rows)
        assert (a.next_p, a.if_full, a.cur_capacity) == (b.next_p, b.if_full, b.cur_capacity)
    assert np.array_equal(a.memory, b.buf_obs)


def test_projection_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "projection_kat.npz"))
    out = L.projection(torch.from_numpy(g["dist"]), torch.from_numpy(g["reward"]),
                       torch.from_numpy(g["done"]), 0.99 ** 3)
    assert np.array_equal(out.numpy().view(np.uint32), g["out"].view(np.uint32))


def _digest(x):
    x = x.detach().float().reshape(-1)
    return np.array([x.sum().item(), x.abs().sum().item(), x.norm().item(),
                     *x[:6].tolist(), *x[-6:].tolist()], dtype=np.float64)


def _names(prefix_nets):
    return [f"{p}.net.{k}.{wb}" for p in prefix_nets for k in (0, 2, 4, 6) for wb in ("weight", "bias")]


@pytest.mark.parametrize("tag,seed,B,O,A,distl,steps", [("doubleq", 1234, 512, 88, 16, False, 3),
                                                        ("c51", 4321, 256, 88, 16, True, 3),
                                                        ("shadow", 77, 128, 211, 20, False, 2)])
def test_learner_oracle_vs_reference(golden_dir, tag, seed, B, O, A, distl, steps):
    g = np.load(os.path.join(golden_dir, f"learner_{tag}.npz"))
    case = inputs.learner_case(seed, B, O, A, distl)
    idx = torch.from_numpy(g["idx"])
    batch = tuple(x[idx] for x in case["batch"])
    v = L.VLearnerOracle(case["q1"], case["q2"], distl=distl)
    for s in range(steps):
        loss = v.learn(batch, case["noises"][s], case["actor"], case["norm"])
        assert loss == pytest.approx(g["v_losses"][s], rel=2e-5), (s, loss)
        gn = np.array([x.norm().item() for x in v.last["grads"]])
        np.testing.assert_allclose(gn, g["v_grad_norms"][s], rtol=2e-4, atol=1e-9)
    for name, t in zip(_names(["net_q1", "net_q2"]), L.flat([v.q1, v.q2])):
        np.testing.assert_allclose(_digest(t), g[f"critic.{name}"], rtol=2e-5, atol=2e-6, err_msg=name)
    for name, t in zip(_names(["net_q1", "net_q2"]), L.flat([v.tq1, v.tq2])):
        np.testing.assert_allclose(_digest(t), g[f"target.{name}"], rtol=2e-5, atol=2e-6, err_msg=name)
    p = L.PLearnerOracle(case["actor"], distl=distl)
    obs = case["batch"][0][idx]
    for s in range(steps):
        loss = p.learn(obs, case["q1"], case["q2"], case["norm"])
        assert loss == pytest.approx(g["p_losses"][s], rel=2e-5, abs=1e-7), (s, loss)
        gn = np.array([x.norm().item() for x in p.last["grads"]])
        np.testing.assert_allclose(gn, g["p_grad_norms"][s], rtol=2e-4, atol=1e-9)
    for name, t in zip([f"net.{k}.{wb}" for k in (0, 2, 4, 6) for wb in ("weight", "bias")], L.flat([p.actor])):
        np.testing.assert_allclose(_digest(t), g[f"actor.{name}"], rtol=2e-5, atol=2e-6, err_msg=name)


@pytest.mark.parametrize("noise_type,plain", [("mixed", False), ("fixed", False), ("mixed", True)])
def test_actor_oracle_vs_reference(golden_dir, noise_type, plain):
    """oracle.actor (RunningMeanStd, exploration noise, trackers, timeout handling, n-step hand-off)
    against the unmodified PQLActor run on the scripted env (tests/golden/make_golden.py::actor_small).
    Same torch CPU seed => same draws; everything else is fp32 arithmetic in the same order."""
    from oracle.actor import ActorOracle
    # plain: cfg.algo.obs_norm = False and handle_timeout = False (fixture actor_small_mixed_raw.npz)
    g = np.load(os.path.join(golden_dir, f"actor_small_{noise_type}{'_raw' if plain else ''}.npz"))
    c = inputs.ACTOR_CASE
    E, O, A = c["E"], c["O"], c["A"]
    env = inputs.ScriptedEnv(c["seed"], E, O, A, c["warm_up"] + sum(c["calls"]))
    torch.manual_seed(c["seed"])
    act = ActorOracle(env, E, O, A, inputs.actor_case_params(c["seed"], O, A), nstep=c["nstep"], noise_type=noise_type,
                      reward_scale=0.01, tracker_len=c["tracker_len"], obs_norm=not plain, do_handle_timeout=not plain)
    act.reset_agent()

    def check(tag, res):
        p_data, v_data, steps = res
        assert steps == int(g[f"{tag}_steps"])
        np.testing.assert_array_equal(p_data.numpy(), g[f"{tag}_p"])
        for name, x in zip(("obs", "next", "done"), (v_data[0], v_data[3], v_data[4])):
            np.testing.assert_array_equal(x.numpy(), g[f"{tag}_{name}"], err_msg=f"{tag} {name}")
        # actions / rewards pass through the policy MLP: MKL-vs-oracle matmul association only
        np.testing.assert_allclose(v_data[1].numpy(), g[f"{tag}_act"], rtol=0, atol=2e-6)
        np.testing.assert_allclose(v_data[2].numpy(), g[f"{tag}_rew"], rtol=0, atol=1e-7)
        if not plain:
            np.testing.assert_array_equal(act.obs_rms.mean.numpy(), g[f"{tag}_rms_mean"])
            np.testing.assert_array_equal(act.obs_rms.var.numpy(), g[f"{tag}_rms_var"])
            assert act.obs_rms.count == float(g[f"{tag}_rms_count"])
        np.testing.assert_allclose(np.array(list(act.return_tracker.moving_average), dtype=np.float64),
                                   g[f"{tag}_ret_window"], rtol=0, atol=1e-5)
        np.testing.assert_array_equal(np.array(list(act.step_tracker.moving_average), dtype=np.float64),
                                      g[f"{tag}_len_window"])
        np.testing.assert_allclose(act.current_returns.numpy(), g[f"{tag}_returns"], rtol=0, atol=1e-5)
        np.testing.assert_array_equal(act.current_lengths.numpy(), g[f"{tag}_lengths"])

    check("warm", act.explore_env(c["warm_up"], random=True))
    for j, T in enumerate(c["calls"]):
        check(f"call{j}", act.explore_env(T, random=False))
    assert (g["call5_len_window"] > 0).sum() >= 3          # the script does finish episodes


def test_noise_schedules_match_reference(golden_dir):
    """pql_b200.utils.{LinearSchedule, ExponentialSchedule} against value sequences recorded from
    pql/utils/schedule_util.py (host-side scalars: exact equality)."""
    from pql_b200.utils.schedule_util import ExponentialSchedule, LinearSchedule
    g = json.load(open(os.path.join(golden_dir, "schedules.json")))
    cases = {"linear_0.8_0.05_7": LinearSchedule(0.8, 0.05, 7), "linear_1_0_3": LinearSchedule(1.0, 0.0, 3),
             "exp_0.8_0.9_0.05": ExponentialSchedule(0.8, 0.9, 0.05), "exp_0.5_0.5_none": ExponentialSchedule(0.5, 0.5)}
    for name, sch in cases.items():
        vals = [sch.val()]
        for _ in range(40):
            vals.append(sch.step())
            vals.append(sch.val())
        assert vals == g[name], name
