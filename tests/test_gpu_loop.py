"""GPU tier: the whole PQL loop (actor worker + V-learner + P-learner at the reference's 8 : 4 : 1
ratios, pql_b200/train.py) on a synthetic vectorised env whose reward depends on the action: the
deterministic policy must get measurably better - kernels, hand-offs and optimiser working together."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class _Space:
    def __init__(self, shape):
        self.shape = shape


class ReachEnv:
    """Contextual-bandit stub: obs ~ N(0, 1) i.i.d. per step, reward = -mean((a - tanh(obs[:, :A]))^2)."""

    def __init__(self, E, O, A, seed=0):
        self.E, self.O, self.A = E, O, A
        self.g = torch.Generator(device=DEV).manual_seed(seed)
        self.observation_space, self.action_space = _Space((O,)), _Space((A,))
        self.obs = None

    def _draw(self):
        return torch.randn(self.E, self.O, device=DEV, generator=self.g)

    def reset(self):
        self.obs = self._draw()
        return self.obs

    def cost(self, obs, action):
        return ((action - torch.tanh(obs[:, :self.A])) ** 2).mean(dim=1)

    def step(self, action):
        reward = -self.cost(self.obs, action)
        self.obs = self._draw()
        done = torch.zeros(self.E, device=DEV)
        return self.obs, reward, done, {}


@pytest.mark.parametrize("distl", [False, True])
def test_policy_improves_on_reach_env(distl):
    from pql_b200.train import LockStepTrainer
    from pql_b200.utils import default_pql_cfg
    E, O, A = 1024, 24, 4
    torch.manual_seed(7)
    # C51: rewards scaled so that the returns (about -cost / (1 - gamma)) stay inside [v_min, v_max]
    cfg = default_pql_cfg(num_envs=E, sim_device=DEV, batch_size=2048, memory_size=200_000, warm_up=8,
                          reward_scale=0.02 if distl else 1.0, distl=distl, v_min=-1.5, v_max=0.25)
    env = ReachEnv(E, O, A)
    tr = LockStepTrainer(env, cfg)
    probe = torch.randn(4096, O, device=DEV, generator=torch.Generator(device=DEV).manual_seed(1))

    def policy_cost():
        a = tr.actor_worker.obs_rms.normalize(probe)
        return float(env.cost(probe, tr.actor(a)).mean())

    tr.warm_up()
    before = policy_cost()
    info = tr.run(max_env_steps=tr.global_steps + 150 * E)
    after = policy_cost()
    print(f"distl={distl}: policy cost {before:.4f} -> {after:.4f}; {info}")
    assert info["train/critic_update_times"] >= 8 * 149 and info["train/actor_update_times"] >= 4 * 149
    assert torch.isfinite(tr.critic.arena.flat).all() and torch.isfinite(tr.actor.arena.flat).all()
    assert after < 0.6 * before
    assert tr.actor_worker.obs_rms.count == pytest.approx(1e-4 + (8 + 150) * E)


def test_lockstep_loop_matches_oracle_loop():
    """SURVEY section 4 'integration' row: K iterations of the whole loop (scripts/train_pql.py:95-158 at the exact
    8 : 4 : 1 schedule: explore_env -> n-step -> both update() exchanges -> 8 critic + 4 actor updates) through
    LockStepTrainer against the same loop built from the CPU oracles (oracle/actor.py, oracle/replay.py,
    oracle/learner.py, each pinned to the reference), fed the SAME random draws: the exploration noise is the twin
    generator's, replay indices and target-policy noise are read back from the plans after every learn().
      * ring / obs-ring pointer state and every action-independent column (obs, next_obs, done): bit-exact;
      * actions and n-step rewards in the ring (they pass through the policy forward): <= 1e-3;
      * the loss of EVERY one of the 24 + 12 updates along the way: <= 1e-3 (BASELINE.json's tolerance; the
        first iteration's: <= 5e-4)."""
    import numpy as np
    from oracle import actor as OA
    from oracle import learner as L
    from oracle import replay as R
    from pql_b200.train import LockStepTrainer
    from pql_b200.utils import default_pql_cfg
    from tests import parity
    from tests.golden import inputs
    from tests.test_gpu_actor import _twin_draws
    E, O, A, B, CAP, WARM, K = 256, 88, 16, 512, 4096, 4, 3
    torch.manual_seed(21)
    cfg = default_pql_cfg(num_envs=E, sim_device=DEV, batch_size=B, memory_size=CAP, warm_up=WARM, sync_loss=True, tracker_len=9)
    T = WARM + K
    env = inputs.ScriptedEnv(5, E, O, A, T, device=DEV)
    tr = LockStepTrainer(env, cfg)
    csd = {k: v.detach().cpu() for k, v in tr.critic.state_dict().items()}
    asd = {k: v.detach().cpu() for k, v in tr.actor.state_dict().items()}
    q1, q2 = L.params_from_state_dict(csd, "net_q1.net."), L.params_from_state_dict(csd, "net_q2.net.")
    actor0 = L.params_from_state_dict(asd, "net.")
    # ---- oracle side
    draws = _twin_draws(tr.actor_worker.generator.initial_seed(), E, A, WARM, K, "mixed")
    ov, op = L.VLearnerOracle(q1, q2), L.PLearnerOracle(actor0)
    oa = OA.ActorOracle(inputs.ScriptedEnv(5, E, O, A, T), E, O, A, [(w.detach(), b.detach()) for w, b in op.actor], nstep=3,
                        reward_scale=0.01, tracker_len=9, replay_draws=draws)
    ring, obs_ring = R.RingOracle(CAP, O, A), R.ObsRingOracle(CAP, O)
    log = []          # (kind, idx, noise or None, loss) of every CUDA update, in launch order

    def observer(kind):
        torch.cuda.synchronize()
        plan = (tr.v_learner if kind == "v" else tr.p_learner)._plan
        log.append((kind, plan.idx.cpu().clone(), plan.noise.cpu().clone() * 0.8 if kind == "v" else None, float(plan.loss.item())))
    tr.observer = observer

    def oracle_exchange(p_data, v_data):
        ring.insert(*(x.numpy() for x in v_data))
        obs_ring.insert(p_data.numpy())
        snap = lambda ps: [(w.detach().clone(), b.detach().clone()) for w, b in ps]      # noqa: E731
        norm = (oa.obs_rms.mean.clone(), oa.obs_rms.var.clone(), 1e-4)
        return snap(op.actor), snap(ov.q1), snap(ov.q2), norm

    oa.reset_agent()
    tr.warm_up()
    p_data, v_data, _ = oa.explore_env(WARM, random=True)
    actor_v, cq1, cq2, norm = oracle_exchange(p_data, v_data)
    worst = {"v": 0.0, "p": 0.0, "v_first": 0.0, "p_first": 0.0}
    for it in range(K):
        log.clear()
        tr.step()
        p_data, v_data, _ = oa.explore_env(1, random=False)
        actor_v, cq1, cq2, norm = oracle_exchange(p_data, v_data)
        for kind, idx, noise, got in log:
            if kind == "v":
                batch = tuple(torch.from_numpy(x) for x in ring.gather(idx.numpy()))
                ref = ov.learn(batch, noise, actor_v, norm)
            else:
                ref = op.learn(torch.from_numpy(obs_ring.gather(idx.numpy())), cq1, cq2, norm)
            err = abs(got - ref) / max(abs(ref), 1e-7 if kind == "v" else 1e-3)
            worst[kind] = max(worst[kind], err)
            if it == 0:
                worst[kind + "_first"] = max(worst[kind + "_first"], err)
        assert len(log) == 12
        oa.actor = [(w.detach(), b.detach()) for w, b in op.actor]          # the actor worker explores with the P-learner's current policy
        # replay state after this iteration
        mem = tr.v_learner.memory
        assert (mem.next_p, mem.if_full, mem.cur_capacity) == (ring.next_p, ring.if_full, ring.cur_capacity)
        pl = tr.p_learner
        assert (pl.next_p, pl.if_full, pl.cur_capacity) == (obs_ring.next_p, obs_ring.if_full, obs_ring.cur_capacity)
        n = ring.cur_capacity
        assert np.array_equal(mem.buf_obs[:n].cpu().numpy(), ring.buf_obs[:n])
        assert np.array_equal(mem.buf_next_obs[:n].cpu().numpy(), ring.buf_next_obs[:n])
        assert np.array_equal(mem.buf_done[:n].cpu().numpy().astype(np.float32).reshape(n, -1), ring.buf_done[:n].astype(np.float32).reshape(n, -1))
        assert np.array_equal(pl.memory[:n].cpu().numpy(), obs_ring.memory[:n])
        assert parity.rel(mem.buf_action[:n], torch.from_numpy(ring.buf_action[:n])) <= 1e-3
        assert parity.rel(mem.buf_reward[:n], torch.from_numpy(ring.buf_reward[:n])) <= 1e-3
    print("loop parity, worst relative loss error:", worst)
    # measured on B200: 8.3e-5 (critic) / 7.8e-5 (actor) over the three iterations
    assert worst["v_first"] <= 5e-4 and worst["p_first"] <= 5e-4, worst
    assert worst["v"] <= 1e-3 and worst["p"] <= 1e-3, worst


@pytest.mark.parametrize("distl", [False, True])
def test_step_graph_equals_individual_updates(distl):
    """LockStepTrainer replays all updates of an env step as ONE CUDA graph (two branches + the look-ahead sampler's
    side branches, train.py:_learn_block).  Same seeds, same schedule issued as twelve learn() calls (cfg.step_graph
    off): weights, Polyak target, loss windows, update counts and the rings must be bit-identical."""
    from pql_b200.train import LockStepTrainer
    from pql_b200.utils import default_pql_cfg
    E, O, A = 512, 24, 4

    def run(step_graph):
        torch.manual_seed(11)
        cfg = default_pql_cfg(num_envs=E, sim_device=DEV, batch_size=1024, memory_size=50_000, warm_up=4,
                              reward_scale=0.02 if distl else 1.0, distl=distl, v_min=-1.5, v_max=0.25)
        cfg.step_graph = step_graph
        cfg.learner_streams = True
        tr = LockStepTrainer(ReachEnv(E, O, A, seed=5), cfg)
        infos = [tr.step() for _ in range(5)]
        torch.cuda.synchronize()
        used = tr._block is not None
        v, p = tr.v_learner, tr.p_learner
        m, n = v.memory, v.memory.cur_capacity         # the ring is torch.empty storage: compare the filled rows' columns only
        ring = torch.cat([m.buf_obs[:n], m.buf_next_obs[:n], m.buf_action[:n], m.buf_reward[:n], m.buf_done[:n].float()], 1)
        return (v.critic.arena.flat.clone(), v._plan.t_flat.clone(), p.actor.arena.flat.clone(), v.loss_tracker.window.clone(),
                p.loss_tracker.window.clone(), ring, v.update_count, p.update_count,
                [(i["train/critic_loss"], i["train/actor_loss"]) for i in infos], used, int(v._plan.opt.step), int(p._plan.opt.step))

    a, b = run(True), run(False)
    assert a[9] and not b[9], "the per-step graph was not the path that ran"
    for x, y in zip(a[:6], b[:6]):
        assert torch.equal(x, y)
    assert a[6:9] == b[6:9] and a[10:] == b[10:]
    assert a[6] == 5 * 8 and a[7] == 5 * 4 and a[10] == a[6] and a[11] == a[7]
