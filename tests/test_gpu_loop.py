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
