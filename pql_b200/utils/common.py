"""Host-side helpers mirroring pql/utils/common.py for the learner path."""
import numpy as np
import torch


class DeviceTracker:
    """Tracker whose window lives on the GPU: the last kernel of every update writes its loss into
    ``window[update_index % max_len]`` (pqlb_grad_reduce_finish) instead of the reference's
    ``loss.item()`` host sync after every update (pql_v_learner.py:111).  ``mean()`` reads the window
    (a host synchronisation); ``mean_lagged()`` - what ``update()`` uses once per env step unless
    ``cfg.sync_loss`` - does not block.  Same pre-filled-with-zeros window semantics as the reference's
    Tracker (pql/utils/common.py:103-126)."""

    def __init__(self, max_len, device):
        self.max_len = max_len
        self.window = torch.zeros(max_len, dtype=torch.float32, device=device)
        self._host, self._event, self._last = None, None, 0.0

    def mean(self):
        return float(np.mean(self.window.tolist()))

    def mean_lagged(self):
        """Non-blocking read-back: returns the window mean as of the PREVIOUS call and starts the
        device-to-host copy (pinned memory, current stream) that the next call will read.  The caller
        never waits for the work it has just enqueued, so it can keep one env step of launches queued
        ahead of the GPU.  In the reference the main loop receives a learner's loss asynchronously as
        well (``ray.wait(..., timeout=0)``, scripts/train_pql.py:101-109)."""
        if self._host is None:
            self._host = torch.zeros(self.max_len, dtype=torch.float32).pin_memory()
        if self._event is not None:
            self._event.synchronize()            # the copy of the previous call: long done
            self._last = float(np.mean(self._host.numpy().astype(np.float64)))
        self._host.copy_(self.window, non_blocking=True)
        self._event = torch.cuda.Event()
        self._event.record()
        return self._last

    def std(self):
        return float(np.std(self.window.tolist()))

    def max(self):
        return float(np.max(self.window.tolist()))


class AttrDict(dict):
    """Minimal stand-in for the OmegaConf DictConfig the reference passes around."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def default_pql_cfg(**overrides):
    """Values of pql/cfg/algo/pql_algo.yaml + actor_critic.yaml + default.yaml that the learner
    path reads."""
    cfg = AttrDict(available_gpus=1, artifact=None, num_envs=4096, sim_device="cuda:0", info_track_keys=None, seed=42,
                   algo=AttrDict(v_learner_gpu=0, p_learner_gpu=0, distl=False, cri_class="DoubleQ",
                                 act_class="TanhMLPPolicy", v_min=-10, v_max=10, num_atoms=51, critic_lr=5e-4,
                                 actor_lr=5e-4, memory_size=int(5e6), batch_size=8192, obs_norm=True, gamma=0.99,
                                 nstep=3, tau=0.05, max_grad_norm=0.5, tracker_len=100, reward_scale=0.01,
                                 handle_timeout=True, warm_up=32, horizon_len=1, critic_actor_ratio=2,
                                 critic_sample_ratio=8,
                                 noise=AttrDict(type="mixed", decay=None, std_max=0.8, std_min=0.05, tgt_pol_std=0.8,
                                                tgt_pol_noise_bound=0.2)))
    for k, v in overrides.items():
        if k in cfg.algo:
            cfg.algo[k] = v
        else:
            cfg[k] = v
    return cfg
