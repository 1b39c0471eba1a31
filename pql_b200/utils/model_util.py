"""Checkpoint format of the reference: ``torch.save({'obs_rms': (mean, var, eps) | None,
'actor': actor.state_dict(), 'critic': critic.state_dict()})`` as written by
pql/utils/model_util.py:24-41 from pql/utils/evaluator.py:112-119, with the reference's state_dict
keys (``net.{0,2,4,6}.{weight,bias}``, ``net_q{1,2}.net...``).  A file written by the reference loads
into the pql_b200 modules and vice versa.  The W&B artifact upload / download around it is the
caller's business (wandb is not a dependency)."""
import torch


def save_model(path, actor, critic, rms, wandb_run=None, description=None, **_ignored):
    """model_util.py:24-41.  ``actor`` / ``critic`` are state_dicts (or modules), ``rms`` is
    ``obs_rms.get_states()`` or None.  Tensors are stored on the CPU."""
    def sd(m):
        m = m.state_dict() if hasattr(m, "state_dict") else m
        return {k: v.detach().cpu() for k, v in m.items()}

    if isinstance(actor, list):
        checkpoint = {'obs_rms': _rms_cpu(rms), 'critic': sd(critic)}
        for i, a in enumerate(actor):
            checkpoint[f'actor_{i}'] = sd(a)
    else:
        checkpoint = {'obs_rms': _rms_cpu(rms), 'actor': sd(actor), 'critic': sd(critic)}
    torch.save(checkpoint, path)
    return checkpoint


def _rms_cpu(rms):
    if rms is None:
        return None
    return tuple(x.detach().cpu() if torch.is_tensor(x) else x for x in rms)


def load_model(model, model_type, path):
    """model_util.py:9-21 without the artifact download: ``model_type`` in {'actor', 'critic',
    'obs_rms'}; ``model`` is the module (or RunningMeanStd) to fill.  Returns False when the file
    holds no such entry (the reference logs a warning and carries on)."""
    weights = torch.load(path, map_location="cpu")
    if model_type not in weights:
        return False
    if model_type == "obs_rms" and weights[model_type] is None:
        return False
    model.load_state_dict(weights[model_type])
    return True
