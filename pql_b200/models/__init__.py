"""Class lookup by name, like pql/models/__init__.py + load_class_from_path
(pql/utils/common.py:34-42) do for cfg.algo.cri_class / cfg.algo.act_class."""
from .mlp import DistributionalDoubleQ, DoubleQ, MLPNet, TanhMLPPolicy  # noqa: F401

model_name_to_class = {c.__name__: c for c in (MLPNet, TanhMLPPolicy, DoubleQ, DistributionalDoubleQ)}


def load_class(name):
    try:
        return model_name_to_class[name]
    except KeyError:
        raise KeyError(f"{name!r} is not on the PQL learner path; available: {sorted(model_name_to_class)}")
