from .common import AttrDict, DeviceTracker, default_pql_cfg  # noqa: F401
from .schedule_util import ExponentialSchedule, LinearSchedule  # noqa: F401
from .torch_util import RunningMeanStd  # noqa: F401
from .model_util import load_model, save_model  # noqa: F401
