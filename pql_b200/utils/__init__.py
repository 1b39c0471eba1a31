from .common import AttrDict, DeviceTracker, Tracker, default_pql_cfg  # noqa: F401
