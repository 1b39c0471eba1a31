from .simple_replay import ReplayBuffer, create_buffer  # noqa: F401
from .nstep_replay import NStepReplay  # noqa: F401
