from .pql_p_learner import PQLPLearner  # noqa: F401
from .pql_v_learner import PQLVLearner  # noqa: F401
from .pql_actor import PQLActor  # noqa: F401
