from .min_snap import Trajectory, plan, plan_named  # noqa: F401
from .paths import PATHS  # noqa: F401
