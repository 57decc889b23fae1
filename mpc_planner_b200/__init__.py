"""B200-native batched SQP backend for the NLP solve loop of tud-amr/mpc_planner."""
from .problem import (  # noqa: F401
    CONFIGS,
    ProblemDesc,
    ProblemSpec,
    SolveInfo,
    config_basic,
    config_safe_horizon,
    config_tmpc,
    config_tmpc_decomp,
)
