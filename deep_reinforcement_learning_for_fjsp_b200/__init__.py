"""B200-native batched FJSP scheduling environment (drop-in for the reference's
environments/SO_DFJSP.py, MO_DFJSP.py, MO_DFJSP_breakdown.py and the process-pool rollout
of utilities/Parallel_Experience_Generator.py)."""
from .instance import FJSPInstance  # noqa: F401


def __getattr__(name):
    if name in ("FJSPVecEnv", "FJSPEnv", "MyError"):
        from . import vec_env
        return getattr(vec_env, name)
    raise AttributeError(name)
