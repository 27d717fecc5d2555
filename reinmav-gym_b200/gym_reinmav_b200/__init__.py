"""gym_reinmav_b200 — B200-native batched drop-in for reinmav-gym's native quadrotor envs.

Registered ids (same strings as /root/reference/gym_reinmav/__init__.py:8-26):
``quadrotor2d-v0``, ``quadrotor2d-slungload-v0``, ``quadrotor3d-v0``, ``quadrotor3d-slungload-v0``
and ``reinmav-v0`` (/root/reference/gym_reinmav/__init__.py:3-6).

    import gym_reinmav_b200 as rmv
    env = rmv.make('quadrotor3d-v0')                         # gym.Env view, N = 1
    venv = rmv.make_vec_env('quadrotor3d-v0', num_env=1 << 20)   # baselines-style VecEnv, one kernel per step

If ``gym`` (or ``gymnasium``) is installed the ids are also registered there, so
``gym.make('quadrotor3d-v0')`` resolves to these classes.
"""
from ._capi import LIB_PATH, RmvError  # noqa: F401
from .batch import ENV_IDS, QuadBatch, fill_uniform_actions  # noqa: F401
from .vec_env import QuadVecEnv, SparseInfos  # noqa: F401


def RolloutBuffer(*args, **kwargs):
    """On-device PPO2-shaped rollout buffers (imports torch lazily)."""
    from .rollout import RolloutBuffer as _RB
    return _RB(*args, **kwargs)

__version__ = "0.1.0"

ENTRY_POINTS = {
    "quadrotor2d-v0": "gym_reinmav_b200.envs.native:Quadrotor2D",
    "quadrotor2d-slungload-v0": "gym_reinmav_b200.envs.native:Quadrotor2DSlungload",
    "quadrotor3d-v0": "gym_reinmav_b200.envs.native:Quadrotor3D",
    "quadrotor3d-slungload-v0": "gym_reinmav_b200.envs.native:Quadrotor3DSlungload",
    "reinmav-v0": "gym_reinmav_b200.envs.native:ReinmavEnv",
}


def make(env_id, **kwargs):
    """``gym.make`` for the four native ids (no TimeLimit wrapper: the reference registers
    no ``max_episode_steps``)."""
    if env_id not in ENTRY_POINTS:
        raise KeyError("unknown env id %r (have %s)" % (env_id, sorted(ENTRY_POINTS)))
    from importlib import import_module
    mod, cls = ENTRY_POINTS[env_id].split(":")
    return getattr(import_module(mod), cls)(**kwargs)


def make_vec_env(env_id, env_type="native", num_env=1, seed=0, device=0, **kwargs):
    """Same call shape as baselines' ``make_vec_env(env_id, env_type, num_env, seed, ...)`` used at
    /root/reference/gym_reinmav/run.py:89."""
    return QuadVecEnv(env_id, num_env, device=device, seed=0 if seed is None else seed, **kwargs)


def _register_with_gym():
    for mod in ("gym", "gymnasium"):
        try:
            reg = __import__(mod + ".envs.registration", fromlist=["register"])
        except Exception:
            continue
        for env_id, entry in ENTRY_POINTS.items():
            try:
                reg.register(id=env_id, entry_point=entry)
            except Exception:
                pass                                  # already registered (e.g. by the reference package)


_register_with_gym()
