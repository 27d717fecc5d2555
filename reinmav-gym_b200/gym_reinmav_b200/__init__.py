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

def VecNormalize(*args, **kwargs):
    """baselines' VecNormalize on the device (run.py:91-92; imports torch lazily)."""
    from .vec_normalize import VecNormalize as _VN
    return _VN(*args, **kwargs)


__version__ = "0.2.0"

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


def make_vec_env(env_id, env_type="native", num_env=1, seed=0, wrapper_kwargs=None, env_kwargs=None, start_index=0,
                 reward_scale=1.0, flatten_dict_observations=True, gamestate=None, initializer=None, force_dummy=False,
                 **options):
    """baselines' ``make_vec_env`` for the native ids, with its signature — the reference's launcher calls it as
    ``make_vec_env(env_id, env_type, args.num_env or 1, seed, reward_scale=args.reward_scale,
    flatten_dict_observations=flatten_dict_observations)`` (/root/reference/gym_reinmav/run.py:89) and that call works
    here verbatim.  What the baselines arguments mean on this path:

    * ``num_env``, ``seed``: N envs in ONE handle; ``seed`` keys the Philox reset stream (``None`` -> 0).
    * ``start_index``: baselines seeds sub-env ``i`` with ``seed + start_index + i``; here it is the global id of env 0
      (``env_id_offset``), which is what the reset stream is keyed by.
    * ``reward_scale``: baselines wraps each env in ``RewardScaler`` OUTSIDE ``Monitor`` when it is not 1 — the returned
      rewards are multiplied inside the step kernel (``rmv_params.reward_scale``), ``info['episode']['r']`` is not.
    * ``flatten_dict_observations``: observations are flat Boxes already (nothing to flatten).
    * ``force_dummy``: irrelevant (there is no process fan-out).
    * ``wrapper_kwargs`` / ``env_kwargs`` / ``gamestate`` / ``initializer``: atari / retro / user hooks that have no
      meaning for these envs; anything but ``None`` / ``{}`` is refused rather than silently dropped.

    ``options`` are QuadVecEnv's own keywords (``device``, ``params``, ``monitor``, ``return_obs``, ``copy_outputs``,
    ``state_layout``)."""
    if env_type not in (None, "native"):
        raise ValueError("env_type %r: this package provides the 'native' envs only" % (env_type,))
    for name, value in (("wrapper_kwargs", wrapper_kwargs), ("env_kwargs", env_kwargs), ("gamestate", gamestate),
                        ("initializer", initializer)):
        if value:
            raise ValueError("make_vec_env(%s=%r) has no meaning for the native quadrotor envs" % (name, value))
    params = dict(options.pop("params", None) or {})
    if reward_scale is not None and float(reward_scale) != 1.0:
        params["reward_scale"] = float(reward_scale)
    if "env_id_offset" in options:
        start_index = options.pop("env_id_offset")
    return QuadVecEnv(env_id, int(num_env or 1), seed=0 if seed is None else seed, env_id_offset=int(start_index),
                      params=params or None, **options)


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
