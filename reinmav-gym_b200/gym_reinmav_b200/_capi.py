"""ctypes binding of librmv.so (include/rmv.h).  No torch, no NumPy arithmetic here.

There is deliberately no fallback: if the shared library is missing or a call fails the
caller gets an exception (``RmvError``) carrying ``rmv_last_error()``.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RMV_LIB", os.path.join(_HERE, "lib", "librmv.so"))

QUAD2D, QUAD2D_SLUNG, QUAD3D, QUAD3D_SLUNG, REINMAV = 0, 1, 2, 3, 4
ACT_AOS, ACT_SOA = 0, 1
MANY_STREAM, MANY_ROLLOUT, MANY_LAUNCHES = 0, 1, 2
VARIANTS = {-1: "none", 0: "k_step", 1: "k_step_rows", 2: "k_step_tma", 3: "k_rollout", 4: "k_step_stream", 5: "k_step_stream_pair"}
TRACK_EPISODES = 1
STATE_ROWS = 2
EVER_DONE_BIT = 0x80000000
ABI_VERSION = 2


class RmvError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("librmv error %d: %s" % (code, message))
        self.code = code


class Params(ctypes.Structure):
    """struct rmv_params (include/rmv.h)."""
    _fields_ = [("mass", ctypes.c_float), ("load_mass", ctypes.c_float), ("dt", ctypes.c_float),
                ("gravity", ctypes.c_float), ("tether_len", ctypes.c_float), ("pos_done", ctypes.c_float),
                ("vel_done", ctypes.c_float), ("thrust_scale", ctypes.c_float),
                ("thrust_clamp", ctypes.c_int32), ("first_done_reward", ctypes.c_float),
                ("later_done_reward", ctypes.c_float), ("reward_scale", ctypes.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_vp, _i, _i64, _u64, _u32, _f = (ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint64,
                                 ctypes.c_uint32, ctypes.c_float)
_pi, _pi64, _pf, _pd = ctypes.POINTER(_i), ctypes.POINTER(_i64), ctypes.POINTER(_f), ctypes.POINTER(ctypes.c_double)

# name -> (restype, argtypes).  Every symbol include/rmv.h declares must be listed here;
# tests/test_capi_symbols.py checks the two against each other.
SIGNATURES = {
    "rmv_abi_version": (_i, []),
    "rmv_last_error": (ctypes.c_char_p, []),
    "rmv_kind_dims": (_i, [_i, _pi, _pi]),
    "rmv_default_params": (_i, [_i, ctypes.POINTER(Params)]),
    "rmv_padded_ld": (_i64, [_i64]),
    "rmv_action_bounds": (_i, [_i, _pf, _pf]),
    "rmv_create": (_i, [ctypes.POINTER(_vp), _i, _i64, _i64, _i, _u64, ctypes.POINTER(Params), _u32]),
    "rmv_destroy": (_i, [_vp]),
    "rmv_dims": (_i, [_vp, _pi, _pi, _pi64, _pi64]),
    "rmv_get_params": (_i, [_vp, ctypes.POINTER(Params)]),
    "rmv_set_seed": (_i, [_vp, _u64]),
    "rmv_stream_pays": (_i, [_i, _i64]),
    "rmv_last_variant": (_i, [_vp, _pi]),
    "rmv_get_time": (_i, [_vp, ctypes.POINTER(ctypes.c_double)]),
    "rmv_set_time": (_i, [_vp, ctypes.c_double]),
    "rmv_set_tuning": (_i, [_vp, _i, _i]),
    "rmv_set_host_chunks": (_i, [_vp, _i]),
    "rmv_launch_shape": (_i, [_i, _i64, ctypes.POINTER(_i), ctypes.POINTER(_i)]),
    "rmv_set_kernel": (_i, [_vp, _i, _i]),
    "rmv_set_prefetch": (_i, [_vp, _i]),
    "rmv_set_l2_persist": (_i, [_vp, _i]),
    "rmv_state_ptr": (_i, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_vp)]),
    "rmv_set_state": (_i, [_vp, _vp, _i, _vp]),
    "rmv_get_state": (_i, [_vp, _vp, _i, _vp]),
    "rmv_set_episode": (_i, [_vp, _vp, _i, _vp]),
    "rmv_get_episode": (_i, [_vp, _vp, _i, _vp]),
    "rmv_reset": (_i, [_vp, _vp, _vp, _vp]),
    "rmv_step": (_i, [_vp, _vp, _i, _i64, _vp, _vp, _vp, _i, _vp]),
    "rmv_step_rows": (_i, [_vp, _vp, _vp, _i, _i64, _vp, _vp, _vp, _i, _vp]),
    "rmv_step_many": (_i, [_vp, _i, _vp, _i, _i64, _vp, _vp, _i64, _i, _vp]),
    "rmv_set_stream": (_i, [_vp, _i, _i, _i]),
    "rmv_stream_shape": (_i, [_i64, _i, _i, _pi, _pi, _pi, _pi]),
    "rmv_step_host": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "rmv_reset_host": (_i, [_vp, _vp, _vp]),
    "rmv_control": (_i, [_vp, _pf, _pf, _vp, _i, _i64, _vp]),
    "rmv_control_step": (_i, [_vp, _pf, _pf, _vp, _vp, _vp, _vp, _i, _vp]),
    "rmv_control_step_many": (_i, [_vp, _i, _pf, _pf, _vp, _vp, _vp, _i64, _vp]),
    "rmv_control_host": (_i, [_vp, _pf, _pf, _vp, _vp]),
    "rmv_fill_uniform_actions": (_i, [_vp, _i64, _i, _i, _i64, _f, _f, _u64, _i64, _u32, _i, _vp]),
    "rmv_episode_stats": (_i, [_vp, ctypes.POINTER(ctypes.c_double), _vp]),
    "rmv_gather_episodes": (_i, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "rmv_episode_ptr": (_i, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_vp)]),
    "rmv_norm_create": (_i, [ctypes.POINTER(_vp), _i64, _i, _i, _i, _i, _f, _f, ctypes.c_double, ctypes.c_double]),
    "rmv_norm_destroy": (_i, [_vp]),
    "rmv_norm_step": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "rmv_norm_reset": (_i, [_vp, _vp, _vp, _i, _vp]),
    "rmv_norm_get": (_i, [_vp, _pd, _pd, _pd, _pd, _pd, _pd, _vp]),
    "rmv_norm_set": (_i, [_vp, _pd, _pd, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, _vp]),
    "rmv_host_alloc": (_i, [ctypes.POINTER(_vp), ctypes.c_size_t]),
    "rmv_host_free": (_i, [_vp]),
}

_lib = None


def lib():
    """Load librmv.so once.  Raises if it has not been built — there is no other path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "librmv.so not found at %s — build it with `make -C reinmav-gym_b200/csrc` "
                "(or __graft_entry__.build()); this package has no CPU implementation" % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        if handle.rmv_abi_version() != ABI_VERSION:
            raise ImportError("librmv.so ABI version %d, expected %d" % (handle.rmv_abi_version(), ABI_VERSION))
        _lib = handle
    return _lib


def check(code):
    if code != 0:
        raise RmvError(code, lib().rmv_last_error().decode("utf-8", "replace"))
    return code


def kind_dims(kind):
    s, a = _i(), _i()
    check(lib().rmv_kind_dims(kind, ctypes.byref(s), ctypes.byref(a)))
    return s.value, a.value


def default_params(kind):
    p = Params()
    check(lib().rmv_default_params(kind, ctypes.byref(p)))
    return p


def action_bounds(kind):
    lo, hi = _f(), _f()
    check(lib().rmv_action_bounds(kind, ctypes.byref(lo), ctypes.byref(hi)))
    return lo.value, hi.value
