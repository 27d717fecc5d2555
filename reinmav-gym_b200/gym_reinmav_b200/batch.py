"""QuadBatch — Python owner of one ``rmv_env`` handle (N envs of one kind on one GPU).

Thin by design: it converts arguments to raw device pointers / a stream handle and calls the
C-ABI.  Device buffers may be passed as ``torch.Tensor`` (CUDA), as any object with
``__cuda_array_interface__``, or as an integer address.
"""
import ctypes

import numpy as np

from . import _capi
from ._capi import ACT_AOS, ACT_SOA, check, lib

ENV_IDS = {
    "quadrotor2d-v0": _capi.QUAD2D,
    "quadrotor2d-slungload-v0": _capi.QUAD2D_SLUNG,
    "quadrotor3d-v0": _capi.QUAD3D,
    "quadrotor3d-slungload-v0": _capi.QUAD3D_SLUNG,
    "reinmav-v0": _capi.REINMAV,
}


HOST = "host"


def _ptr(x, dtype=None, where=None, what="buffer"):
    """Raw address of a device (or pinned host) buffer; None -> NULL.

    The C-ABI takes plain pointers, so this is the only layer that can refuse a buffer of the wrong element type or on
    the wrong side of the bus before a kernel reinterprets it (a float64 action tensor read as float32, a CPU tensor's
    address dereferenced on the GPU — an illegal-address fault kills the CUDA context).  ``dtype``: expected element
    type name ("float32", "uint8", ...); ``where``: a device ordinal (the buffer must live on that GPU), ``HOST`` (it
    must be host memory) or None (not checked).  Integer addresses cannot be checked."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):                       # torch.Tensor
        if not x.is_contiguous():
            raise ValueError("%s: tensor passed to librmv must be contiguous" % what)
        if dtype is not None and str(x.dtype) != "torch." + dtype:
            raise TypeError("%s must be %s, got %s" % (what, dtype, x.dtype))
        if where == HOST and x.is_cuda:
            raise ValueError("%s must be a host buffer, got a tensor on %s" % (what, x.device))
        if isinstance(where, int) and (not x.is_cuda or x.device.index != where):
            raise ValueError("%s must be on cuda:%d, got a tensor on %s" % (what, where, x.device))
        return x.data_ptr()
    cai = getattr(x, "__cuda_array_interface__", None)
    if cai is not None:
        if dtype is not None and np.dtype(cai["typestr"]) != np.dtype(dtype):
            raise TypeError("%s must be %s, got %s" % (what, dtype, np.dtype(cai["typestr"])))
        if where == HOST:
            raise ValueError("%s must be a host buffer, got a device array" % what)
        return cai["data"][0]
    if isinstance(x, np.ndarray):                    # host buffer (for the *_host entry points)
        if not x.flags["C_CONTIGUOUS"]:
            raise ValueError("%s: array passed to librmv must be C-contiguous" % what)
        if dtype is not None and x.dtype != np.dtype(dtype):
            raise TypeError("%s must be %s, got %s" % (what, dtype, x.dtype))
        if isinstance(where, int):
            raise ValueError("%s must be a device buffer on cuda:%d, got a NumPy array" % (what, where))
        return x.ctypes.data
    raise TypeError("cannot take the address of %r" % type(x))


def _need(x, count, what):
    """Refuse a tensor / array that is too small for what the kernel is about to read or write (raw addresses cannot
    be checked).  The C-ABI takes plain pointers, so this is the only place an undersized buffer can be caught."""
    if x is None or isinstance(x, int):
        return
    have = x.numel() if hasattr(x, "numel") else getattr(x, "size", None)
    if have is not None and have < count:
        raise ValueError("%s holds %d elements, %d needed" % (what, have, count))


_TORCH_CUDA = None          # torch, once it is imported and reports a CUDA device (asked once: this sits on the per-step path)


def _stream(stream, device):
    if stream is None:
        global _TORCH_CUDA
        if _TORCH_CUDA is None:
            import sys
            torch = sys.modules.get("torch")
            if torch is None or not torch.cuda.is_available():
                return None
            _TORCH_CUDA = torch
        return _TORCH_CUDA.cuda.current_stream(device).cuda_stream
    if isinstance(stream, int):
        return stream or None
    return stream.cuda_stream                        # torch.cuda.Stream


class _DeviceView:
    """Zero-copy window onto handle-owned device memory (``torch.as_tensor(view, device=...)``)."""

    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False),
                                         "version": 2, "strides": None}
        self._owner = owner                           # keeps the handle alive


class QuadBatch:
    def __init__(self, kind, num_envs, device=0, seed=0, env_id_offset=0, params=None, track_episodes=False,
                 state_layout="tiles"):
        """``state_layout``: "tiles" (float32 [ld/128, S, 128], the default) or "rows" (float32 [ld, S]: the state
        buffer is the observation matrix itself, stepped in place — ``RMV_STATE_ROWS``)."""
        if isinstance(kind, str):
            if kind not in ENV_IDS:
                raise KeyError("unknown env id %r (have %s)" % (kind, sorted(ENV_IDS)))
            kind = ENV_IDS[kind]
        self.kind, self.device, self.seed = int(kind), int(device), int(seed)
        self._h = ctypes.c_void_p()
        p = None
        if params is not None:
            p = _capi.default_params(self.kind)
            for k, v in dict(params).items():
                if not hasattr(p, k):
                    raise KeyError("rmv_params has no field %r" % k)
                setattr(p, k, v)
        if state_layout not in ("tiles", "rows"):
            raise ValueError("state_layout must be 'tiles' or 'rows', got %r" % (state_layout,))
        self.state_layout = state_layout
        flags = (_capi.TRACK_EPISODES if track_episodes else 0) | (_capi.STATE_ROWS if state_layout == "rows" else 0)
        check(lib().rmv_create(ctypes.byref(self._h), self.kind, int(num_envs), int(env_id_offset), self.device,
                               self.seed & 0xFFFFFFFFFFFFFFFF, ctypes.byref(p) if p is not None else None, flags))
        s, a, n, ld = ctypes.c_int(), ctypes.c_int(), ctypes.c_int64(), ctypes.c_int64()
        check(lib().rmv_dims(self._h, ctypes.byref(s), ctypes.byref(a), ctypes.byref(n), ctypes.byref(ld)))
        self.state_dim, self.action_dim, self.num_envs, self.ld = s.value, a.value, n.value, ld.value
        self.env_id_offset = int(env_id_offset)
        self.track_episodes = bool(track_episodes)
        self.launches = 0                               # kernels launched through this handle

    # ---------------------------------------------------------------- lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().rmv_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def params(self):
        p = _capi.Params()
        check(lib().rmv_get_params(self._h, ctypes.byref(p)))
        return p.as_dict()

    def set_seed(self, seed):
        self.seed = int(seed)
        check(lib().rmv_set_seed(self._h, self.seed & 0xFFFFFFFFFFFFFFFF))

    @property
    def last_kernel(self):
        """Name of the kernel family the most recent step call launched ("k_step_stream_pair", ...)."""
        v = ctypes.c_int()
        check(lib().rmv_last_variant(self._h, ctypes.byref(v)))
        return _capi.VARIANTS.get(v.value, str(v.value))

    def _one_launch(self):
        """Whether the multi-step call just made was ONE kernel launch (the library decides: step stream / rollout kernel
        against one launch per step, by kind, batch size and state layout)."""
        return self.last_kernel in ("k_step_stream", "k_step_stream_pair", "k_rollout")

    def set_tuning(self, vec=0, block=0):
        check(lib().rmv_set_tuning(self._h, int(vec), int(block)))

    @property
    def sim_time(self):
        """reinmav-v0: the env clock ``t`` (seconds)."""
        t = ctypes.c_double()
        check(lib().rmv_get_time(self._h, ctypes.byref(t)))
        return t.value

    @sim_time.setter
    def sim_time(self, t):
        check(lib().rmv_set_time(self._h, float(t)))

    def set_kernel(self, tma=False, ctas_per_sm=0):
        """Pick the step-kernel variant (register path / TMA-pipelined persistent); results are identical."""
        check(lib().rmv_set_kernel(self._h, 1 if tma else 0, int(ctas_per_sm)))

    def set_prefetch(self, mode=-1):
        """Episode-word prefetch of the step kernel: -1 auto, 0 off, 1 on; results are identical."""
        check(lib().rmv_set_prefetch(self._h, int(mode)))

    def set_l2_persist(self, mode=-1):
        """L2 access-policy window on the state (state stays L2-resident between steps): -1 auto, 0 off, 1 on."""
        check(lib().rmv_set_l2_persist(self._h, int(mode)))

    def set_stream(self, warps_per_sm=0, block=0, action_prefetch=-1):
        """Tuning of the step-stream kernel behind ``step_many(fused=0)``; results are identical."""
        check(lib().rmv_set_stream(self._h, int(warps_per_sm), int(block), int(action_prefetch)))

    def set_host_chunks(self, chunks=0):
        check(lib().rmv_set_host_chunks(self._h, int(chunks)))

    # ---------------------------------------------------------------- hot path (device buffers)
    def reset(self, mask=None, obs=None, stream=None):
        d = self.device
        check(lib().rmv_reset(self._h, _ptr(mask, "uint8", d, "mask"), _ptr(obs, "float32", d, "obs"), _stream(stream, d)))
        self.launches += 1

    def _check_io(self, actions, reward, done, layout, act_ld, steps=1, out_ld=0):
        n, A = self.num_envs, self.action_dim
        per_step = A * (int(act_ld) if layout == ACT_SOA and act_ld else n)
        _need(actions, steps * per_step, "actions")
        _need(reward, (steps - 1) * int(out_ld or n) + n, "reward")
        _need(done, (steps - 1) * int(out_ld or n) + n, "done")

    def step(self, actions, reward, done, obs=None, auto_reset=True, layout=ACT_AOS, act_ld=0, stream=None):
        self._check_io(actions, reward, done, layout, act_ld)
        _need(obs, self.num_envs * self.state_dim, "obs")
        d = self.device
        check(lib().rmv_step(self._h, _ptr(actions, "float32", d, "actions"), layout, int(act_ld),
                             _ptr(reward, "float32", d, "reward"), _ptr(done, "uint8", d, "done"),
                             _ptr(obs, "float32", d, "obs"), 1 if auto_reset else 0, _stream(stream, d)))
        self.launches += 1

    def step_rows(self, rows_in, actions, reward, done, rows_out, auto_reset=True, layout=ACT_AOS, act_ld=0, stream=None):
        """Step on row-major state: read [N, S] rows_in, write [N, S] rows_out (may alias); the handle's own tiled
        state is not touched (``set_state_device(rows_out)`` re-synchronises it)."""
        self._check_io(actions, reward, done, layout, act_ld)
        _need(rows_in, self.num_envs * self.state_dim, "rows_in")
        _need(rows_out, self.num_envs * self.state_dim, "rows_out")
        d = self.device
        check(lib().rmv_step_rows(self._h, _ptr(rows_in, "float32", d, "rows_in"), _ptr(actions, "float32", d, "actions"),
                                  layout, int(act_ld), _ptr(reward, "float32", d, "reward"), _ptr(done, "uint8", d, "done"),
                                  _ptr(rows_out, "float32", d, "rows_out"), 1 if auto_reset else 0, _stream(stream, d)))
        self.launches += 1

    def set_state_device(self, rows, stream=None):
        """Load [N, S] device rows into the handle's tiled state."""
        check(lib().rmv_set_state(self._h, _ptr(rows, "float32", self.device, "rows"), 1, _stream(stream, self.device)))

    def step_many(self, n_steps, actions, reward=None, done=None, out_ld=0, layout=ACT_SOA, act_ld=0, fused=False,
                  stream=None):
        """``actions`` is [T, A, act_ld] (``ACT_SOA``, the default; ``act_ld`` 0 -> the handle's padded env count
        ``self.ld``) or [T, N, A] (``ACT_AOS``).  ``fused``: False / 0 = the persistent step stream (state through
        memory every step, one launch), True / 1 = state in registers for the whole call, 2 = one launch per step."""
        if layout == ACT_SOA and not act_ld:
            act_ld = self.ld
        if self.action_dim:
            self._check_io(actions, reward, done, layout, act_ld, steps=int(n_steps), out_ld=out_ld)
        d = self.device
        check(lib().rmv_step_many(self._h, int(n_steps), _ptr(actions, "float32", d, "actions"), layout, int(act_ld),
                                  _ptr(reward, "float32", d, "reward"), _ptr(done, "uint8", d, "done"),
                                  int(out_ld or self.num_envs), int(fused), _stream(stream, d)))
        self.launches += 1 if self._one_launch() else int(n_steps)

    # ---------------------------------------------------------------- hot path (host buffers)
    def step_host(self, actions, reward, done, obs=None, auto_reset=True, stream=None):
        check(lib().rmv_step_host(self._h, _ptr(actions, "float32", HOST, "actions"), _ptr(reward, "float32", HOST, "reward"),
                                  _ptr(done, "uint8", HOST, "done"), _ptr(obs, "float32", HOST, "obs"),
                                  1 if auto_reset else 0, _stream(stream, self.device)))
        self.launches += 1

    def reset_host(self, obs, stream=None):
        check(lib().rmv_reset_host(self._h, _ptr(obs, "float32", HOST, "obs"), _stream(stream, self.device)))
        self.launches += 1

    # ---------------------------------------------------------------- controllers
    @staticmethod
    def _ref(v):
        if v is None:
            return None
        v = [float(x) for x in v]
        return (ctypes.c_float * len(v))(*v)

    def control(self, actions, ref_pos=None, ref_vel=None, layout=ACT_AOS, act_ld=0, stream=None):
        """The reference's geometric controller for every env -> device action buffer."""
        check(lib().rmv_control(self._h, self._ref(ref_pos), self._ref(ref_vel),
                                _ptr(actions, "float32", self.device, "actions"), layout, int(act_ld),
                                _stream(stream, self.device)))
        self.launches += 1

    def control_step(self, reward, done, obs=None, actions_out=None, ref_pos=None, ref_vel=None, auto_reset=True, stream=None):
        """``action = control(); step(action)`` as ONE kernel (no action buffer is read); ``actions_out`` (optional
        [N, A] device buffer) receives the controller's actions."""
        d = self.device
        _need(reward, self.num_envs, "reward"); _need(done, self.num_envs, "done")
        _need(obs, self.num_envs * self.state_dim, "obs"); _need(actions_out, self.num_envs * self.action_dim, "actions_out")
        check(lib().rmv_control_step(self._h, self._ref(ref_pos), self._ref(ref_vel), _ptr(actions_out, "float32", d, "actions_out"),
                                     _ptr(reward, "float32", d, "reward"), _ptr(done, "uint8", d, "done"),
                                     _ptr(obs, "float32", d, "obs"), 1 if auto_reset else 0, _stream(stream, d)))
        self.launches += 1

    def control_step_many(self, n_steps, reward=None, done=None, out_ld=0, actions_out=None, ref_pos=None, ref_vel=None, stream=None):
        """``n_steps`` closed-loop steps with reset-on-done in one launch of the step stream."""
        d, n = self.device, self.num_envs
        _need(reward, (int(n_steps) - 1) * int(out_ld or n) + n, "reward"); _need(done, (int(n_steps) - 1) * int(out_ld or n) + n, "done")
        _need(actions_out, int(n_steps) * n * self.action_dim, "actions_out")
        check(lib().rmv_control_step_many(self._h, int(n_steps), self._ref(ref_pos), self._ref(ref_vel),
                                          _ptr(actions_out, "float32", d, "actions_out"), _ptr(reward, "float32", d, "reward"),
                                          _ptr(done, "uint8", d, "done"), int(out_ld or n), _stream(stream, d)))
        self.launches += 1 if self._one_launch() else int(n_steps)

    def control_host(self, actions, ref_pos=None, ref_vel=None, stream=None):
        check(lib().rmv_control_host(self._h, self._ref(ref_pos), self._ref(ref_vel),
                                     _ptr(actions, "float32", HOST, "actions"), _stream(stream, self.device)))
        self.launches += 1

    # ---------------------------------------------------------------- state access
    def get_state_device(self, rows, stream=None):
        """Write the current state as [N, S] rows into a device buffer (a kernel on ``stream``; no synchronisation)."""
        _need(rows, self.num_envs * self.state_dim, "rows")
        check(lib().rmv_get_state(self._h, _ptr(rows, "float32", self.device, "rows"), 1, _stream(stream, self.device)))

    def get_state(self, stream=None):
        out = np.empty((self.num_envs, self.state_dim), dtype=np.float32)
        check(lib().rmv_get_state(self._h, out.ctypes.data, 0, _stream(stream, self.device)))
        return out

    def set_state(self, rows, stream=None):
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        if rows.shape != (self.num_envs, self.state_dim):
            raise ValueError("state must be [%d, %d]" % (self.num_envs, self.state_dim))
        check(lib().rmv_set_state(self._h, rows.ctypes.data, 0, _stream(stream, self.device)))

    def get_episode_words(self, stream=None):
        out = np.empty(self.num_envs, dtype=np.uint32)
        check(lib().rmv_get_episode(self._h, out.ctypes.data, 0, _stream(stream, self.device)))
        return out

    def set_episode_words(self, words, stream=None):
        words = np.ascontiguousarray(words, dtype=np.uint32)
        check(lib().rmv_set_episode(self._h, words.ctypes.data, 0, _stream(stream, self.device)))

    def state_view(self):
        """Device view of the handle's state.  Tiles: float32 [ld/128, S, 128], env e is [e // 128, :, e % 128]
        (lanes of the last tile beyond N are padding).  Rows: float32 [ld, S], env e is row e (rows beyond N are
        padding) — the observation matrix itself."""
        sp, ep = ctypes.c_void_p(), ctypes.c_void_p()
        check(lib().rmv_state_ptr(self._h, ctypes.byref(sp), ctypes.byref(ep)))
        if self.state_layout == "rows":
            return _DeviceView(sp.value, (self.ld, self.state_dim), "<f4", self)
        return _DeviceView(sp.value, (self.ld // 128, self.state_dim, 128), "<f4", self)

    def episode_stats(self, stream=None):
        out = (ctypes.c_double * 4)()
        check(lib().rmv_episode_stats(self._h, out, _stream(stream, self.device)))
        return {"episodes": out[0], "return_sum": out[1], "length_sum": out[2], "env_steps": out[3]}

    def gather_episodes(self, env_indices, stream=None):
        """(returns float32[k], lengths uint32[k]) of the most recently finished episode of the listed envs."""
        idx = np.ascontiguousarray(env_indices, dtype=np.int64)
        ret = np.empty(idx.shape[0], dtype=np.float32)
        length = np.empty(idx.shape[0], dtype=np.uint32)
        check(lib().rmv_gather_episodes(self._h, idx.ctypes.data, idx.shape[0], ret.ctypes.data, length.ctypes.data,
                                        _stream(stream, self.device)))
        return ret, length

    def last_episode_views(self):
        rp, lp = ctypes.c_void_p(), ctypes.c_void_p()
        check(lib().rmv_episode_ptr(self._h, ctypes.byref(rp), ctypes.byref(lp)))
        return (_DeviceView(rp.value, (self.ld,), "<f4", self), _DeviceView(lp.value, (self.ld,), "<u4", self))


def fill_uniform_actions(actions, num_envs, action_dim, low, high, seed, env_id_offset=0, step_index=0,
                         layout=ACT_AOS, act_ld=0, device=0, stream=None):
    """Random-policy actions ~ U[low, high) written on the device (Philox, keyed by global env id)."""
    _need(actions, int(action_dim) * (int(act_ld) if layout == ACT_SOA and act_ld else int(num_envs)), "action buffer")
    check(lib().rmv_fill_uniform_actions(_ptr(actions, "float32", int(device), "action buffer"), int(num_envs), int(action_dim), layout, int(act_ld),
                                         float(low), float(high), int(seed) & 0xFFFFFFFFFFFFFFFF, int(env_id_offset),
                                         int(step_index), int(device), _stream(stream, device)))
