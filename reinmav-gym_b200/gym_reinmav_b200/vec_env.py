"""QuadVecEnv — N quadrotor envs behind the baselines ``VecEnv`` protocol.

The reference vectorises by process fan-out: ``make_vec_env(env_id, 'native', num_env, ...)``
(/root/reference/gym_reinmav/run.py:89) builds a ``DummyVecEnv``/``SubprocVecEnv`` of N Python
env objects behind ``Monitor`` wrappers, and ``ppo2``'s Runner drives it with
``obs, rews, dones, infos = env.step(actions)`` (run.py:63-68, --play loop :192-211).  This class
exposes the same surface (``num_envs``, ``observation_space``, ``action_space``, ``reset``,
``step_async``/``step_wait``/``step``, ``close``, auto-reset on done with the reset observation
returned, ``infos[i]['episode'] = {'r','l'}`` at episode end) over ONE fused CUDA kernel per step.

Two ways to step:

* ``step(actions_numpy)`` — host buffers in, host buffers out (pinned staging, H2D + kernel +
  D2H inside ``rmv_step_host``); what an unmodified baselines/stable-baselines learner calls.
* ``step_tensor(actions_cuda_tensor)`` — everything stays on the device; returns torch tensors
  that alias internal buffers (valid until the next call).
"""
import ctypes
import time
from collections.abc import Sequence

import numpy as np

from . import _capi
from ._capi import check, lib
from .batch import ENV_IDS, QuadBatch
from .spaces import Box

# action_space / observation_space as the reference declares them
# (quadrotor2d.py:62-65, quadrotor2d_slungload.py:68-71, quadrotor3d.py:70-71, quadrotor3d_slungload.py:75-78)
_ACTION_BOX = {_capi.QUAD2D: (-10.0, 10.0), _capi.QUAD2D_SLUNG: (-10.0, 10.0),
               _capi.QUAD3D: (0.0, 10.0), _capi.QUAD3D_SLUNG: (-10.0, 10.0)}


def make_spaces(kind):
    s, a = _capi.kind_dims(kind)
    if kind == _capi.REINMAV:       # the reference leaves both spaces undefined (reinmav_env.py:65-70); no action input
        return Box(low=-np.inf, high=np.inf, shape=(s,), dtype=np.float32), None
    lo, hi = _ACTION_BOX[kind]
    # quadrotor3d.py:70-71 declares both spaces with dtype=np.float (float64); the other three kinds say float32.
    # (The arrays step()/reset() return stay float32 — the kernel's arithmetic type; float32 is contained in a
    # float64 Box.)
    dt = np.float64 if kind == _capi.QUAD3D else np.float32
    return (Box(low=-10.0, high=10.0, shape=(s,), dtype=dt),
            Box(low=lo, high=hi, shape=(a,), dtype=dt))


class _Pinned:
    """Page-locked host array (cudaHostAlloc through the C-ABI) exposed as a NumPy array."""

    def __init__(self, shape, dtype):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._p = ctypes.c_void_p()
        check(lib().rmv_host_alloc(ctypes.byref(self._p), self.nbytes))
        raw = (ctypes.c_char * max(self.nbytes, 1)).from_address(self._p.value)
        self.array = np.frombuffer(raw, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self._p:
            self.array = None
            lib().rmv_host_free(self._p)
            self._p = ctypes.c_void_p()


class SparseInfos(Sequence):
    """``infos`` without building N dicts per step: ``infos[i]`` is ``{}`` unless env i just
    finished an episode, in which case it is ``{'episode': {'r': return, 'l': length, 't': seconds}}``
    (baselines Monitor's record; ppo2's Runner collects it with ``info.get('episode')``).

    The finished envs are kept as arrays (sorted env indices, returns, lengths); a dict is only built for an entry
    somebody reads — at N = 2^20 about 13 000 envs finish per step, and building their dicts eagerly would cost ten times
    the step."""

    _EMPTY = {}

    def __init__(self, n, finished=None, idx=None, returns=None, lengths=None, t=0.0):
        self._n = n
        if idx is None:                                   # a ready-made {env index: info} mapping
            fin = dict(finished or {})
            idx = np.array(sorted(fin), dtype=np.int64)
            self._dicts = {int(i): fin[int(i)] for i in idx}
            self._r = self._l = None
        else:
            self._dicts, self._r, self._l = {}, returns, lengths
        self._idx, self._t = idx, t

    def __len__(self):
        return self._n

    def _info(self, k):
        i = int(self._idx[k])
        d = self._dicts.get(i)
        if d is None:
            d = self._dicts[i] = {"episode": {"r": float(self._r[k]), "l": int(self._l[k]), "t": self._t}}
        return d

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(self._n))]
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError(i)
        k = int(np.searchsorted(self._idx, i))
        if k < len(self._idx) and self._idx[k] == i:
            return self._info(k)
        return self._EMPTY

    def __iter__(self):
        """All N entries in order (what ``for info in infos`` of ppo2's Runner does) without N binary searches."""
        nxt = 0
        for k in range(len(self._idx)):
            i = int(self._idx[k])
            for _ in range(nxt, i):
                yield self._EMPTY
            yield self._info(k)
            nxt = i + 1
        for _ in range(nxt, self._n):
            yield self._EMPTY

    def finished_indices(self):
        """Sorted indices of the envs that finished an episode in this step."""
        return self._idx

    def episodes(self):
        """Only the finished ones: {env index: {'r':..., 'l':..., 't':...}}."""
        return {int(self._idx[k]): self._info(k)["episode"] for k in range(len(self._idx))}


class QuadVecEnv:
    metadata = {"render.modes": ["human"]}
    reward_range = (-float("inf"), float("inf"))
    spec = None

    def __init__(self, env_id, num_envs, device=0, seed=0, env_id_offset=0, params=None, monitor=True,
                 return_obs=True, copy_outputs=True, state_layout=None):
        """``state_layout``: "rows" keeps the state as the [N, S] observation matrix and steps it in place — no
        observation copy, ``step_tensor`` returns a view of the state itself (15 vs 19 us per quadrotor3d step at
        N = 2^20); "tiles" keeps the tiled struct-of-arrays and writes observation rows beside it.  ``None`` picks rows
        for the 3D kinds and tiles for the planar ones (where the row transposes cost what the copy saves).

        ``copy_outputs=False`` returns views of the pinned staging buffers (valid until the next
        ``step``/``reset``) instead of fresh arrays — at N = 2^20 the three NumPy copies cost more than the
        PCIe transfers.  ``action_buffer`` is the pinned input buffer: filling it in place and passing it to
        ``step`` skips the input copy as well."""
        self.env_id = env_id
        kind = ENV_IDS.get(env_id, env_id) if isinstance(env_id, str) else env_id
        if state_layout is None:
            state_layout = "rows" if kind in (_capi.QUAD3D, _capi.QUAD3D_SLUNG) else "tiles"
        self.batch = QuadBatch(env_id, num_envs, device=device, seed=seed, env_id_offset=env_id_offset,
                               params=params, track_episodes=monitor, state_layout=state_layout)
        self.state_layout = self.batch.state_layout
        self.num_envs = self.batch.num_envs
        self.device = device
        self.monitor = monitor
        self.return_obs = return_obs
        self.copy_outputs = copy_outputs
        self.observation_space, self.action_space = make_spaces(self.batch.kind)
        n, S, A = self.num_envs, self.batch.state_dim, self.batch.action_dim
        self._act = _Pinned((n, A), np.float32)
        self._obs = _Pinned((n, S), np.float32)
        self._rew = _Pinned((n,), np.float32)
        self._done = _Pinned((n,), np.uint8)
        self._pending = None
        self._t = None            # torch buffers for step_tensor, made on first use
        self._tstart = time.time()
        self.closed = False
        self.viewer = None

    # ------------------------------------------------------------------ VecEnv protocol (host)
    def seed(self, seed=None):
        if seed is not None:
            self.batch.set_seed(seed)
        return [seed]

    @property
    def action_buffer(self):
        """The pinned [N, A] float32 input buffer ``step`` uploads from."""
        return self._act.array

    def _out(self, a):
        return a.copy() if self.copy_outputs else a

    def reset(self):
        self.batch.reset_host(self._obs.array)
        return self._out(self._obs.array)

    def step_async(self, actions):
        if actions is not self._act.array:
            a = np.asarray(actions, dtype=np.float32)
            if a.shape != self._act.array.shape:
                raise ValueError("actions must have shape %s, got %s" % (self._act.array.shape, a.shape))
            self._act.array[...] = a
        self._pending = True

    def step_wait(self):
        if not self._pending:
            raise RuntimeError("step_wait() without step_async()")
        self._pending = None
        self.batch.step_host(self._act.array, self._rew.array, self._done.array,
                             self._obs.array if self.return_obs else None, auto_reset=True)
        done = self._done.array.astype(bool)
        infos = self._finished(done) if self.monitor else SparseInfos(self.num_envs)
        obs = self._out(self._obs.array) if self.return_obs else None
        return obs, self._out(self._rew.array), done, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def _finished(self, done):
        """SparseInfos of this step: Monitor's episode record for the envs whose done flag is set."""
        idx = np.flatnonzero(done)
        if idx.size == 0:
            return SparseInfos(self.num_envs)
        r, l = self.batch.gather_episodes(idx)
        t = round(time.time() - self._tstart, 6)          # Monitor: seconds since the env was created
        return SparseInfos(self.num_envs, idx=idx, returns=r, lengths=l, t=t)

    # ------------------------------------------------------------------ device-resident stepping
    def _tensors(self):
        if self._t is None:
            import torch
            dev = "cuda:%d" % self.device
            n, S = self.num_envs, self.batch.state_dim
            state = torch.as_tensor(self.batch.state_view(), device=dev)
            # row-state handle: the observation matrix IS the state buffer (obs == state, quadrotor3d.py:104)
            obs = state[:n] if self.state_layout == "rows" else torch.empty((n, S), dtype=torch.float32, device=dev)
            self._t = dict(obs=obs, state=state,
                           rew=torch.empty((n,), dtype=torch.float32, device=dev),
                           done=torch.empty((n,), dtype=torch.uint8, device=dev))
            for k in ("obs", "rew", "done"):
                self._t[k + "_ptr"] = self._t[k].data_ptr()
        return self._t

    def reset_tensor(self):
        t = self._tensors()
        self.batch.reset(obs=None if self.state_layout == "rows" else t["obs"])
        return t["obs"]

    def step_tensor(self, actions, want_obs=True):
        """actions: float32 CUDA tensor [N, A] (row per env).  Returns (obs [N,S] or None,
        reward [N], done [N] uint8) as tensors aliasing internal buffers."""
        t = self._tensors()
        rows = self.state_layout == "rows"          # the step updates t["obs"] (= the state) in place: nothing to copy
        # the internal buffers go down as the raw addresses validated when they were made: only `actions` is checked per call
        self.batch.step(actions, t["rew_ptr"], t["done_ptr"], obs=t["obs_ptr"] if (want_obs and not rows) else None, auto_reset=True)
        return (t["obs"] if want_obs else None), t["rew"], t["done"]

    def step_many_tensor(self, actions, rewards=None, dones=None, want_obs=True):
        """T steps with auto-reset from a pre-generated action block (replays, random-policy runs, shooting): ``actions``
        float32 CUDA [T, N, A]; ``rewards`` float32 / ``dones`` uint8 [T, N] are filled (allocated if None).  ONE launch of
        the step stream where the library finds it pays, else T launches; the same results as T ``step_tensor`` calls, bit
        for bit.  Returns (obs after the last step [N, S] or None, rewards, dones); Monitor statistics keep counting."""
        import torch
        T = int(actions.shape[0])
        if actions.dim() != 3 or tuple(actions.shape[1:]) != (self.num_envs, self.batch.action_dim):
            raise ValueError("actions must have shape [T, %d, %d], got %s" % (self.num_envs, self.batch.action_dim, tuple(actions.shape)))
        dev = "cuda:%d" % self.device
        if rewards is None:
            rewards = torch.empty((T, self.num_envs), dtype=torch.float32, device=dev)
        if dones is None:
            dones = torch.empty((T, self.num_envs), dtype=torch.uint8, device=dev)
        t = self._tensors()
        self.batch.step_many(T, actions, rewards, dones, out_ld=self.num_envs, layout=0, fused=0)
        if want_obs and self.state_layout != "rows":
            self.batch.get_state_device(t["obs"])       # tiled state -> observation rows (a row-state handle IS the rows)
        return (t["obs"] if want_obs else None), rewards, dones

    def control_tensor(self, ref_pos=None, ref_vel=None):
        """Actions [N, A] of the reference's geometric controller for the current states (device tensor)."""
        import torch
        t = self._tensors()
        if "act" not in t:
            t["act"] = torch.empty((self.num_envs, self.batch.action_dim), dtype=torch.float32,
                                   device="cuda:%d" % self.device)
        self.batch.control(t["act"], ref_pos=ref_pos, ref_vel=ref_vel)
        return t["act"]

    def state_tensor(self):
        """The handle's state buffer itself (a view; obs == state in these envs): tiled [N_pad/128, S, 128], env e is
        ``[e // 128, :, e % 128]`` — or, with ``state_layout='rows'``, [N_pad, S]."""
        return self._tensors()["state"]

    def episode_stats(self):
        return self.batch.episode_stats()

    # ------------------------------------------------------------------ misc protocol
    def render(self, mode="human"):
        return None           # the reference's viewers (pyglet / vpython) need a display; not part of this path

    def get_images(self):
        return []

    @property
    def unwrapped(self):
        return self

    def close(self):
        if self.closed:
            return
        self.closed = True
        self._t = None
        for b in (self._act, self._obs, self._rew, self._done):
            b.free()
        self.batch.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
