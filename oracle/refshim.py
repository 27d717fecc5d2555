"""T0-ref: run the reference's OWN source under stub third-party modules.

TEST INFRASTRUCTURE, BUILD-CONTAINER ONLY (``/root/reference`` does not exist on
the GPU box).  Nothing is copied: the four env source files are read from
``/root/reference/gym_reinmav/envs/native/`` at call time, patched in memory
(list below) and ``exec``'d.  The only consumer is
``oracle/make_golden.py`` (fixtures for tests/golden) and the container-only
test that re-checks those fixtures.

Why stubs: ``gym`` and ``pyquaternion`` are not installed and there is no
index.  The stubs provide exactly the surface the four files touch:

* ``gym.Env``, ``gym.spaces.Box``, ``gym.utils.seeding.np_random``,
  ``gym.logger.warn``, ``gym.error``
* ``pyquaternion.Quaternion`` — the subset called at quadrotor3d.py:94-102,139,
  169,176 and quadrotor3d_slungload.py:107-145,184,218-226, restated from the
  published behaviour of pyquaternion 0.9.x (pin: ``pyquaternion>0.9``,
  /root/reference/requirements.txt:1).  Third-party, un-vendored: see DESIGN.md.

In-memory patches (never written back):

  P1  quadrotor2d.py:95-98   the ``done`` expression does not parse (missing
      line continuation).  repair 'A' drops the two leftover template lines
      (:97-98); repair 'B' adds the missing continuation to :96.
  P2  quadrotor2d.py:93, quadrotor2d_slungload.py:134   ``att`` -> ``att[0]``
      inside the state tuple (ragged np.array is an error on NumPy >= 1.24).
  P3  quadrotor3d.py:70-71   ``np.float`` -> ``float`` (alias removed in 1.24).
  P4  quadrotor2d.py:136, quadrotor2d_slungload.py:181   controller returns
      ``np.array([thrust, w])`` with 1-element ``w`` -> ``w[0]``.
  P5  ``print("reset")`` -> ``pass``.
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("RMV_REFERENCE_ROOT", "/root/reference")
NATIVE_DIR = os.path.join(REFERENCE_ROOT, "gym_reinmav", "envs", "native")


def available():
    return os.path.isdir(NATIVE_DIR)


# --------------------------------------------------------------------------- pyquaternion subset
class Quaternion:
    """Restated subset of pyquaternion 0.9.x (float64, scalar-first)."""

    def __init__(self, *args, **kwargs):
        if "matrix" in kwargs:
            self.q = self._from_matrix(np.asarray(kwargs["matrix"], dtype=float))
        elif "vector" in kwargs:
            v = np.asarray(kwargs["vector"], dtype=float).reshape(3)
            self.q = np.array([0.0, v[0], v[1], v[2]])
        elif "array" in kwargs:
            self.q = np.asarray(kwargs["array"], dtype=float).reshape(4).copy()
        elif len(args) == 1 and isinstance(args[0], Quaternion):
            self.q = args[0].q.copy()
        elif len(args) == 1 and np.isscalar(args[0]):
            self.q = np.array([float(args[0]), 0.0, 0.0, 0.0])
        elif len(args) == 1:
            self.q = np.asarray(args[0], dtype=float).reshape(4).copy()
        else:
            raise TypeError("unsupported Quaternion constructor in the oracle stub")

    @staticmethod
    def _from_matrix(R, rtol=1e-05, atol=1e-08):
        if R.shape != (3, 3):
            raise ValueError("matrix must be 3x3")
        if not np.allclose(np.dot(R, R.conj().transpose()), np.eye(3), rtol=rtol, atol=atol):
            raise ValueError("Matrix must be orthogonal, i.e. its transpose should be its inverse")
        if not np.isclose(np.linalg.det(R), 1.0, rtol=rtol, atol=atol):
            raise ValueError("Matrix must be special orthogonal i.e. its determinant must be +1.0")
        m = R.conj().transpose()
        if m[2, 2] < 0:
            if m[0, 0] > m[1, 1]:
                t = 1 + m[0, 0] - m[1, 1] - m[2, 2]
                q = [m[1, 2] - m[2, 1], t, m[0, 1] + m[1, 0], m[2, 0] + m[0, 2]]
            else:
                t = 1 - m[0, 0] + m[1, 1] - m[2, 2]
                q = [m[2, 0] - m[0, 2], m[0, 1] + m[1, 0], t, m[1, 2] + m[2, 1]]
        else:
            if m[0, 0] < -m[1, 1]:
                t = 1 - m[0, 0] - m[1, 1] + m[2, 2]
                q = [m[0, 1] - m[1, 0], m[2, 0] + m[0, 2], m[1, 2] + m[2, 1], t]
            else:
                t = 1 + m[0, 0] + m[1, 1] + m[2, 2]
                q = [t, m[1, 2] - m[2, 1], m[2, 0] - m[0, 2], m[0, 1] - m[1, 0]]
        return np.array(q, dtype=float) * (0.5 / np.sqrt(t))

    # -- algebra
    def _q_matrix(self):
        w, x, y, z = self.q
        return np.array([[w, -x, -y, -z], [x, w, -z, y], [y, z, w, -x], [z, -y, x, w]])

    def _q_bar_matrix(self):
        w, x, y, z = self.q
        return np.array([[w, -x, -y, -z], [x, w, z, -y], [y, -z, w, x], [z, y, -x, w]])

    def _sum_of_squares(self):
        return np.dot(self.q, self.q)

    def _normalise(self):
        if not abs(1.0 - self._sum_of_squares()) < 1e-14:
            n = np.sqrt(self._sum_of_squares())
            if n > 0:
                self.q = self.q / n

    def __mul__(self, other):
        if isinstance(other, Quaternion):
            return Quaternion(array=np.dot(self._q_matrix(), other.q))
        return self * Quaternion(other)

    def __rmul__(self, other):
        return Quaternion(other) * self

    @property
    def conjugate(self):
        return Quaternion(array=self.q * np.array([1.0, -1.0, -1.0, -1.0]))

    @property
    def rotation_matrix(self):
        self._normalise()          # mutates: later calls on the same object see the unit quaternion
        product = np.dot(self._q_matrix(), self._q_bar_matrix().conj().transpose())
        return product[1:][:, 1:]

    def derivative(self, rate):
        return 0.5 * self * Quaternion(vector=np.asarray(rate, dtype=float))

    @property
    def elements(self):
        return self.q


# --------------------------------------------------------------------------- gym subset
class _Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)


class _Env:
    metadata = {"render.modes": []}


def _np_random(seed=None):
    rng = np.random.RandomState(None if seed is None else int(seed) % (2 ** 32))
    return rng, seed


def _stub_modules():
    gym = types.ModuleType("gym")
    gym.Env = _Env
    spaces = types.ModuleType("gym.spaces")
    spaces.Box = _Box
    utils = types.ModuleType("gym.utils")
    seeding = types.ModuleType("gym.utils.seeding")
    seeding.np_random = _np_random
    utils.seeding = seeding
    logger = types.ModuleType("gym.logger")
    logger.warn = lambda *a, **k: None
    error = types.ModuleType("gym.error")
    gym.spaces, gym.utils, gym.logger, gym.error = spaces, utils, logger, error
    pyq = types.ModuleType("pyquaternion")
    pyq.Quaternion = Quaternion
    return {"gym": gym, "gym.spaces": spaces, "gym.utils": utils, "gym.utils.seeding": seeding,
            "gym.logger": logger, "gym.error": error, "pyquaternion": pyq}


# --------------------------------------------------------------------------- loader
def _patched_source(fname, repair):
    with open(os.path.join(NATIVE_DIR, fname), "r") as fh:
        lines = fh.read().split("\n")
    src = None
    if fname == "quadrotor2d.py":
        # P1: locate the broken expression by content, not by line number alone
        i = next(k for k, l in enumerate(lines) if "linalg.norm(vel, 2) > 10.0" in l)
        assert not lines[i].rstrip().endswith("\\"), "reference changed: P1 no longer needed?"
        if repair == "A":
            del lines[i + 1:i + 3]
        elif repair == "B":
            lines[i] = lines[i].rstrip() + " \\"
        else:
            raise ValueError("repair must be 'A' or 'B'")
    src = "\n".join(lines)
    if fname in ("quadrotor2d.py", "quadrotor2d_slungload.py"):
        src = src.replace("(pos[0], pos[1], att, vel[0], vel[1]", "(pos[0], pos[1], att[0], vel[0], vel[1]")  # P2
        src = src.replace("np.array([thrust, w])", "np.array([thrust, w[0]])")                                  # P4
    src = src.replace("dtype=np.float,", "dtype=float,")                                                        # P3
    src = src.replace('print("reset")', "pass")                                                                  # P5
    return src


_FILES = {
    "quadrotor2d-v0": ("quadrotor2d.py", "Quadrotor2D"),
    "quadrotor2d-slungload-v0": ("quadrotor2d_slungload.py", "Quadrotor2DSlungload"),
    "quadrotor3d-v0": ("quadrotor3d.py", "Quadrotor3D"),
    "quadrotor3d-slungload-v0": ("quadrotor3d_slungload.py", "Quadrotor3DSlungload"),
}


def load_class(env_id, repair="A"):
    """Return the reference env class for ``env_id`` (exec'd from /root/reference)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    fname, cls = _FILES[env_id]
    stubs = _stub_modules()
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        ns = {"__name__": "rmv_refshim_" + fname[:-3]}
        exec(compile(_patched_source(fname, repair), os.path.join(NATIVE_DIR, fname), "exec"), ns)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return ns[cls]


class ScriptedUniform:
    """Stands in for ``env.np_random``: ``uniform()`` hands back scripted draws.

    Used to make the reference's ``reset()`` consume the same Philox numbers the
    CUDA path uses (the values are exact in float32 and float64 alike)."""

    def __init__(self, draws_fn):
        self._fn = draws_fn

    def uniform(self, low=-1.0, high=1.0, size=None):
        n = int(np.prod(size))
        v = np.asarray(self._fn(n), dtype=np.float64)
        assert v.shape == (n,) and low == -1.0 and high == 1.0
        return v.reshape(size)
