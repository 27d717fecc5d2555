"""tests/golden/reinmav-v0.npz: the reference's ReinmavEnv (reinmav_env.py) stepped 120 times from its
own initial state and from 5 perturbed states.  Build-container only.  Run: python -m oracle.make_golden_reinmav"""
import os
import sys
import types

import numpy as np

from . import refshim


def load_reinmav_class():
    stubs = refshim._stub_modules()
    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    stubs.update({"matplotlib": mpl, "matplotlib.pyplot": plt})
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        src = open(os.path.join(refshim.NATIVE_DIR, "reinmav_env.py")).read()
        src = src.replace('print("__init__ called")', "pass").replace('print("reset")', "pass")
        ns = {"__name__": "rmv_refshim_reinmav"}
        exec(compile(src, os.path.join(refshim.NATIVE_DIR, "reinmav_env.py"), "exec"), ns)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return ns["ReinmavEnv"]


def main():
    cls = load_reinmav_class()
    rng = np.random.RandomState(7)
    T, E = 120, 6
    init = np.tile(np.array([0.0] * 6 + [1.0] + [0.0] * 6), (E, 1))
    pert = rng.uniform(-1, 1, size=(E - 1, 13)) * np.array([0.02] * 3 + [0.05] * 3 + [0.01] * 4 + [0.1] * 3)
    init[1:] += pert
    init = init.astype(np.float32).astype(np.float64)          # float32-exact inputs
    traj = np.empty((T, E, 13))
    for e in range(E):
        env = cls()
        env.state = init[e].copy()
        for t in range(T):
            s, r, d, _ = env.step()
            assert r == 90.0 and d is True
            traj[t, e] = np.asarray(s, dtype=float).reshape(-1)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reinmav-v0.npz")
    np.savez_compressed(out, init=init, traj=traj)
    print("wrote", out, "final x,y,z of env 0:", traj[-1, 0, :3])


if __name__ == "__main__":
    main()
