"""Build tests/host_twin/libtwin.so (host compile of the kernel arithmetic header).  TEST ONLY."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "libtwin.so")


def build(force=False):
    src = os.path.join(HERE, "twin.cpp")
    csrc = os.path.join(ROOT, "reinmav-gym_b200", "csrc")
    deps = [src] + [os.path.join(csrc, h) for h in ("rmv_math.cuh", "rmv_f64.cuh")]
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(d) for d in deps):
        return OUT
    cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
           "-x", "c++", src, "-I", os.path.join(ROOT, "reinmav-gym_b200", "csrc"),
           "-I", os.path.join(ROOT, "include"), "-o", OUT]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
