"""CPU run of the accounted float32-vs-float64 comparisons (tests/accounting.py) with the float32 MODEL as the
stepper.  The CUDA path equals that model bit for bit (tests/test_gpu_parity.py), so the counts asserted here are
the counts the GPU suite sees at the same (env, n, steps, seed)."""
import pytest

from tests import accounting
from tests.util import load_golden

ENV_IDS = ["quadrotor2d-v0", "quadrotor2d-slungload-v0", "quadrotor3d-v0", "quadrotor3d-slungload-v0"]


def check_counts(env_id, c):
    """Cumulative bounds on the admissible events of a free run (shared with the GPU suite)."""
    steps = c["env_steps"]
    assert c["dones"] > 0.005 * steps                              # episodes ended and were re-drawn on both sides
    assert c["done_flips"] <= max(2, steps // 2_000_000), c        # threshold within 1e-5 of the reference's norm
    assert c["worst"] <= accounting.TOL, c
    if "slung" in env_id:
        assert c["taut"] > 0.1 * steps, c                          # both branches are exercised
        # a state on the switching surface takes either branch; about half of them do, never more
        assert c["branch_flips"] <= 0.6 * c["near_surface"], c
        assert c["drift_flips"] <= max(2, steps // 1_000_000), c   # off the surface but inside the state tolerance
    else:
        assert c["near_surface"] == c["branch_flips"] == c["drift_flips"] == 0, c


@pytest.mark.parametrize("env_id", ENV_IDS)
def test_model_free_running_vs_float64_accounted(env_id):
    c = accounting.free_running(accounting.ModelStepper, env_id, n=1024, steps=300, seed=3)
    check_counts(env_id, c)


@pytest.mark.parametrize("env_id", ["quadrotor2d-slungload-v0", "quadrotor3d-slungload-v0"])
def test_model_follows_slung_golden_trajectories(env_id):
    compared, flips = accounting.golden_trajectory(accounting.ModelStepper, env_id, load_golden(env_id))
    assert compared == 2400 and flips < 0.25 * compared
