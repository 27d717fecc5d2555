"""pyquaternion is a third-party dependency of the reference (requirements.txt:1, `pyquaternion>0.9`), absent from
/root/reference and from this image; oracle/refshim.py restates the subset the envs call (quadrotor3d.py:94-102,
126-180).  These tests pin that restatement to an INDEPENDENT implementation (scipy.spatial.transform.Rotation) and to
the Hamilton algebra it must satisfy, so the 3D golden fixtures do not rest on one hand-written class alone."""
import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st
from scipy.spatial.transform import Rotation

from oracle.refshim import Quaternion

comp = st.floats(min_value=-2.0, max_value=2.0, allow_nan=False, allow_infinity=False)
quat = st.tuples(comp, comp, comp, comp).filter(lambda q: sum(x * x for x in q) > 1e-3)
rate = st.tuples(*[st.floats(min_value=-20.0, max_value=20.0, allow_nan=False)] * 3)


def hamilton(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([aw * bw - ax * bx - ay * by - az * bz,
                     aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw])


@settings(max_examples=300, deadline=None)
@given(quat)
def test_rotation_matrix_matches_scipy_and_normalises_in_place(q):
    obj = Quaternion(q)
    assert np.array_equal(obj.elements, np.array(q, dtype=float))           # the constructor does not normalise
    R = obj.rotation_matrix
    unit = np.array(q, dtype=float) / np.linalg.norm(q)
    assert np.allclose(obj.elements, unit, rtol=0, atol=1e-15)                # ... rotation_matrix does, in place
    want = Rotation.from_quat([unit[1], unit[2], unit[3], unit[0]]).as_matrix()    # scipy is scalar-last
    assert np.allclose(R, want, rtol=0, atol=1e-12)
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and abs(np.linalg.det(R) - 1.0) < 1e-12
    # the column the envs use: R e3 = [2(xz+wy), 2(yz-wx), w^2-x^2-y^2+z^2] of the unit quaternion
    w, x, y, z = unit
    assert np.allclose(R @ [0, 0, 1], [2 * (x * z + w * y), 2 * (y * z - w * x), w * w - x * x - y * y + z * z], atol=1e-12)


@settings(max_examples=300, deadline=None)
@given(quat, rate)
def test_derivative_is_half_q_times_pure_rate(q, w):
    obj = Quaternion(q)
    d = obj.derivative(w).elements
    assert np.allclose(d, 0.5 * hamilton(q, (0.0,) + tuple(w)), rtol=1e-13, atol=1e-13)
    assert abs(np.dot(d, q)) <= 1e-11 * (1 + np.linalg.norm(q) ** 2 * np.linalg.norm(w))     # tangent to the sphere |q| = const
    # the envs call rotation_matrix FIRST on the same object (quadrotor3d.py:96 then :101): derivative then sees q-hat
    _ = obj.rotation_matrix
    unit = np.array(q, dtype=float) / np.linalg.norm(q)
    assert np.allclose(obj.derivative(w).elements, 0.5 * hamilton(unit, (0.0,) + tuple(w)), rtol=1e-12, atol=1e-12)


@settings(max_examples=300, deadline=None)
@given(quat, quat)
def test_product_and_conjugate(a, b):
    p = (Quaternion(a) * Quaternion(b)).elements
    assert np.allclose(p, hamilton(a, b), rtol=1e-13, atol=1e-13)
    assert np.isclose(np.linalg.norm(p), np.linalg.norm(a) * np.linalg.norm(b), rtol=1e-12)
    c = Quaternion(a).conjugate
    assert np.array_equal(c.elements, np.array(a, dtype=float) * [1, -1, -1, -1])     # no normalisation
    assert np.allclose((Quaternion(a) * c).elements, [np.dot(a, a), 0, 0, 0], atol=1e-12)


@settings(max_examples=300, deadline=None)
@given(st.integers(min_value=0, max_value=2 ** 31 - 1))
def test_from_matrix_inverts_rotation_matrix(seed):
    R = Rotation.random(random_state=seed).as_matrix()
    q = Quaternion(matrix=R)
    assert np.isclose(np.linalg.norm(q.elements), 1.0, atol=1e-12)
    assert np.allclose(q.rotation_matrix, R, atol=1e-12)
    xyzw = Rotation.from_matrix(R).as_quat()
    ref = np.array([xyzw[3], xyzw[0], xyzw[1], xyzw[2]])
    assert np.allclose(q.elements, ref, atol=1e-9) or np.allclose(q.elements, -ref, atol=1e-9)
    assert np.max(q.elements) >= 0.5 - 1e-9               # the pivot t >= 1 of the chosen branch: a POSITIVE component >= 1/2

def test_from_matrix_rejects_what_pyquaternion_rejects():
    with pytest.raises(ValueError):
        Quaternion(matrix=np.eye(3) * 1.1)                  # not orthogonal
    with pytest.raises(ValueError):
        Quaternion(matrix=np.diag([1.0, 1.0, -1.0]))        # orthogonal, determinant -1
