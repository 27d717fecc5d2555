"""The C-ABI library loads, exports every symbol include/rmv.h declares, and its device-free
queries answer correctly.  No compute call is made here (CPU only)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "rmv.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rmv_[a-z_0-9]+)\s*\(", text)))


@pytest.fixture(scope="module")
def capi():
    from gym_reinmav_b200 import _capi
    _capi.lib()
    return _capi


def test_header_and_binding_agree(capi):
    declared = _declared()
    assert len(declared) >= 25
    assert sorted(capi.SIGNATURES) == declared


def test_library_exports_every_declared_symbol(capi):
    raw = ctypes.CDLL(capi.LIB_PATH)
    for name in _declared():
        assert hasattr(raw, name), name


def test_device_free_queries(capi):
    lib = capi.lib()
    assert lib.rmv_abi_version() == 2
    assert [capi.kind_dims(k) for k in range(4)] == [(5, 2), (9, 2), (10, 4), (16, 4)]
    assert capi.action_bounds(capi.QUAD3D) == (0.0, 10.0)
    assert capi.action_bounds(capi.QUAD2D) == (-10.0, 10.0)
    assert lib.rmv_padded_ld(1) == 128 and lib.rmv_padded_ld(128) == 128 and lib.rmv_padded_ld(129) == 256
    assert lib.rmv_padded_ld(1 << 20) == 1 << 20


def test_default_params_are_the_reference_constants(capi):
    from oracle.batch import default_params
    for kind in range(4):
        got = capi.default_params(kind).as_dict()
        want = default_params(kind)
        for k, v in want.items():
            assert got[k] == np.float32(v) if isinstance(v, float) else got[k] == v, (kind, k)


def test_errors_are_reported_not_raised_in_c(capi):
    lib = capi.lib()
    assert lib.rmv_kind_dims(7, None, None) == -1
    assert b"unknown env kind" in lib.rmv_last_error()
    with pytest.raises(capi.RmvError):
        capi.kind_dims(-1)


def test_sparse_infos():
    from gym_reinmav_b200.vec_env import SparseInfos
    infos = SparseInfos(5, {3: {"episode": {"r": 1.5, "l": 7}}})
    assert len(infos) == 5 and infos[0] == {} and infos[3]["episode"]["l"] == 7 and infos[-2]["episode"]["r"] == 1.5
    assert [i.get("episode") for i in infos].count(None) == 4
    with pytest.raises(IndexError):
        infos[5]
    # the lazy form the VecEnv builds: arrays of finished envs, dicts only for what is read
    lazy = SparseInfos(8, idx=np.array([1, 6]), returns=np.array([-2.5, 4.0], np.float32), lengths=np.array([9, 3], np.uint32), t=0.25)
    assert len(lazy) == 8 and lazy[0] == {} and lazy[6] == {"episode": {"r": 4.0, "l": 3, "t": 0.25}} and lazy[-7]["episode"]["l"] == 9
    assert [bool(i) for i in lazy] == [False, True, False, False, False, False, True, False]
    assert lazy.episodes() == {1: {"r": -2.5, "l": 9, "t": 0.25}, 6: {"r": 4.0, "l": 3, "t": 0.25}}
    assert list(lazy.finished_indices()) == [1, 6] and lazy[1:3] == [lazy[1], {}]
    assert len(SparseInfos(3)) == 3 and SparseInfos(3)[2] == {} and list(SparseInfos(2)) == [{}, {}]


def test_undersized_buffers_are_refused_before_the_abi():
    """The C-ABI takes plain pointers; the Python binding is the only place an undersized tensor can be caught."""
    import numpy as np
    import pytest
    from gym_reinmav_b200.batch import _need
    _need(None, 10, "x")
    _need(12345, 10, "x")                                  # raw address: cannot be checked
    _need(np.empty(10, np.float32), 10, "x")
    with pytest.raises(ValueError, match="actions holds 8 elements, 16 needed"):
        _need(np.empty((4, 2), np.float32), 16, "actions")


def test_wrong_dtype_or_side_of_the_bus_is_refused_before_the_abi():
    """A float64 / int buffer of the right element count would be reinterpreted as float32, and a host address
    dereferenced by a kernel kills the CUDA context: the binding refuses both."""
    import torch
    from gym_reinmav_b200.batch import HOST, _ptr
    f32, f64 = np.zeros((4, 2), np.float32), np.zeros((4, 2), np.float64)
    assert _ptr(f32, "float32", HOST, "actions") == f32.ctypes.data
    assert _ptr(None, "float32", 0) is None and _ptr(12345, "float32", 0) == 12345
    with pytest.raises(TypeError, match="actions must be float32, got float64"):
        _ptr(f64, "float32", HOST, "actions")
    with pytest.raises(TypeError, match="done must be uint8"):
        _ptr(np.zeros(4, bool), "uint8", HOST, "done")
    with pytest.raises(ValueError, match="device buffer on cuda:0"):
        _ptr(f32, "float32", 0, "actions")
    t = torch.zeros((4, 2), dtype=torch.float32)
    assert _ptr(t, "float32", HOST, "actions") == t.data_ptr()
    with pytest.raises(ValueError, match="must be on cuda:0, got a tensor on cpu"):
        _ptr(t, "float32", 0, "actions")
    with pytest.raises(TypeError, match="actions must be float32, got torch.float64"):
        _ptr(t.double(), "float32", HOST, "actions")
    with pytest.raises(ValueError, match="contiguous"):
        _ptr(t.t(), "float32", HOST, "actions")


def test_auto_launch_shape_policy(capi):
    """What "auto" resolves to (profiles/r1b_sweep_block_prefetch.txt): CTA 256 for quadrotor2d while one
    step's working set stays in L2, 128 otherwise; the episode word is fetched with the state for quadrotor2d always,
    for the slung-load kinds only when the working set streams from HBM, never for quadrotor3d."""
    import ctypes
    import os
    if os.environ.get("RMV_BLOCK") or os.environ.get("RMV_PREFETCH"):
        pytest.skip("developer overrides set")
    lib = capi.lib()

    def shape(kind, n):
        b, p = ctypes.c_int(), ctypes.c_int()
        assert lib.rmv_launch_shape(kind, n, ctypes.byref(b), ctypes.byref(p)) == 0
        return b.value, p.value
    small, big = 1 << 20, 1 << 23
    assert shape(capi.QUAD2D, small) == (256, 1) and shape(capi.QUAD2D, big) == (128, 1)
    assert shape(capi.QUAD2D_SLUNG, small) == (128, 0) and shape(capi.QUAD2D_SLUNG, big) == (128, 1)
    assert shape(capi.QUAD3D, small) == (128, 0) and shape(capi.QUAD3D, big) == (128, 0)
    assert shape(capi.QUAD3D_SLUNG, small) == (128, 1) and shape(capi.QUAD3D_SLUNG, big) == (128, 1)
    assert shape(capi.QUAD3D_SLUNG, 1 << 16) == (128, 0)
    assert lib.rmv_launch_shape(capi.REINMAV, small, None, None) != 0


def test_step_stream_shape_and_policy(capi):
    """Device-free queries of the persistent step stream (DESIGN.md §3 "The step stream"): every warp owns the same number
    of 32-env units and all of them are co-resident; small batches use one-warp CTAs; batches whose step streams several
    hundred MB through HBM keep one launch per step."""
    import ctypes
    import os
    lib = capi.lib()

    def shape(n, sms=148, resident=64):
        k, w, g, b = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        assert lib.rmv_stream_shape(n, sms, resident, ctypes.byref(k), ctypes.byref(w), ctypes.byref(g), ctypes.byref(b)) == 0
        return k.value, w.value, g.value, b.value
    assert shape(1 << 20) == (4, 8192, 2048, 128)            # 8192 warps x 4 units on 148 x 64 slots
    assert shape(1 << 16) == (1, 2048, 2048, 32)             # one unit per warp, 32-thread CTAs spread over the SMs
    assert shape(1 << 23) == (28, 9364, 2341, 128)
    k, w, g, b = shape(17)
    assert (k, w) == (1, 4) and g * b >= w * 32
    for n in (1, 31, 33, 4095, 70001, (1 << 19) + 77):
        k, w, g, b = shape(n)
        units = (n + 31) // 32
        assert w % 4 == 0 and k * w >= units and (k - 1) * w < units + 4 and g * b >= w * 32 and w <= 148 * 64 + 3
    assert lib.rmv_stream_shape(0, 148, 64, None, None, None, None) != 0
    if os.environ.get("RMV_NO_STREAM"):
        pytest.skip("developer override set")
    assert lib.rmv_stream_pays(capi.QUAD3D, 1 << 20) == 1 and lib.rmv_stream_pays(capi.QUAD3D, 1 << 21) == 1
    assert lib.rmv_stream_pays(capi.QUAD3D, 1 << 22) == 0 and lib.rmv_stream_pays(capi.QUAD3D, 1 << 23) == 0
    assert lib.rmv_stream_pays(capi.QUAD2D_SLUNG, 1 << 18) == 1 and lib.rmv_stream_pays(capi.QUAD2D_SLUNG, 1 << 20) == 0
    assert lib.rmv_stream_pays(capi.QUAD2D, 1 << 22) == 1 and lib.rmv_stream_pays(capi.QUAD3D_SLUNG, 1 << 20) == 1
    assert lib.rmv_stream_pays(capi.REINMAV, 1 << 10) == 0
