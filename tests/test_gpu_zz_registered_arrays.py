"""Page-locking caller-owned arrays (kem_host_register / MembraneModel.register_host_array).
Kept in its own, alphabetically last GPU module: it is the one test whose outcome depends on
the host allocator's page layout."""
import numpy as np
import pytest

from test_gpu_api import Func, make_pair

pytestmark = pytest.mark.gpu


def test_registered_host_arrays_take_the_direct_copy_path(built):
    """register_host_array page-locks the caller's `u.x.array` once; the unmodified setters and
    getters then copy it without staging and give the same values."""
    from knpemi_b200._cabi import host_is_pinned
    n = 40003          # 320 kB per array: above malloc's mmap threshold, so no two arrays share a page
    gpu, cpu, X, rng = make_pair("hh_tissue", n)
    k_e, v, back = Func(3.0 + 0.01 * rng.normal(size=n)), Func(-70.0 + rng.normal(size=n)), Func(np.zeros(n))
    assert not host_is_pinned(k_e.x.array)
    for u in (k_e, v, back, k_e):                       # registering twice is a no-op
        assert gpu.register_host_array(u) is u
    assert host_is_pinned(k_e.x.array) and host_is_pinned(back.x.array)
    for m in (gpu, cpu):
        m.set_parameter('K_e', k_e)
        m.set_membrane_potential(v)
    loc = lambda x: x[1] < 30e-6      # noqa: E731
    for m in (gpu, cpu):
        m.set_parameter('K_e', v, locator=loc)          # masked write from a registered array
    ref = Func(np.zeros(n))
    gpu.get_parameter('K_e', back)
    cpu.get_parameter('K_e', ref)
    assert np.array_equal(back.x.array, ref.x.array)
    gpu.get_membrane_potential(back)
    assert np.array_equal(back.x.array, v.x.array)
    from knpemi_b200._cabi import KemError
    with pytest.raises(KemError):
        gpu.register_host_array(np.zeros(4, dtype=np.float32))
    gpu.close()
    assert not host_is_pinned(k_e.x.array)              # close() unregisters
