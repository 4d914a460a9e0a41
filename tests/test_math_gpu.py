"""GPU unit test of the arithmetic primitives every kernel is built from (csrc/ek_common.cuh), through the C ABI
(ek_selftest_math): the setp/selp forms of max(x,0) / min(x,0) written in PTX (EK_ASM_MAX0) against the C form on
+-0, denormals, +-inf, NaN and random operands -- bit for bit -- and the MUFU-seeded reciprocal / square root /
reciprocal square root / division against IEEE arithmetic (<= 2 ulp; the 1e-12 parity contract has 4 decades of margin)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_max0_min0_select_forms_equal_the_c_form():
    from euler_kokkos_b200 import capi

    rng = np.random.default_rng(3)
    tiny = np.nextafter(0.0, 1.0)
    special = np.array([0.0, -0.0, tiny, -tiny, 2.2250738585072014e-308, -2.2250738585072014e-308, 1e-310, -1e-310,
                        np.inf, -np.inf, np.nan, 1.0, -1.0, 1.7976931348623157e308, -1.7976931348623157e308])
    x = np.concatenate([special, rng.standard_normal(4096) * 10.0 ** rng.integers(-300, 300, 4096)])
    r = capi.selftest_math(x)
    with np.errstate(invalid="ignore"):
        want_max = np.where(x > 0.0, x, 0.0)  # the C form of dmax0 / dmin0: 0 for NaN, +0 for -0
        want_min = np.where(x < 0.0, x, 0.0)
    assert np.array_equal(r["max0"].view(np.uint64), want_max.view(np.uint64))
    assert np.array_equal(r["min0"].view(np.uint64), want_min.view(np.uint64))


def test_fast_division_and_square_roots_are_within_2_ulp():
    from euler_kokkos_b200 import capi

    rng = np.random.default_rng(5)
    x = np.abs(rng.standard_normal(1 << 16)) * 10.0 ** rng.integers(-100, 100, 1 << 16) + 1e-300
    r = capi.selftest_math(x)
    for name, want in (("rcp", 1.0 / x), ("sqrt", np.sqrt(x)), ("rsqrt", 1.0 / np.sqrt(x)), ("div", (1.0 + x) / x)):
        ulp = np.abs(r[name] - want) / np.spacing(np.abs(want))
        assert ulp.max() <= 2.0, (name, ulp.max())
    # sqrt_fast(0) = 0 (zero guard); negative zero too
    z = capi.selftest_math(np.array([0.0, -0.0, 1e-320]))
    assert np.all(z["sqrt"] == 0.0)
