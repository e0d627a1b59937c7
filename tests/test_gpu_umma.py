"""GPU: the hand-written tcgen05 (UMMA) tile engine vs a host fp64 reference, through the C ABI."""
import ctypes as C

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant,k,n", [(0, 32, 64), (0, 64, 64), (0, 64, 16), (0, 128, 32),
                                         (1, 64, 64), (1, 16, 64), (1, 64, 32),
                                         (2, 128, 64), (2, 128, 32),
                                         (3, 128, 16), (4, 16, 64), (4, 16, 32),
                                         (5, 64, 64), (5, 32, 16),
                                         (6, 64, 64), (6, 32, 16), (6, 64, 16), (7, 64, 64), (7, 16, 64)])
def test_umma_tile_gemm(ctx, variant, k, n):
    import dependence_free_rl_b200 as D
    err = C.c_float(-1.0)
    D._lib.check(D._lib.lib.dfrl_umma_selftest(ctx.h, variant, k, n, C.byref(err)))
    # bf16 hi/lo split, three products: ~2^-17 relative per product, fp32 accumulation
    assert 0 <= err.value < 2e-5, (variant, k, n, err.value)
