"""Pins the tcgen05 SWIZZLE_NONE descriptor semantics the tensor-core planner relies on:
K-major A/B operands, row-block offsets, LBO/SBO roles (see csrc/tc_common.cuh).  (MN-major with
SWIZZLE_NONE returned zeros for kind::tf32 on B200, so the adjoint pass streams its own
pre-transposed K-major weight images instead.)"""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def tf32_exact(rng, shape):
    x = rng.standard_normal(shape).astype(np.float32)
    return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


@pytest.mark.parametrize("NB", [32, 64])
@pytest.mark.parametrize("K", [8, 24, 200])
def test_k_major_a(K, NB, built_lib):
    """A stored [k/4][m][4]: LBO = stride between 16-byte k-chunks, SBO = 128 (8-row groups)."""
    from gan_mpc_b200 import _lib
    rng = np.random.default_rng(K + NB)
    A, B = tf32_exact(rng, (128, K)), tf32_exact(rng, (NB, K))
    a_s1 = 128 * 16                   # one k-chunk slab: 128 rows x 16 B
    b_s1 = NB * 16 + 16               # padded slab (bank-conflict-free epilogue stores)
    a_bytes = (K // 4) * a_s1
    smem = a_bytes + (K // 4) * b_s1 + 256
    D = _lib.tc_probe(A, B, 0, a_lbo=a_s1, a_sbo=128, a_s1=a_s1, a_s2=128, a_kstep=2 * a_s1,
                      b_lbo=b_s1, b_sbo=128, b_s1=b_s1, b_s2=128, a_bytes=a_bytes, smem_bytes=smem)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    assert np.abs(D - ref).max() < 1e-4 * np.abs(ref).max()


def test_second_row_block_and_short_image(built_lib):
    """The planner addresses the second 128-row block by advancing the start address by 16 row
    groups (2048 B) and lets unused rows alias whatever follows: rows 128..199 of a 200-row
    operand must come out right."""
    from gan_mpc_b200 import _lib
    rng = np.random.default_rng(5)
    K, NB, rows = 16, 32, 200
    W, B = tf32_exact(rng, (rows, K)), tf32_exact(rng, (NB, K))
    lbo = rows * 16
    img = np.zeros((K // 4) * lbo // 4 + 2048, np.float32)          # + slack for aliased rows
    for r in range(rows):
        for k in range(K):
            img[((k // 4) * lbo + (r // 8) * 128 + (r % 8) * 16 + (k % 4) * 4) // 4] = W[r, k]
    b_s1 = NB * 16 + 16
    a_bytes = img.size * 4
    # raw-image mode (a_major=2): descriptor start = image + 2048 -> rows 128..255
    D = _lib.tc_probe(img[512:].reshape(1, -1), B, 2, a_lbo=lbo, a_sbo=128, a_s1=0, a_s2=0,
                      a_kstep=2 * lbo, b_lbo=b_s1, b_sbo=128, b_s1=b_s1, b_s2=128,
                      a_bytes=a_bytes - 2048, smem_bytes=a_bytes + (K // 4) * b_s1 + 256)
    ref = W[128:].astype(np.float64) @ B.astype(np.float64).T
    assert np.abs(D[:72] - ref).max() < 1e-4 * np.abs(ref).max()
