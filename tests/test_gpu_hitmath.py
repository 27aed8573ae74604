"""The S1 hit loop's bit-built -ln(u) and sincos against libm (numpy, float64 + longdouble)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_neg_log_and_sincos_accuracy(ctx):
    rng = np.random.default_rng(12345)
    n = 400000
    bits = rng.integers(0, 2 ** 32, size=(n, 3), dtype=np.uint64).astype(np.uint32)
    # edge cases: all-zero / all-one words, many leading zeros, cell boundaries of the log table
    edge = [(0, 0, 0), (0, 1, 0), (0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF), (0x80000000, 0, 0x40000000),
            (0x7FFFFFFF, 0xFFFFFFFF, 0x3FFFFFFF), (1, 0, 0x20000000), (0, 0x80000000, 0x1FFFFFFF),
            (0x84000000, 0, 0x80000000), (0x83FFFFFF, 0xFFFFFFFF, 0xC0000000), (0xFFFFFFFF, 0, 0xBFFFFFFF)]
    for k in range(40):
        edge.append((0xFFFFFFFF >> k if k < 32 else 0, 0xFFFFFFFF >> max(k - 32, 0), (k * 0x08000000) & 0xFFFFFFFF))
    bits[:len(edge)] = np.array(edge, dtype=np.uint64).astype(np.uint32)
    out = ctx.debug_hitmath(bits)
    x = (bits[:, 0].astype(np.uint64) << np.uint64(32)) | bits[:, 1].astype(np.uint64)
    x = x | np.uint64(1)  # the kernel forces the lowest bit (u > 0; a 2^-64 shift of u)
    # reference in extended precision; the kernel truncates x to its top 53 bits
    xl = x.astype(np.longdouble)
    E = -(np.log(xl) - 64 * np.log(np.longdouble(2)))
    err = np.abs(out[:, 0] - E.astype(np.float64))
    tol = 4e-16 * np.maximum(E.astype(np.float64), 1.0) + 2.5e-16  # 53-bit truncation of u: <= 2^-53 relative in u
    assert np.all(err <= tol), (err.max(), bits[err.argmax()])
    ang = (bits[:, 2].astype(np.longdouble) + 0.5) * (2 * np.pi / np.longdouble(2) ** 32)
    assert np.abs(out[:, 1] - np.sin(ang).astype(np.float64)).max() < 4e-16
    assert np.abs(out[:, 2] - np.cos(ang).astype(np.float64)).max() < 4e-16
    # the pair is a unit vector
    assert np.abs(out[:, 1] ** 2 + out[:, 2] ** 2 - 1).max() < 1e-15
