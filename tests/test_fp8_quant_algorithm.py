"""CPU check of the arithmetic shortcut inside the activation quantiser kernel (magi-1_b200/csrc/fp8_quant.cu): the
kernel does not call the IEEE division per element; it computes the correctly rounded quotient with a reciprocal and two
fused-multiply-add corrections, takes the clamp from the conversion's saturation and restores the sign of zero from x.
Here that sequence is restated in numpy (fp32 fma emulated in 80-bit arithmetic) and compared, byte for byte, with the
oracle's div_clamp_to (= the reference's eager chain, pinned by tests/golden/fp8_quant.npz) for EVERY bf16 input the
fast path accepts against random, 'nice' and extreme divisors. The kernel itself is checked on the B200 by
tests/test_fp8_gpu.py::test_fp8_quantize_exhaustive_over_bf16_inputs."""
import numpy as np

import magi_oracle as O

F32, LD = np.float32, np.longdouble


def fma(a, b, c):
    return (a.astype(LD) * b.astype(LD) + c.astype(LD)).astype(F32)


def kernel_fast_path(x, s):
    """quant_row8 of csrc/fp8_quant.cu for one divisor s in [2^-60, 2^60] and finite |x| < 2^65."""
    sv, y = np.full_like(x, s), np.full_like(x, F32(1.0) / s)        # y = RN(1 / s)  (__frcp_rn)
    q = (x * y).astype(F32)
    for _ in range(2):
        e = fma(q, -sv, x)
        q = fma(e, y, q)
    b = O.bf16_round(q)
    b = (b.view(np.uint32) | (x.view(np.uint32) & np.uint32(0x80000000))).view(F32)     # sign of x
    return O.e4m3_encode(O.e4m3_round(np.clip(b, -448, 448)))         # cvt.rn.satfinite.e4m3x2


def test_reciprocal_and_two_fma_corrections_reproduce_the_division_bytes():
    x = (np.arange(65536, dtype=np.uint32) << 16).view(F32)
    x = x[np.isfinite(x) & (np.abs(x) < 2.0 ** 65)]
    rng = np.random.default_rng(0)
    scales = np.concatenate([rng.uniform(0.004, 0.02, 24), rng.uniform(0.5, 1.5, 8), 2.0 ** rng.uniform(-59, 59, 16),
                             [1.0, 0.0078125, 3.0, 0.1, 7.0, 448.0, 4 / 448, 0.3, 0.75, 6.0, 2.0 ** -60, 2.0 ** 60]])
    with np.errstate(all="ignore"):
        for s in scales.astype(F32):
            ref = O.e4m3_encode(O.div_clamp_to(x, np.asarray([s], dtype=F32)))
            assert np.array_equal(kernel_fast_path(x, s), ref), float(s)
