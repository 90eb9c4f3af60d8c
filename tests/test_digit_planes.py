"""Host-side restatement of the integer digit-plane arithmetic behind the tensor-core moments
(plspm-python_b200/csrc/kernels_digits.cuh): the scaling, the carry-free balanced base-128
digits, the int32 accumulation bound and the fp64 recombination.  Checks the claims the CUDA path relies on:
the digits reproduce the scaled integer exactly, the int8 x int8 -> int32 sums cannot overflow within a GEMM
chunk, and the recombined moments agree with fp64 accumulation to rounding level."""
import numpy as np

DIGITS = 6
OFFSET = sum(64 * 128 ** k for k in range(DIGITS))      # I8_OFFSET
KCHUNK = 262144                                          # I8_KCHUNK


def scale_exponent(col_absmax: np.ndarray) -> np.ndarray:
    """digit_scale_kernel: 2^e > max|x| (frexp), e = 0 for an all-zero column."""
    e = np.zeros(col_absmax.shape, dtype=np.int64)
    nz = col_absmax > 0
    e[nz] = np.frexp(col_absmax[nz])[1]
    return e


def digit_planes(q: np.ndarray) -> np.ndarray:
    """digit_bytes(): digit k of q is ((q + OFFSET) >> 7k & 127) - 64 -- no carries."""
    u = (q + OFFSET).astype(np.uint64)
    return np.stack([((u >> np.uint64(7 * k)) & np.uint64(127)).astype(np.int64) - 64 for k in range(DIGITS)])


def test_offset_constant_and_digit_range():
    assert OFFSET == 64 * ((1 << 42) - 1) // 127 == 2216338399296
    rng = np.random.default_rng(0)
    q = np.concatenate([rng.integers(-(1 << 40), (1 << 40) + 1, size=100000), [-(1 << 40), 1 << 40, 0, -1, 1, 63, 64, -64, -65]])
    d = digit_planes(q)
    assert d.min() >= -64 and d.max() <= 63
    assert ((q + OFFSET) >= 0).all() and ((q + OFFSET) < (1 << 42)).all()
    np.testing.assert_array_equal(sum(d[k] * 128 ** k for k in range(DIGITS)), q)


def test_int32_accumulation_cannot_overflow_within_a_chunk():
    assert KCHUNK * 64 * 127 < 2 ** 31          # |digit| <= 64, multiplicity <= 127 (larger ones take the fp64 route)


def test_column_sums_and_gram_from_planes_match_fp64():
    rng = np.random.default_rng(1)
    N, P = 20000, 6
    X = rng.normal(size=(N, P)) * np.array([1e-3, 1.0, 40.0, 1e4, 0.3, 7.0]) + rng.normal(size=P) * 0.01
    X = X - X.mean(axis=0)                       # the engine stores globally centred columns
    counts = np.bincount(rng.integers(0, N, size=N), minlength=N).astype(np.int64)   # bootstrap multiplicities
    assert counts.max() <= 127
    e = scale_exponent(np.abs(X).max(axis=0))
    # column sums: q = rint(x 2^(40-e)), colsum = 2^(e-40) sum_k 128^k (counts . d_k)
    q = np.rint(X * np.ldexp(1.0, 40 - e)).astype(np.int64)
    assert np.abs(q).max() <= 1 << 40
    d = digit_planes(q)                          # [6, N, P]
    S = np.einsum("i,kip->kp", counts, d)        # exact integer sums (int64 here; int32 on the device)
    assert np.abs(S).max() < 2 ** 31
    colsum = sum(S[k].astype(np.float64) * 128.0 ** k for k in range(DIGITS)) * np.ldexp(1.0, e - 40)
    ref = (counts[:, None] * X).sum(axis=0)
    scale = np.abs(counts[:, None] * X).sum(axis=0)
    assert (np.abs(colsum - ref) <= 1e-12 * scale).all()
    # Gram: z = x_p x_q scaled by 2^(40 - e_p - e_q); |z| < 2^(e_p + e_q) so |q| <= 2^40 again
    for p in range(P):
        for r in range(p, P):
            z = X[:, p] * X[:, r]
            qz = np.rint(z * np.ldexp(1.0, 40 - e[p] - e[r])).astype(np.int64)
            assert np.abs(qz).max() <= 1 << 40
            dz = digit_planes(qz)
            Sz = (counts[None, :] * dz).sum(axis=1)
            assert np.abs(Sz).max() < 2 ** 31
            g = sum(float(Sz[k]) * 128.0 ** k for k in range(DIGITS)) * np.ldexp(1.0, int(e[p] + e[r]) - 40)
            gref = float((counts * z).sum())
            assert abs(g - gref) <= 1e-12 * float((counts * np.abs(z)).sum())


def test_rounding_bound_of_one_element():
    # |x - q 2^(e-40)| <= 2^(e-41): half a unit of the least significant digit
    x = np.random.default_rng(2).normal(size=50000) * 13.7
    e = scale_exponent(np.array([np.abs(x).max()]))[0]
    q = np.rint(x * np.ldexp(1.0, 40 - e))
    assert np.abs(x - q * np.ldexp(1.0, e - 40)).max() <= np.ldexp(1.0, e - 41)
