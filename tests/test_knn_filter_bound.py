"""Host-side check of the inequality csrc/knn.cu rests on: the float32 value the filter computes,

    lb = fl32( n~_j - 2 * dot32(a, b) ),   n~ = ||p||^2 * (1 - (DP + 8) * 2^-24)  (rounded down to float32),

compared with  ru32( tau - n~_i ),  never rejects a pair whose true squared distance is <= tau.  The kernel's
arithmetic is replayed in numpy: float32 operands, a DP-term dot product accumulated by fused multiply-adds in
the kernel's order (emulated in float64: a 24 x 24-bit product is exact there), the final fma, the directed
roundings of the two bounds.  Inputs are chosen to stress it: large common offsets that centring cannot remove,
tight clusters, wildly mixed scales, ties (tau exactly the pair's own distance)."""

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

EPS = 2.0 ** -24


def _f32(x):
    return np.asarray(x, dtype=np.float64).astype(np.float32).astype(np.float64)


def _round_down32(x):
    y = np.asarray(x, dtype=np.float64).astype(np.float32)
    y = np.where(y.astype(np.float64) > x, np.nextafter(y, np.float32(-np.inf)), y)
    return y.astype(np.float64)


def _round_up32(x):
    y = np.asarray(x, dtype=np.float64).astype(np.float32)
    y = np.where(y.astype(np.float64) < x, np.nextafter(y, np.float32(np.inf)), y)
    return y.astype(np.float64)


def _filter_passes(A, B, tau, DP):
    """A: (r, d) queries, B: (r, d) points (already centred, float64), tau: (r,) true squared-distance bounds.
    Returns the kernel's verdict per pair."""
    d = A.shape[1]
    shrink = 1.0 - (DP + 8) * EPS
    na, nb = np.sum(A * A, axis=1), np.sum(B * B, axis=1)
    a32, b32 = _f32(A), _f32(B)
    acc = np.zeros(A.shape[0])
    for k in range(d):                                  # acc = fmaf(a[k], b[k], acc): one rounding per step
        acc = _f32(a32[:, k] * b32[:, k] + acc)
    nb_lo = _round_down32(nb * shrink)                  # __double2float_rd(nrm64 * shrink)
    lb = _f32(nb_lo - 2.0 * acc)                        # fmaf(-2, acc, nb_lo)
    tq = _round_up32(tau - na * shrink)                 # __double2float_ru(tau - nlo64)
    return lb <= tq


def _true_d2(A, B):
    return np.sum((A - B) ** 2, axis=1)


@settings(max_examples=60, deadline=None)
@given(seed=st.integers(0, 2**31 - 1), d=st.sampled_from([1, 3, 8, 15, 30, 32, 47, 64]),
       offset=st.sampled_from([0.0, 1.0, 1e3, 1e6]), spread=st.sampled_from([1e-6, 1e-3, 1.0, 1e3]))
def test_no_true_neighbour_is_filtered_out(seed, d, offset, spread):
    DP = 8 if d <= 8 else 16 if d <= 16 else 32 if d <= 32 else 64
    rng = np.random.default_rng(seed)
    r = 4000
    A = offset + rng.normal(size=(r, d)) * spread * np.logspace(-2, 2, d)
    B = A + rng.normal(size=(r, d)) * spread * rng.choice([1e-9, 1e-6, 1e-3, 1.0], size=(r, 1))
    B[: r // 8] = A[: r // 8]                            # coincident pairs
    d2 = _true_d2(A, B)
    # a pair must pass whenever tau >= its true distance: try tau exactly equal (tie) and slightly above
    for tau in (d2, d2 * (1 + 1e-12), d2 + 1e-300):
        assert np.all(_filter_passes(A, B, tau, DP)), "the filter rejected a pair that is within the bound"


def test_the_filter_does_filter():
    """...and it is not vacuous: far pairs are rejected once tau is the scale of near pairs."""
    rng = np.random.default_rng(3)
    A = rng.normal(size=(5000, 30)) * 0.01
    B = rng.normal(size=(5000, 30)) * 0.01
    d2 = _true_d2(A, B)
    tau = np.full(5000, np.quantile(d2, 0.01))
    passed = _filter_passes(A, B, tau, 32)
    assert passed[d2 <= tau].all()
    assert passed.mean() < 0.02


# ---- the tensor-core filter: tf32 hi/lo split, three passes, fp32 accumulation of unknown rounding ----------

def _tf32(x):
    """cvt.rna.tf32.f32: round to nearest (ties away) on a 10-bit mantissa, as float64 values."""
    x = np.asarray(x, dtype=np.float64).astype(np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    u = ((u + 0x1000) & ~np.uint64(0x1FFF)).astype(np.uint32)
    return u.view(np.float32).astype(np.float64)


def _trunc32(x):
    """float64 -> float32 toward zero: the most pessimistic accumulator rounding a tensor core could use."""
    y = np.asarray(x, dtype=np.float64).astype(np.float32)
    over = np.abs(y.astype(np.float64)) > np.abs(x)
    y = np.where(over, np.nextafter(y, np.float32(0)), y)
    return y.astype(np.float64)


def _mma_filter_passes(A, B, tau, DP, accumulate):
    d = A.shape[1]
    shrink = 1.0 - 2.0 ** -14
    na, nb = np.sum(A * A, axis=1), np.sum(B * B, axis=1)
    ah = _tf32(_f32(A)); al = _tf32(_f32(A - ah))
    bh = _tf32(_f32(B)); bl = _tf32(_f32(B - bh))
    acc = np.zeros(A.shape[0])
    for k0 in range(0, d, 8):                           # one mma.m16n8k8 per pass and k-step
        for x, y in ((al, bh), (ah, bl), (ah, bh)):
            for k in range(k0, min(k0 + 8, d)):
                acc = accumulate(acc + x[:, k] * y[:, k])
    nb_lo = _round_down32(nb * shrink)
    lb = _f32(nb_lo - 2.0 * acc)
    tq = _round_up32(tau - na * shrink)
    return lb <= tq


@settings(max_examples=40, deadline=None)
@given(seed=st.integers(0, 2**31 - 1), d=st.sampled_from([1, 3, 8, 15, 30, 32]),
       offset=st.sampled_from([0.0, 1.0, 1e3, 1e6]), spread=st.sampled_from([1e-6, 1e-3, 1.0, 1e3]))
def test_tensor_core_filter_keeps_every_true_neighbour(seed, d, offset, spread):
    """csrc/knn.cu, knn_scan_mma_kernel: with the norms shrunk by 2^-14 the bound holds whether the tensor core
    rounds its fp32 accumulator to nearest or chops it after every single product."""
    DP = 8 if d <= 8 else 16 if d <= 16 else 32
    rng = np.random.default_rng(seed)
    r = 3000
    A = offset + rng.normal(size=(r, d)) * spread * np.logspace(-2, 2, d)
    B = A + rng.normal(size=(r, d)) * spread * rng.choice([1e-9, 1e-6, 1e-3, 1.0], size=(r, 1))
    B[: r // 8] = A[: r // 8]
    d2 = _true_d2(A, B)
    for accumulate in (_f32, _trunc32):
        for tau in (d2, d2 * (1 + 1e-12)):
            assert np.all(_mma_filter_passes(A, B, tau, DP, accumulate))


def test_tensor_core_filter_does_filter():
    rng = np.random.default_rng(4)
    A = rng.normal(size=(5000, 30)) * 0.01
    B = rng.normal(size=(5000, 30)) * 0.01
    d2 = _true_d2(A, B)
    tau = np.full(5000, np.quantile(d2, 0.01))
    passed = _mma_filter_passes(A, B, tau, 32, _trunc32)
    assert passed[d2 <= tau].all()
    assert passed.mean() < 0.03
