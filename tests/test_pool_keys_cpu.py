"""Bit tricks of the fused train-mode max-pool (csrc/gemm.cuh, EPI_STATS_POOL; reference: `torch.max(x, 2)` of pcs.py:114,
whose arg-max decides where the gradient goes) restated in numpy and checked against plain float arithmetic.

Inside a 128-row tile a candidate is ONE signed 32-bit key: the order-preserving image of +-x (x is a bf16, so the low 16 bits
of its fp32 pattern are free) OR-ed with 0xFFFF - row; an integer max then keeps the FIRST row among equal values.  The tile
winner is converted to the packed 64-bit form (orderable(value) << 32 | ~row_in_cloud) that `atomicMax` merges across tiles
and that `k_maxpool_finish` decodes (pointwise.cuh: float_orderable / float_from_orderable)."""
import numpy as np


def _bf16_bits(rng, n):
    """random bf16 values as fp32 bit patterns (low 16 bits zero): normals of both signs, zeros, repeated values"""
    v = rng.standard_normal(n).astype(np.float32) * rng.choice([1e-3, 1.0, 50.0], n).astype(np.float32)
    b = v.view(np.uint32) & np.uint32(0xFFFF0000)
    b[rng.random(n) < 0.05] = 0                               # +0
    dup = rng.random(n) < 0.3                                 # ties: copy another element's value
    b[dup] = b[rng.integers(0, n, dup.sum())]
    return b


def _key32(bits, flip, row):
    f = bits ^ flip
    m = (f.view(np.int32) >> 31).view(np.uint32) & np.uint32(0x7FFF0000)
    return ((f ^ m) | (np.uint32(0xFFFF) - row.astype(np.uint32))).view(np.int32)


def _orderable(b):          # pointwise.cuh: float_orderable
    return np.where(b & np.uint32(0x80000000), ~b, b | np.uint32(0x80000000)).astype(np.uint32)


def _from_orderable(k):     # pointwise.cuh: float_from_orderable
    return np.where(k & np.uint32(0x80000000), k & np.uint32(0x7FFFFFFF), ~k).astype(np.uint32)


def test_key32_orders_like_floats_and_prefers_the_first_row():
    rng = np.random.default_rng(0)
    for flip in (np.uint32(0), np.uint32(0x80000000)):        # gamma >= 0: max, gamma < 0: min (= max of -x)
        for _ in range(200):
            bits = _bf16_bits(rng, 128)
            rows = np.arange(128)
            keys = _key32(bits, flip, rows)
            x = (bits ^ flip).view(np.float32)                # the value the max runs over
            win = int(np.argmax(keys))                        # (argmax of distinct keys: rows differ, so no key ties)
            assert len(np.unique(keys)) == 128
            # first index of the float maximum; -0.0 and +0.0 compare equal as floats but -0 < +0 as keys, which can only
            # matter when the maximum itself is a zero: skip those draws (exact zeros of both signs in one column of a
            # GEMM accumulator do not occur: a sum that contains a +0 term is +0)
            if x.max() == 0.0 and np.signbit(x[x == 0.0]).any() and not np.signbit(x[x == 0.0]).all():
                continue
            assert win == int(np.argmax(x)), (win, int(np.argmax(x)))


def test_key32_to_packed_key64_roundtrip():
    rng = np.random.default_rng(1)
    bits = _bf16_bits(rng, 4096)
    rows = rng.integers(0, 128, 4096)
    row0 = rng.integers(0, 1 << 20, 4096) * 128               # first row of the tile inside its cloud
    for flip in (np.uint32(0), np.uint32(0x80000000)):
        k = _key32(bits, flip, rows).view(np.uint32)
        sord = k & np.uint32(0xFFFF0000)
        rwin = np.uint32(0xFFFF) - (k & np.uint32(0xFFFF))
        ordv = (sord ^ np.uint32(0x80000000)) | np.where(sord & np.uint32(0x80000000), np.uint32(0xFFFF), np.uint32(0))
        key64 = (ordv.astype(np.uint64) << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - row0.astype(np.uint64) - rwin.astype(np.uint64))
        # the packed form is what the straddling-tile path (and round 1) builds directly from the float
        f = bits ^ flip
        want = (_orderable(f).astype(np.uint64) << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - (row0 + rows).astype(np.uint64))
        assert (key64 == want).all()
        # and the decoder gets value and row back
        assert (_from_orderable((key64 >> np.uint64(32)).astype(np.uint32)) == f).all()
        assert ((np.uint64(0xFFFFFFFF) - (key64 & np.uint64(0xFFFFFFFF))) == (row0 + rows).astype(np.uint64)).all()


def test_packed_keys_merge_across_tiles_like_a_global_argmax():
    rng = np.random.default_rng(2)
    for _ in range(50):
        n_tiles = 6
        bits = _bf16_bits(rng, 128 * n_tiles)
        flip = np.uint32(0x80000000) if rng.random() < 0.5 else np.uint32(0)
        best = np.uint64(0)
        for t in range(n_tiles):
            b = bits[128 * t:128 * (t + 1)]
            k = int(_key32(b, flip, np.arange(128)).max())
            ku = np.uint32(k & 0xFFFFFFFF)
            sord = ku & np.uint32(0xFFFF0000)
            rwin = np.uint32(0xFFFF) - (ku & np.uint32(0xFFFF))
            ordv = (sord ^ np.uint32(0x80000000)) | (np.uint32(0xFFFF) if (sord & np.uint32(0x80000000)) else np.uint32(0))
            key64 = (np.uint64(ordv) << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - np.uint64(128 * t) - np.uint64(rwin))
            best = max(best, key64)
        x = (bits ^ flip).view(np.float32)
        if x.max() == 0.0:
            continue
        assert int(np.uint64(0xFFFFFFFF) - (best & np.uint64(0xFFFFFFFF))) == int(np.argmax(x))
