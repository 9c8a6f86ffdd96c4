"""CPU checks of the bit-level and error-bound arguments the tensor-core kNN path relies on
(deep_gcns_torch_b200/csrc/knn_tc.cuh).  They restate the device expressions with numpy /
torch on the host; no GPU, no product code path involved."""
import numpy as np
import torch


def _f2u(x):
    return np.asarray(x, dtype=np.float32).view(np.uint32)


def _u2f(u):
    return np.asarray(u, dtype=np.uint32).view(np.float32)


def test_lop3_entry_packing_identity():
    """The filter builds a list entry with ONE lop3 (LUT 0xE6 = (a & b & c) | (b ^ c)): a = accumulator
    bits, b = 0xFFFFF000 | (first column of the 32-column chunk), c = 0xFFFFF000 | (column within the chunk).
    It must equal (a & 0xFFFFF000) | column for every column below 4096."""
    rng = np.random.default_rng(0)
    a = rng.integers(0, 2**32, size=4096, dtype=np.uint64).astype(np.uint32)
    for base in range(0, 4096, 32):
        b = np.uint32(0xFFFFF000 | base)
        for i in range(32):
            c = np.uint32(0xFFFFF000 | i)
            got = (a & b & c) | (b ^ c)
            want = (a & np.uint32(0xFFFFF000)) | np.uint32(base + i)
            assert np.array_equal(got, want), (base, i)
    # the LUT constant itself: bit (a<<2 | b<<1 | c) of 0xE6 is f(a, b, c)
    for av in (0, 1):
        for bv in (0, 1):
            for cv in (0, 1):
                assert ((0xE6 >> (av << 2 | bv << 1 | cv)) & 1) == ((av & bv & cv) | (bv ^ cv))


def test_packed_entries_stay_lower_bounds_of_the_key():
    """Flush restores an UPPER bound of the accumulator from an entry whose low 12 mantissa bits were
    replaced by the index (positive: set them, negative: clear them), so key = -2 acc is bounded from
    BELOW; the list then stores the distance with its low 12 bits cleared - again a lower bound - and the
    admission threshold is one truncation step above the worst entry."""
    rng = np.random.default_rng(1)
    acc = np.concatenate([rng.normal(0, 40, 200000), rng.normal(0, 1e-3, 1000), [0.0, -0.0, 1e-30, -1e30]]).astype(np.float32)
    idx = rng.integers(0, 4096, size=acc.size).astype(np.uint32)
    en = (_f2u(acc) & np.uint32(0xFFFFF000)) | idx
    neg = (en & np.uint32(0x80000000)) != 0
    ub = _u2f(np.where(neg, en & np.uint32(0xFFFFF000), en | np.uint32(0xFFF)))
    assert np.all(ub >= acc)
    assert np.all(np.abs(ub - acc) <= np.abs(acc) * 2.0**-11 + 1e-37)
    assert np.array_equal(en & np.uint32(0xFFF), idx)
    sqq = np.float32(63.7)
    key_lb = np.float32(-2.0) * ub
    assert np.all(key_lb <= np.float32(-2.0) * acc)
    d2 = np.maximum(key_lb + sqq, np.float32(0))               # fmaf(-2, ub, sqq) rounds once; monotone either way
    stored = _u2f(_f2u(d2) & np.uint32(0xFFFFF000))
    assert np.all(stored <= d2) and np.all(stored >= 0)
    tau = _u2f((_f2u(stored) & np.uint32(0xFFFFF000)) + np.uint32(0x1000)) - sqq     # admission threshold (key units)
    # A rejected candidate has key > tau, hence distance > stored - (rounding of the fp32 subtraction), which
    # the 2^-20 (|x_i|^2 + max|x_j|^2) term of eps absorbs with room to spare.
    finite = np.isfinite(tau) & (stored < 3e38)
    t64, s64 = tau.astype(np.float64)[finite], stored.astype(np.float64)[finite]
    assert np.all(t64 + float(sqq) >= s64 - 2.0**-23 * (float(sqq) + s64))
    assert np.all(2.0**-23 * (float(sqq) + s64) <= 9.537e-7 * (float(sqq) + s64) / 8)


def test_three_product_bf16_split_error_is_inside_eps():
    """|approx - exact| of the pre-filter key against the certificate's eps: x = hi + mid in bf16, products
    hi*hi + hi*mid + mid*hi (mid*mid dropped), -|x_j|^2/2 as three bf16 terms, fp32 accumulation."""
    g = torch.Generator().manual_seed(0)
    for C, scale in ((64, 1.0), (64, 30.0), (16, 0.01), (3, 5.0), (48, 1e3), (16, -1.0), (64, -1.0)):
        cpad = (C + 15) // 16 * 16
        if scale < 0:
            # adversarial for the split: every component sits at the worst case of both roundings
            # (1 + 2^-8 + 2^-17: hi drops 2^-8, mid drops 2^-17), mixed with bf16-exact values and signs
            worst = 1.0 + 2.0**-8 + 2.0**-17
            pick = torch.randint(0, 3, (512, C), generator=g)
            sign = torch.randint(0, 2, (512, C), generator=g) * 2.0 - 1.0
            x = torch.where(pick == 0, torch.tensor(worst), torch.where(pick == 1, torch.tensor(1.5), torch.tensor(worst * 2)))
            x = (x * sign).float()
        else:
            x = torch.randn(512, C, generator=g) * scale
        x[:8] = x[8:16] * (1 + 1e-4)                               # near-duplicates
        hi = x.to(torch.bfloat16).float()
        mid = (x - hi).to(torch.bfloat16).float()
        sq = (x.double() ** 2).sum(1)
        s = (-0.5 * sq).float()
        parts, rem = [], s.clone()
        for _ in range(3):
            h = rem.to(torch.bfloat16).float()
            parts.append(h)
            rem = rem - h
        # fp32 accumulation in an arbitrary (here: product-major) order, like the tensor core's
        acc = torch.zeros(512, 512)
        for a, b in ((hi, hi), (hi, mid), (mid, hi)):
            acc = acc + (a @ b.t())
        acc = acc + sum(parts)[None, :]
        key = -2.0 * acc                                           # approximate |x_j|^2 - 2 x_i.x_j
        exact = (sq[None, :] - 2.0 * (x.double() @ x.double().t()))
        smax = sq.max()
        eps = (2.0 * (3.0518e-5 + (5.0 * cpad + 8.0) * 1.1921e-7)) * torch.sqrt(sq[:, None] * smax) \
            + 9.537e-7 * (sq[:, None] + smax)
        err = (key.double() - exact).abs()
        # the kernel's exact side is an fp32 FMA chain: allow its own rounding (C * 2^-24 relative) on top
        fp32_chain = (C + 2) * 2.0**-24 * (2 * torch.sqrt(sq[:, None] * sq[None, :]) + sq[None, :] + sq[:, None])
        assert torch.all(err <= eps.double() - fp32_chain), (C, scale, float((err / eps).max()))
