"""CPU restatements of the two pieces of knn_tc4_kernel (deep_gcns_torch_b200/csrc/knn_tc4.cuh) whose
correctness is an argument rather than a measurement: the sorting-network flush of the register list, and the
membership-by-interval-arithmetic path.  numpy only; no GPU, no product code path involved."""
import numpy as np

SENT = np.uint32(0xFFFFFFFF)


def _ce(v, i, j):
    lo, hi = np.minimum(v[i], v[j]), np.maximum(v[i], v[j])
    v[i], v[j] = lo, hi


def _t4_sort(v):
    """t4_sort<NN>: bitonic sorting network, ascending (the loops of the device code, verbatim)."""
    n = len(v)
    k = 2
    while k <= n:
        j = k >> 1
        while j > 0:
            for i in range(n):
                l = i ^ j
                if l > i:
                    if (i & k) == 0:
                        _ce(v, i, l)
                    else:
                        _ce(v, l, i)
            j >>= 1
        k <<= 1


def _t4_merge(v):
    """t4_merge<NN>: bitonic sequence -> ascending."""
    n = len(v)
    j = n >> 1
    while j > 0:
        for i in range(n):
            l = i ^ j
            if l > i:
                _ce(v, i, l)
        j >>= 1


def _flush(lk, batch):
    """One warp-wide flush: lk (32 ascending entries per lane), batch (<= 24 buffered entries per lane; slots
    16..23 only exist for some lanes).  All lanes at once: arrays are [entry][lane]."""
    lanes = lk.shape[1]
    for base, nb in ((0, 16), (16, 8)):
        bv = np.full((nb, lanes), SENT, dtype=np.uint32)
        for i in range(nb):
            have = batch[1] > base + i
            bv[i, have] = batch[0][base + i, have]
        if base == 16 and not np.any(batch[1] > 16):
            break
        rows = [bv[i].copy() for i in range(nb)]
        _t4_sort(rows)
        lst = [lk[i].copy() for i in range(32)]
        for i in range(nb):
            lst[32 - nb + i] = np.minimum(lst[32 - nb + i], rows[nb - 1 - i])
        _t4_merge(lst)
        lk[:] = np.stack(lst)
    return lk


def test_network_flush_keeps_the_32_smallest_sorted():
    rng = np.random.default_rng(0)
    lanes = 64
    lk = np.full((32, lanes), SENT, dtype=np.uint32)
    seen = [[] for _ in range(lanes)]
    for step in range(40):
        cnt = rng.integers(0, 25, size=lanes)
        if step % 7 == 0:
            cnt[:] = rng.integers(17, 25)                       # force the second (slots 16..23) pass
        vals = rng.integers(0, 2**31, size=(24, lanes), dtype=np.int64).astype(np.uint32)
        if step % 5 == 0:
            vals[:, ::3] = vals[0, ::3]                         # duplicates
        for ln in range(lanes):
            seen[ln].extend(int(x) for x in vals[:cnt[ln], ln])
        lk = _flush(lk, (vals, cnt))
        for ln in range(0, lanes, 7):
            want = sorted(seen[ln])[:32]
            got = [int(x) for x in lk[:, ln] if x != SENT]
            assert got == want[:len(got)] and len(got) == min(32, len(want))
        assert np.all(lk[1:] >= lk[:-1])


def _membership(a, eps, delta_of, K, cut, exact):
    """The set-only path for one query.  a: ascending approximate lower bounds of the listed candidates, exact: their
    exact keys (distance, index); returns the chosen set or None (query handed to the exact completion kernel)."""
    vK, vK1 = a[K - 1], a[K]
    hi = vK + delta_of(vK) + 2 * eps
    lo = vK1 - delta_of(vK1) - 2 * eps
    if not hi < cut:
        return None
    chosen = [u for u in range(len(a)) if a[u] < lo and u < K]
    band = [u for u in range(len(a)) if not a[u] < lo and a[u] <= hi]
    if len(band) > 12:
        return None
    need = K - len(chosen)
    band.sort(key=lambda u: exact[u])
    return set(chosen) | set(band[:need])


def test_interval_membership_equals_exact_top_k():
    """Random lists with gaps both far above and far below the error band, duplicates straddling rank K, and
    errors drawn up to the bound: whenever the path answers, its set is the exact top-K set; and a candidate outside
    the list (exact >= cut - eps) can never belong to that set."""
    rng = np.random.default_rng(1)
    K, KP = 20, 28
    answered = 0
    for trial in range(4000):
        scale = 10.0 ** rng.uniform(-3, 1)                      # candidate spacing from far below to far above the band
        exact_d = np.sort(rng.uniform(60, 60 + scale * KP, size=KP + 40))
        if trial % 4 == 0:
            exact_d[K - 2:K + 2] = exact_d[K - 1]               # exact ties across the boundary: the index decides
        eps = 4e-3
        delta_of = lambda v: 2.0 ** -10 * (v + 64.0)            # noqa: E731
        err = rng.uniform(-eps, eps, size=exact_d.size)
        trunc = rng.uniform(0, 1, size=exact_d.size) * delta_of(exact_d) * 0.5
        approx_lb = exact_d + err - trunc                       # a <= approx <= a + delta, |approx - exact| <= eps
        order = np.argsort(approx_lb, kind="stable")
        listed, unlisted = order[:KP], order[KP:]
        a = approx_lb[listed]
        cut = a[KP - 1]                                         # every unlisted candidate has approx >= cut
        assert np.all(approx_lb[unlisted] >= cut)
        idx = rng.permutation(exact_d.size)
        exact_key = [(exact_d[c], idx[c]) for c in listed]
        got = _membership(a, eps, delta_of, K, cut, exact_key)
        if got is None:
            continue
        answered += 1
        all_keys = sorted((exact_d[c], idx[c]) for c in range(exact_d.size))[:K]
        want = {u for u in range(KP) if exact_key[u] in set(all_keys)}
        assert got == want and len(got) == K, trial
    assert answered > 2000


def _f2u(x):
    return np.asarray(x, dtype=np.float32).view(np.uint32)


def _u2f(u):
    return np.asarray(u, dtype=np.uint32).view(np.float32)


def test_list_value_brackets_the_approximate_distance_within_delta():
    """The interval test assumes a_c <= approx_c <= a_c + delta(a_c) with delta(v) = 2^-10 (v + |x_i|^2) for the value
    a_c a list entry carries.  Restate the device path of a candidate - accumulator -> 20-bit packed buffer entry ->
    upper bound of the accumulator -> d2 = max(fma(-2, ab, |x_i|^2), 0) -> 20-bit list value - in float32 and check the
    bracket against the accumulator's own distance, over magnitudes from 1e-3 to 1e6 and both accumulator signs."""
    rng = np.random.default_rng(2)
    for scale in (1e-3, 1.0, 64.0, 1e3, 1e6):
        sqq = np.float32(scale * rng.uniform(0.2, 2.0))
        # approximate squared distances around and far above |x_i|^2 (accumulator negative and positive)
        d2 = np.concatenate([rng.uniform(0, 4 * scale, 200000), rng.uniform(0, 1e-3 * scale, 2000),
                             np.geomspace(1e-6 * scale, 64 * scale, 2000)])
        acc = ((np.float64(sqq) - d2) * 0.5).astype(np.float32)         # the tensor core's accumulator = -key/2
        approx = np.float64(sqq) - 2.0 * acc.astype(np.float64)          # the distance that accumulator stands for
        idx = rng.integers(0, 4096, size=acc.size).astype(np.uint32)
        en = (_f2u(acc) & np.uint32(0xFFFFF000)) | idx                    # filter: one LOP3
        neg = (en & np.uint32(0x80000000)) != 0
        ab = _u2f(np.where(neg, en & np.uint32(0xFFFFF000), en | np.uint32(0xFFF)))
        d2p = np.maximum((np.float64(-2.0) * ab.astype(np.float64) + np.float64(sqq)).astype(np.float32), np.float32(0))
        a = _u2f(_f2u(d2p) & np.uint32(0xFFFFF000)).astype(np.float64)   # the list value
        slack = 2.0 ** -22 * (np.abs(approx) + float(sqq))               # one fp32 rounding of the fma, inside eps
        assert np.all(a <= np.maximum(approx, 0) + slack)
        delta = 2.0 ** -10 * (a + float(sqq))
        assert np.all(np.maximum(approx, 0) <= a + delta + slack)


def test_shuffle_bitonic_sort_of_the_128_key_sample():
    """select_rows_fast_kernel sorts its 128-key sample in registers: element e = 32 u + lane lives in smp[u] of
    lane `lane`; exchange distances below 32 are __shfl_xor, 32 and 64 are register pairs.  A wrong network would
    not break results (the sample only bounds the K-th distance) - it would silently cost retries - so the
    direction logic is restated here and checked against a plain sort."""
    rng = np.random.default_rng(3)
    for trial in range(50):
        keys = rng.integers(0, 2**32, size=128, dtype=np.uint64).astype(np.uint32)
        if trial % 5 == 0:
            keys[rng.integers(0, 128, size=40)] = keys[0]                 # duplicates
        smp = keys.reshape(4, 32).copy()                                  # smp[u][lane]
        lane = np.arange(32)
        kk = 2
        while kk <= 128:
            j = kk >> 1
            while j > 0:
                if j < 32:
                    for u in range(4):
                        other = smp[u][lane ^ j]                          # __shfl_xor_sync(..., j)
                        up = ((32 * u + lane) & kk) == 0
                        lower = (lane & j) == 0
                        smp[u] = np.where(lower == up, np.minimum(smp[u], other), np.maximum(smp[u], other))
                else:
                    jr = j >> 5
                    for u in range(4):
                        if (u & jr) == 0:
                            up = ((32 * u) & kk) == 0
                            lo, hi = np.minimum(smp[u], smp[u ^ jr]), np.maximum(smp[u], smp[u ^ jr])
                            smp[u], smp[u ^ jr] = (lo, hi) if up else (hi, lo)
                j >>= 1
            kk <<= 1
        assert np.array_equal(smp.reshape(-1), np.sort(keys))             # lane l holds samples l, l+32, l+64, l+96
