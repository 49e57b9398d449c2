"""CPU: an executable specification of the register bitonic networks and the merge-tree bootstrap of
torchpq_b200/csrc/common.cuh (reg_sort / reg_merge / warp_bitonic_merge_desc / warp_merge_run / warp_merge_runs /
CtaTopK::cta_bootstrap).  The functions below are line-by-line transliterations of the index arithmetic of those
routines (element r*32 + lane of a warp's run, shuffle distance j, direction bit k, run regions 64 keys apart, the
gap-closing copy for 32-key runs) with "all lanes at once" semantics for a shuffle stage; they are checked against
`sorted`.  The GPU tests check the kernels' results; this file pins the networks' index math where no GPU is needed."""
import random

import pytest


def _cswap(x, a, b, lane, desc):
    hi, lo = max(x[a][lane], x[b][lane]), min(x[a][lane], x[b][lane])
    x[a][lane], x[b][lane] = (hi, lo) if desc else (lo, hi)


def _shuffle_stage(x, R, j, k, asc, merge):
    new = [[0] * 32 for _ in range(R)]
    for lane in range(32):
        up = (lane & j) != 0
        for r in range(R):
            y = x[r][lane ^ j]                                   # __shfl_xor_sync(x[r], j)
            if merge:
                new[r][lane] = min(x[r][lane], y) if (up != asc) else max(x[r][lane], y)
            else:
                desc = ((((r * 32) | lane) & k) == 0) != asc
                new[r][lane] = max(x[r][lane], y) if (desc != up) else min(x[r][lane], y)
    for r in range(R):
        x[r] = new[r]


def reg_merge(x, R, asc=False):
    for lane in range(32):
        if R == 4:
            _cswap(x, 0, 2, lane, not asc); _cswap(x, 1, 3, lane, not asc)
        if R >= 2:
            for r in range(0, R, 2):
                _cswap(x, r, r + 1, lane, not asc)
    j = 16
    while j > 0:
        _shuffle_stage(x, R, j, 0, asc, True)
        j >>= 1


def reg_sort(x, R, asc=False):
    k = 2
    while k <= 32 * R:
        j = k >> 1
        if R == 4 and k == 128:
            for lane in range(32):
                _cswap(x, 0, 2, lane, not asc); _cswap(x, 1, 3, lane, not asc)
            j = 32
        if R >= 2 and j == 32:
            for lane in range(32):
                for r in range(0, R, 2):
                    _cswap(x, r, r + 1, lane, (((r * 32) & k) == 0) != asc)
            j = 16
        while j > 0:
            _shuffle_stage(x, R, j, k, asc, False)
            j >>= 1
        k <<= 1


def _flat(x, R):
    return [x[r][l] for r in range(R) for l in range(32)]


def _regs(v, R):
    return [[v[r * 32 + l] for l in range(32)] for r in range(R)]


def bitonic_merge_desc(a, off, n):
    j = n >> 1
    while j >= 64:
        for t in range(n >> 1):
            lo = ((t & ~(j - 1)) << 1) | (t & (j - 1)); hi = lo + j
            if a[off + lo] < a[off + hi]:
                a[off + lo], a[off + hi] = a[off + hi], a[off + lo]
        j >>= 1
    if n == 32:
        x = _regs(a[off:off + 32], 1); reg_merge(x, 1); a[off:off + 32] = _flat(x, 1)
    else:
        for c in range(0, n, 64):
            x = _regs(a[off + c:off + c + 64], 2); reg_merge(x, 2); a[off + c:off + c + 64] = _flat(x, 2)


def merge_run(A, n, x, R):
    tail = n - 32 * R
    if n == 32 * R:
        y = [[max(A[tail + r * 32 + l], x[r][l]) for l in range(32)] for r in range(R)]
        reg_merge(y, R)
        A[tail:] = _flat(y, R)
        return
    for r in range(R):
        for l in range(32):
            A[tail + r * 32 + l] = max(A[tail + r * 32 + l], x[r][l])
    bitonic_merge_desc(A, 0, n)


def merge_runs(buf, a, b, ln, top_only):
    for t in range(ln):
        x, y = buf[a + t], buf[b + ln - 1 - t]
        buf[a + t] = max(x, y)
        if not top_only:
            buf[b + ln - 1 - t] = min(x, y)
    bitonic_merge_desc(buf, a, ln)
    if not top_only:
        bitonic_merge_desc(buf, b, ln)


def cta_bootstrap(keys_per_warp, nw, R, kp):
    scratch = [0] * (nw * 64)
    ln = 32 if (R == 1 or kp < 64) else 64
    for w in range(nw):
        x = _regs(keys_per_warp[w] + [0] * (64 - 32 * R), 2) if R == 2 else _regs(keys_per_warp[w], 1)
        reg_sort(x, 2 if R == 2 else 1)
        scratch[w * 64:w * 64 + ln] = _flat(x, 2 if R == 2 else 1)[:ln]
    s = 1
    while s < nw:
        top_only = 2 * ln > kp
        for w in range(nw):
            if (w & (2 * s - 1)) == 0 and w + s < nw:
                a, b = w * 64, (w + s) * 64
                merge_runs(scratch, a, b, ln, top_only)
                if not top_only and b != a + ln:
                    scratch[a + ln:a + 2 * ln] = scratch[b:b + ln]
        if not top_only:
            ln <<= 1
        s <<= 1
    lst = [0] * kp
    for t in range(ln):                                        # warp_merge_desc(list, kp, scratch, len)
        i = kp - ln + t
        lst[i] = max(lst[i], scratch[ln - 1 - t])
    bitonic_merge_desc(lst, 0, kp)
    return lst


@pytest.mark.parametrize("R", [1, 2, 4])
@pytest.mark.parametrize("asc", [False, True])
def test_register_sort_and_merge(R, asc):
    rng = random.Random(R * 2 + asc)
    for _ in range(25):
        v = [rng.randrange(0, 60) for _ in range(32 * R)]         # many ties
        x = _regs(v, R)
        reg_sort(x, R, asc)
        assert _flat(x, R) == sorted(v, reverse=not asc)
        h = 16 * R
        x = _regs(sorted(v[:h], reverse=True) + sorted(v[h:]), R)  # a bitonic sequence
        reg_merge(x, R, asc)
        assert _flat(x, R) == sorted(v, reverse=not asc)


@pytest.mark.parametrize("kp", [32, 64, 128, 256, 1024])
@pytest.mark.parametrize("R", [1, 2, 4])
def test_merge_run_into_list(kp, R):
    if 32 * R > kp:
        pytest.skip("the run is capped to the list length before the merge (merge_run_capped)")
    rng = random.Random(kp + R)
    for _ in range(8):
        lst = sorted([rng.randrange(0, 1000) for _ in range(kp)], reverse=True)
        run = sorted(rng.randrange(0, 1000) for _ in range(32 * R))
        A = lst[:]
        merge_run(A, kp, _regs(run, R), R)
        assert A == sorted(lst + run, reverse=True)[:kp]


@pytest.mark.parametrize("nw", [4, 8, 16])
@pytest.mark.parametrize("R", [1, 2])
@pytest.mark.parametrize("kp", [32, 64, 128, 256, 1024])
def test_merge_tree_bootstrap(nw, R, kp):
    rng = random.Random(nw * 100 + R * 10 + kp)
    for _ in range(3):
        keys = [[rng.randrange(1, 10 ** 6) for _ in range(32 * R)] for _ in range(nw)]
        got = cta_bootstrap(keys, nw, R, kp)
        expect = sorted((k for w in keys for k in w), reverse=True)[:kp]
        assert got == expect + [0] * (kp - len(expect))
