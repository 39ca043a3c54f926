"""Executable description of the in-kernel bitonic network (csrc/nms.cuh: bitonic_warp_steps / block_bitonic_run): the same
index arithmetic -- E consecutive keys per thread, partner across lanes by `lane ^ (j / E)`, partner in the thread by `e | j`,
partner in another warp through the shared array, direction from bit k of the key's index -- replayed with numpy for every
size class the kernel uses.  It does not run the CUDA code (tests/test_gpu_parity.py::test_nms_segmented_sort_boundaries_and_score_ties
does); it pins the network itself, so an edit of the index rules can be checked on the CPU first."""
import numpy as np
import pytest

T = 512  # kNmsThreads


def warp_steps(v, E, k, jstart, active):
    tid = np.arange(T)
    lane, base = tid & 31, tid * E
    j = jstart
    while j >= E:                                   # partners in other lanes of the warp
        lm = j // E
        lower = (lane & lm) == 0
        for e in range(E):
            asc = ((base + e) & k) == 0
            o = v[tid ^ lm, e].copy()
            take_min = lower == asc
            new = np.where((o < v[:, e]) == take_min, o, v[:, e])
            v[:, e] = np.where(active, new, v[:, e])
        j >>= 1
    j = E // 2
    while j >= 1:                                   # partners in the same thread
        if j <= jstart:
            for e in range(E):
                if (e & j) == 0:
                    asc = ((base + e) & k) == 0
                    a, b = v[:, e].copy(), v[:, e | j].copy()
                    sw = (a > b) == asc
                    v[:, e], v[:, e | j] = np.where(sw, b, a), np.where(sw, a, b)
        j >>= 1


def block_bitonic_run(keys, npow, E):
    W = 32 * E
    base = np.arange(T) * E
    active = base < npow
    v = np.zeros((T, E), dtype=np.uint64)

    def load():
        for t in np.nonzero(active)[0]:
            v[t] = keys[base[t]:base[t] + E]

    def store():
        for t in np.nonzero(active)[0]:
            keys[base[t]:base[t] + E] = v[t]

    load()
    k = 2
    while k <= min(npow, W):
        warp_steps(v, E, k, k >> 1, active)
        k <<= 1
    k = 2 * W
    while k <= npow:
        store()
        j = k >> 1
        while j >= W:                               # partners in other warps: through shared memory
            for q in range(npow >> 1):
                i = ((q & ~(j - 1)) << 1) | (q & (j - 1))
                a, b = keys[i], keys[i | j]
                if (a > b) == ((i & k) == 0):
                    keys[i], keys[i | j] = b, a
            j >>= 1
        load()
        warp_steps(v, E, k, W >> 1, active)
        k <<= 1
    store()


@pytest.mark.parametrize("n", [1, 63, 64, 65, 129, 778, 1024, 1025, 2048])
def test_network_sorts_every_size_class(n):
    rng = np.random.default_rng(n)
    npow = 64
    while npow < n:
        npow <<= 1
    keys = np.zeros(2048, dtype=np.uint64)
    hi = rng.integers(0, 40, size=n).astype(np.uint64)                     # few distinct ranks: ties decided by the slot
    keys[:n] = (hi << np.uint64(32)) | np.arange(n, dtype=np.uint64)
    keys[n:npow] = np.uint64(0xFFFFFFFFFFFFFFFF)
    want = np.sort(keys[:n].copy())
    block_bitonic_run(keys, npow, 2 if npow <= 1024 else 4)
    assert np.array_equal(keys[:n], want)
    assert np.array_equal((keys[:n] >> np.uint64(32)), np.sort(hi))        # ascending rank, slots ascending inside a rank
