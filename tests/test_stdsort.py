"""CPU test: the step-for-step replay of libstdc++'s std::sort (csrc/dq_stdsort.cuh, the code the frame pipeline runs on
the device to order the palette by r+g+b) against std::sort itself on the reference's element type and comparator
(DivQuantMapColors.cpp:227-238, 314-323).  The permutation of EQUAL keys is what matters: it decides which of two
equidistant palette colours a pixel is mapped to."""
import ctypes as C

import numpy as np

_u32p = C.POINTER(C.c_uint32)


def _perm(lib, keys, replay):
    keys = np.ascontiguousarray(keys, np.uint32)
    out = np.zeros(keys.size, np.uint32)
    lib.dq_host_sort_permutation(keys.ctypes.data_as(_u32p), keys.size, replay, out.ctypes.data_as(_u32p))
    return out


def _killer(n):
    """Musser's median-of-three killer: drives introsort to its depth limit, i.e. into the heap sort."""
    k = n // 2
    a = np.zeros(n, np.uint32)
    for i in range(1, k + 1):
        if i % 2 == 1:
            a[i - 1] = i
            a[i] = k + i
        a[k + i - 1] = 2 * i
    return a


def test_replay_equals_std_sort(pkg):
    lib = pkg.load_library()
    rng = np.random.default_rng(7)
    cases = []
    for n in list(range(1, 40)) + [64, 125, 255, 256, 257, 300, 511, 512, 1000, 3000]:
        cases.append(rng.integers(0, 766, n))                       # sums of a random palette
        cases.append(rng.integers(0, max(2, n // 8), n))            # many equal keys
        cases.append(np.zeros(n, np.uint32))                        # all equal
        cases.append(np.arange(n) % 766)                            # sorted runs
        cases.append((np.arange(n)[::-1]) % 766)                    # reversed
        cases.append(np.minimum(np.arange(n), np.arange(n)[::-1]))  # organ pipe
        cases.append(_killer(n) % 65536)
    vals = [0, 63, 127, 191, 255]
    cases.append(np.array([r + g + b for r in vals for g in vals for b in vals]))  # the live call's 125-colour grid
    for _ in range(300):
        n = int(rng.integers(1, 600))
        cases.append(rng.integers(0, int(rng.integers(1, 766)), n))
    for keys in cases:
        a, b = _perm(lib, keys, 0), _perm(lib, keys, 1)
        assert np.array_equal(a, b), (len(keys), keys[:20])
        assert np.array_equal(np.sort(a), np.arange(len(keys)))     # a permutation
        assert np.all(np.diff(np.asarray(keys)[a].astype(np.int64)) >= 0)  # sorted by key
