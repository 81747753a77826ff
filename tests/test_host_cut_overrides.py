"""CPU test of the host logic behind dq_call_stats::cut_overrides (csrc/dq_context.cu, select_cut_overrides): after the
tie resolver has looked at the flagged cuts of a frame, the ones the reference makes elsewhere (status 3) are forced in
the next run of the split -- outermost first: a flagged cut inside the subtree of another one that changes is decided by
the run after, because its node may not exist any more.  Ranges are [begin, begin + size) of point positions; a node's
range contains the ranges of all nodes below it."""
import ctypes as C

import numpy as np

_u32p = C.POINTER(C.c_uint32)


def _select(lib, ranges, status, already=0):
    b = np.array([r[0] for r in ranges], np.uint32)
    s = np.array([r[1] for r in ranges], np.uint32)
    st = np.array(status, np.uint32)
    out = np.zeros(32, np.uint32)
    n = lib.dq_host_select_cut_overrides(b.ctypes.data_as(_u32p), s.ctypes.data_as(_u32p), st.ctypes.data_as(_u32p), len(ranges),
                                         already, out.ctypes.data_as(_u32p))
    return [int(x) for x in out[:n]]


def test_only_differing_cuts_are_forced(pkg):
    lib = pkg.load_library()
    ranges = [(0, 100), (100, 50), (150, 10)]
    assert _select(lib, ranges, [1, 1, 1]) == []            # every cut separates the same points: nothing to force
    assert _select(lib, ranges, [1, 3, 1]) == [1]
    assert _select(lib, ranges, [3, 2, 3]) == [0, 2]        # status 2 (not resolvable) is not a forced cut either


def test_nested_cuts_wait_for_the_next_run(pkg):
    lib = pkg.load_library()
    # node 0 = [0, 1000); node 1 = [0, 400) below it; node 2 = [400, 600) below it; node 3 = [2000, 300) elsewhere
    ranges = [(0, 1000), (0, 400), (400, 600), (2000, 300)]
    assert _select(lib, ranges, [3, 3, 3, 3]) == [0, 3]     # the two outermost
    assert _select(lib, ranges, [1, 3, 3, 3]) == [1, 2, 3]  # node 0 keeps its cut: its children are independent of each other
    assert _select(lib, ranges, [1, 3, 1, 1]) == [1]
    # a child with the very range of its parent (the other side empty): the earlier entry counts as the outer one
    assert _select(lib, [(10, 20), (10, 20)], [3, 3]) == [0]


def test_capacity_of_sixteen_forced_cuts(pkg):
    lib = pkg.load_library()
    ranges = [(100 * i, 100) for i in range(16)]
    assert _select(lib, ranges, [3] * 16) == list(range(16))
    assert _select(lib, ranges, [3] * 16, already=14) == [0, 1]   # two slots left
    assert _select(lib, ranges, [3] * 16, already=16) == []       # full: the caller falls back to the ordered re-run


def test_random_forests_never_pick_an_inner_cut(pkg):
    lib = pkg.load_library()
    rng = np.random.default_rng(11)
    for _ in range(200):
        # a random binary partition tree over [0, 4096); pick up to 16 of its nodes
        nodes = [(0, 4096)]
        frontier = [(0, 4096)]
        while frontier and len(nodes) < 60:
            b, s = frontier.pop(int(rng.integers(0, len(frontier))))
            if s < 2:
                continue
            cut = int(rng.integers(1, s))
            for child in ((b, cut), (b + cut, s - cut)):
                nodes.append(child)
                frontier.append(child)
        pick = [nodes[i] for i in rng.permutation(len(nodes))[:16]]
        status = [int(x) for x in rng.choice([1, 2, 3], len(pick))]
        got = _select(lib, pick, status)
        three = [i for i, x in enumerate(status) if x == 3]
        inside = lambda a, b: b[0] <= a[0] and a[0] + a[1] <= b[0] + b[1]
        for i in got:
            assert status[i] == 3
            assert not any(j != i and inside(pick[i], pick[j]) and pick[j][1] > pick[i][1] for j in three)
        for i in three:  # every differing cut is either forced now or lies inside one that is
            assert i in got or any(inside(pick[i], pick[j]) for j in got)
