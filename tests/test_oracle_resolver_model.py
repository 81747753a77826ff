"""CPU test: the oracle's model of the device's default large-input path (exact integer sums + tie audit + resolver,
oracle.quant_varpart_resolved) equals the compiled reference on the frames where the plain exact-integer model does not:
config-4 seeds 12410 / 12830 / 13124 (a mean exactly on x.5: one palette entry 1 LSB off without the resolver) and 12832 (a
mean exactly on an integer that the reference's rounding noise puts on the other side: the cut is forced).  The GPU tests
(tests/test_gpu_frames.py) pin the device itself against the reference; this pins the METHOD without a GPU.
tools/model_check.py runs the same comparison over all 1024 + 256 frames (profiles/r02_model_check.txt)."""
import os

import numpy as np
import pytest

from oracle import muted

HAVE_REF = os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref",
                                       "libdivquant_ref.so"))


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built")
@pytest.mark.parametrize("seed,kind", [(12410, "rounding"), (12830, "rounding"), (13124, "rounding"), (12832, "cut"), (12345, "clean")])
def test_resolver_model_equals_reference(oracle, seed, kind):
    from oracle import Reference
    ref = Reference()
    px = oracle.generate(1, 1920, 1080, seed)
    with muted():
        ref_pal, _ = ref.quant_varpart_fast(px, 64)
        plain, _ = oracle.quant_varpart_fast(px, 64, exact_counts=True)
    pal, info = oracle.quant_varpart_resolved(px, 64)
    assert info["left"] == 0
    assert np.array_equal(pal, ref_pal)
    if kind == "rounding":
        assert not np.array_equal(plain, ref_pal) and info["roundings"] >= 1   # the reason the audit exists
        assert int((plain != ref_pal).sum()) == 1
    elif kind == "cut":
        assert info["cuts_forced"] == 1
    else:
        assert np.array_equal(plain, ref_pal) and not any(info.values())
