"""CPU tests: the oracle (oracle/divquant_oracle.cpp) against the reference's golden vectors.

Pins, in order of authority:
  1. the seven known-answer palettes of the reference's Test/DivQuantTest.m (tests/golden/divquant_kat.json);
  2. outputs of the UNMODIFIED reference compiled here (tests/golden/reference_outputs.npz, produced by
     tests/golden/make_golden.py from oracle/_ref);
  3. when oracle/_ref is present, a live fuzz against it.
"""
import numpy as np
import pytest

from oracle import muted
from conftest import small_case_ids


def test_kat_palettes_both_paths(oracle, kats):
    # DivQuantTest.m asserts numActualClusters and every colortable entry (allPixelsUnique=1);
    # the weighted path (=0) gives the same palettes.
    for kat in kats:
        px = np.array(kat["pixels"], np.uint32)
        for uq in (1, 0):
            with muted():
                out, pal = oracle.quant_recurse(px, kat["k"], uq)
            assert [int(x) for x in pal] == kat["palette"], (kat["name"], uq)
            assert [int(x) for x in out] == kat["out_pixels"], (kat["name"], uq)


def test_small_cases_match_reference_outputs(oracle, golden):
    for i in small_case_ids(golden):
        px, k = golden[f"small{i}_in"], int(golden[f"small{i}_k"][0])
        for uq in (0, 1):
            with muted():
                out, pal = oracle.quant_recurse(px, k, uq)
            assert np.array_equal(pal, golden[f"small{i}_u{uq}_palette"]), (i, uq)
            assert np.array_equal(out, golden[f"small{i}_u{uq}_out"]), (i, uq)
        assert np.array_equal(oracle.map_colors_mps(px, golden[f"small{i}_mappal"]), golden[f"small{i}_mapout"])
        col, w, _ = oracle.calc_color_table(px)
        assert np.array_equal(col, golden[f"small{i}_hist_colours"])
        assert np.array_equal(w, golden[f"small{i}_hist_weights"])  # bit-exact doubles


@pytest.mark.parametrize("name", ["batman", "cookie"])
def test_fixture_images_match_reference(oracle, golden, images, name):
    px = images[name]
    for k in (4, 64, 125, 256):
        with muted():
            out, pal = oracle.quant_recurse(px, k, 0)
        assert np.array_equal(pal, golden[f"{name}_k{k}_palette"]), k
        assert oracle.hash_words(out) == int(golden[f"{name}_k{k}_out_hash"][0]), k
    # the one live call of the reference pipeline: 125-colour grid (ClusteringSegmentation.cpp:397-408)
    assert oracle.hash_words(oracle.map_colors_mps(px, golden["grid125"])) == int(golden[f"{name}_grid125_out_hash"][0])


def test_synthetic_1080p_matches_reference(oracle, golden):
    px = oracle.generate(1, 1920, 1080)
    assert oracle.hash_words(px) == int(golden["g1_1080_k64_in_hash"][0])
    assert np.unique(px & 0xFFFFFF).size == int(golden["g1_1080_k64_unique"][0]) == 118773
    with muted():
        out, pal = oracle.quant_recurse(px, 64, 0)
    assert np.array_equal(pal, golden["g1_1080_k64_palette"])
    assert oracle.hash_words(out) == int(golden["g1_1080_k64_out_hash"][0])


def test_closed_form_remap_equals_pruned_search(oracle):
    # the search result is argmin (dist, rank) over the whole sorted palette.
    rng = np.random.default_rng(11)
    for trial in range(30):
        n = int(rng.integers(1, 20000))
        px = rng.integers(0, 1 << 24, n, dtype=np.uint32)
        k = int(rng.integers(1, 300))
        pal = rng.integers(0, 1 << 24, k, dtype=np.uint32)
        if trial % 3 == 0:  # duplicates and equal sums, where the sort order is observable
            pal[rng.integers(0, k, k // 3 + 1)] = pal[0]
        if trial % 5 == 0:
            vals = [0, 63, 127, 191, 255]
            pal = np.array([(r << 16) | (g << 8) | b for r in vals for g in vals for b in vals], np.uint32)
            rng.shuffle(pal)
        assert np.array_equal(oracle.map_colors_mps(px, pal), oracle.map_colors_mps(px, pal, bruteforce=True))


def test_label_image(oracle):
    pal = np.array([0x101010, 0x202020, 0x101010, 0x303030], np.uint32)
    px = np.array([0x111111, 0x2F2F2F, 0x202020], np.uint32)
    q = oracle.map_colors_mps(px, pal)
    assert list(oracle.colortable_indexes(q, pal)) == [2, 3, 1]  # last duplicate wins (OpenCVUtil.cpp:787-849)


def test_device_arithmetic_model_agrees_on_named_configs(oracle, images):
    """The CUDA path sums count*c exactly (integers) instead of sequentially in doubles.  On the
    BASELINE.json configs that model yields the reference's palette bit for bit; it can differ only
    when a mean sits exactly on an integer / a variance tie, which the reference then decides by its
    own rounding noise ('Floating-point order')."""
    cases = [(images["cookie"], 256), (images["batman"], 64), (oracle.generate(1, 640, 360), 64)]
    for px, k in cases:
        with muted():
            a, _ = oracle.quant_varpart_fast(px, k)
            b, _ = oracle.quant_varpart_fast(px, k, exact_counts=True)
        assert np.array_equal(a, b)


def test_split_statistics_of_model_within_tolerance(oracle, images):
    # north_star: split statistics within 1e-6 relative.
    with muted():
        _, _, ra = oracle.quant_varpart_fast(images["cookie"], 64, with_records=True)
        _, _, rb = oracle.quant_varpart_fast(images["cookie"], 64, with_records=True, exact_counts=True)
    assert len(ra) == len(rb) == 63
    for x, y in zip(ra, rb):
        assert (x.old_index, x.cut_axis, x.num_points, x.new_size) == (y.old_index, y.cut_axis, y.num_points, y.new_size)
        for f in ("cut_pos", "new_weight", "old_weight"):
            assert abs(getattr(x, f) - getattr(y, f)) <= 1e-6 * abs(getattr(x, f))
        for c in range(3):
            assert abs(x.new_mean[c] - y.new_mean[c]) <= 1e-6 * max(1.0, abs(x.new_mean[c]))


def test_live_fuzz_against_compiled_reference(oracle, reference):
    rng = np.random.default_rng(5)
    for trial in range(120):
        n = int(rng.integers(1, 2500))
        mode = trial % 4
        if mode == 0:
            px = rng.integers(0, 1 << 24, n, dtype=np.uint32)
        elif mode == 1:
            base = rng.integers(0, 1 << 24, int(rng.integers(1, 40)), dtype=np.uint32)
            px = base[rng.integers(0, base.size, n)]
        elif mode == 2:
            c = rng.integers(0, 256, 3)
            ch = [np.clip(c[j] + rng.integers(-6, 7, n), 0, 255).astype(np.uint32) for j in range(3)]
            px = (ch[0] << 16) | (ch[1] << 8) | ch[2]
        else:
            px = rng.integers(0, 256, n).astype(np.uint32) * 0x010101
        k = int(rng.choice([1, 2, 3, 4, 8, 16, 17, 64, 125, 256, 300]))
        uq = int(rng.integers(0, 2))
        bits, dec, iters = int(rng.choice([8, 8, 7, 5, 3])), int(rng.choice([1, 1, 2, 3])), int(rng.choice([10, 1, 2, 5]))
        with muted():
            op, oe = oracle.quant_varpart_fast(px, k, bits, dec, iters, uq)
        rp, re_ = reference.quant_varpart_fast(px, k, bits, dec, iters, uq)
        assert np.array_equal(op, rp) and oe == re_, (trial, n, k, uq, bits, dec, iters)
        pal = rng.integers(0, 1 << 24, int(rng.integers(1, 300)), dtype=np.uint32)
        assert np.array_equal(oracle.map_colors_mps(px, pal), reference.map_colors_mps(px, pal))
        ou, ru = oracle.calc_color_table(px, dec), reference.calc_color_table(px, dec)
        assert np.array_equal(ou[0], ru[0]) and np.array_equal(ou[1], ru[1])
        cb = tuple(int(x) for x in rng.integers(1, 9, 3))
        assert np.array_equal(oracle.cut_bits(px, *cb), reference.cut_bits(px, *cb))


def test_srm_sorted_edges_oracle_matches_reference_fixtures(oracle, golden):
    """the oracle's edge list + stable bucket sort against what the unmodified SRM/srm.c built
    (fingerprints and one full list written by tests/golden/make_golden.py)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden import srm_image
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    for name in ("batman", "cookie"):
        z = np.load(os.path.join(here, f"{name}_px.npz"))
        img = srm_image(z["px"].reshape(int(z["shape"][0]), int(z["shape"][1])))
        pairs = oracle.srm_sorted_edges(img)
        assert pairs.shape[0] == int(golden[f"srm_{name}_num_pairs"][0])
        assert oracle.hash_words(pairs.reshape(-1)) == int(golden[f"srm_{name}_pairs_hash"][0])
        if name == "cookie":
            small = np.ascontiguousarray(img[100:137, 200:251])
            assert np.array_equal(oracle.srm_sorted_edges(small), golden["srm_small_pairs"])
