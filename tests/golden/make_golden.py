#!/usr/bin/env python
"""Generates the committed golden fixtures under tests/golden/.

Run in the build container, where /root/reference is mounted and oracle/_ref (the UNMODIFIED
reference sources compiled by oracle/Makefile) exists:

    python tests/golden/make_golden.py

Outputs
  divquant_kat.json      the seven known-answer palettes of the reference's Test/DivQuantTest.m
                         (inputs and expected palettes transcribed as data, test:line cited), plus the
                         outPixels the compiled reference produces for them (the reference's tests do
                         not assert those).
  batman_px.npz, cookie_px.npz
                         the two fixture images of the reference (tests/Batman, tests/Cookie) as packed
                         0x00RRGGBB words (cv2.imread(..., IMREAD_COLOR), as SURVEY.md 8c specifies).
  reference_outputs.npz  palettes and output fingerprints of the compiled reference for the
                         BASELINE.json configs and a set of seeded small cases (full outputs).
Nothing here is read at product run time; tests/ and smoke() read the fixtures, never /root/reference.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import Oracle, Reference  # noqa: E402

GREYS = [(i * 25) * 0x010101 for i in range(10)]  # DivQuantTest.m:34-45 (step = 255/10 = 25)
UMBRELLA16 = [0x00EBC58B, 0x00DAD4E7, 0x00D7779D, 0x007E393D, 0x00ABA4BA, 0x00CF4B53, 0x00C49AC7, 0x00AC7292,
              0x00ECEFE7, 0x00DC789D, 0x00A8ABC4, 0x00906E9E, 0x00B54748, 0x00A24F44, 0x00857E77,
              0x007F654B]  # DivQuantTest.m:221-236

KATS = [
    {"name": "testQuantN1", "ref": "Test/DivQuantTest.m:31", "pixels": GREYS, "k": 1, "palette": [0x000000]},
    {"name": "testQuantN2", "ref": "Test/DivQuantTest.m:66", "pixels": GREYS, "k": 2, "palette": [0x323232, 0xAFAFAF]},
    {"name": "testQuantN3", "ref": "Test/DivQuantTest.m:103", "pixels": GREYS, "k": 3,
     "palette": [0x191919, 0xAFAFAF, 0x585858]},
    {"name": "testQuantN3N2", "ref": "Test/DivQuantTest.m:147", "pixels": [0x8DC63F, 0xF26522], "k": 3,
     "palette": [0xF26522, 0x8DC63F]},
    {"name": "testQuantGray2", "ref": "Test/DivQuantTest.m:182", "pixels": [0x0A0A0A, 0xF5F5F5], "k": 2,
     "palette": [0x0A0A0A, 0xF5F5F5]},
    {"name": "testQuant_4x4_N4", "ref": "Test/DivQuantTest.m:218", "pixels": UMBRELLA16, "k": 4,
     "palette": [0xA14D48, 0xC292B3, 0xE6D8C8, 0x96758D]},
    {"name": "testQuant_4x4_N16", "ref": "Test/DivQuantTest.m:263", "pixels": UMBRELLA16, "k": 16,
     "palette": [0x7E393D, 0xD7779D, 0xEBC58B, 0x857E77, 0xDAD4E7, 0xABA4BA, 0xA24F44, 0xAC7292, 0xCF4B53, 0x7F654B,
                 0x906E9E, 0xC49AC7, 0xECEFE7, 0xB54748, 0xA8ABC4, 0xDC789D]},
]


def grid_palette():
    """getSubdividedColors (superpixels/OpenCVUtil.cpp:853-897): 5x5x5 grid over {0,63,127,191,255},
    alpha 0xFF, R outermost and B innermost -- the palette of the one live map_colors_mps call."""
    vals = [0, 63, 127, 191, 255]
    return np.array([0xFF000000 | (r << 16) | (g << 8) | b for r in vals for g in vals for b in vals], np.uint32)


def small_cases(rng, n_cases=48):
    """Seeded small inputs covering the regimes of SURVEY.md 7 (ragged sizes, K>U, greys, few colours)."""
    cases = []
    for i in range(n_cases):
        n = int(rng.integers(1, 1500))
        mode = i % 4
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
        px = px | (rng.integers(0, 2, n, dtype=np.uint32) * np.uint32(0xFF000000))
        k = int(rng.choice([1, 2, 3, 4, 8, 16, 17, 64, 125, 256, 300]))
        cases.append((px.astype(np.uint32), k))
    return cases


CROP_IMAGES = ("batman", "cookie", "g1")   # g1 = oracle.generate(1, 1920, 1080, seed 99)
CROP_DIMS = {"batman": (1000, 1778), "cookie": (1000, 1000), "g1": (1080, 1920)}


def crop_specs(n=36, seed=4242):
    """(image index, y, x, height, width, step, K) of seeded crops; kept when the crop is big enough to have more than
    4096 unique colours in practice (checked by the caller through crop<i>_unique)."""
    rng = np.random.default_rng(seed)
    specs = []
    while len(specs) < n:
        im = int(rng.integers(0, 3))
        h, w = CROP_DIMS[CROP_IMAGES[im]]
        ch, cw = int(rng.integers(150, 700)), int(rng.integers(150, 700))
        y, x = int(rng.integers(0, h - ch + 1)), int(rng.integers(0, w - cw + 1))
        specs.append((im, y, x, ch, cw, int(rng.choice([1, 1, 2])), int(rng.choice([16, 64, 125, 256]))))
    return specs


def crop_pixels(shaped, spec):
    im, y, x, ch, cw, step, _ = (int(v) for v in spec)
    return np.ascontiguousarray(shaped[CROP_IMAGES[im]][y:y + ch:step, x:x + cw:step]).ravel()


def srm_image(px2d):
    """Packed 0x00RRGGBB words (height, width) -> the interleaved B,G,R bytes generateSRM hands to SRM()."""
    return np.stack([px2d & 0xFF, (px2d >> 8) & 0xFF, (px2d >> 16) & 0xFF], axis=-1).astype(np.uint8)


def main():
    import cv2
    o, r = Oracle(), Reference()

    # ---- KATs ----
    kat_out = []
    for kat in KATS:
        px = np.array(kat["pixels"], np.uint32)
        for uq in (1, 0):
            out, pal = r.quant_recurse(px, kat["k"], uq)
            assert [int(x) for x in pal] == kat["palette"], (kat["name"], uq, [hex(x) for x in pal])
        out1, _ = r.quant_recurse(px, kat["k"], 1)
        kat_out.append(dict(kat, out_pixels=[int(x) for x in out1]))
    with open(os.path.join(HERE, "divquant_kat.json"), "w") as f:
        json.dump({"source": "caomw/ClusteringSegmentation-1 Test/DivQuantTest.m; out_pixels from oracle/_ref",
                   "cases": kat_out}, f, indent=1)

    # ---- fixture images ----
    images = {}
    for name, rel in (("batman", "tests/Batman/batman.png"), ("cookie", "tests/Cookie/cookie.png")):
        im = cv2.imread(os.path.join("/root/reference", rel), cv2.IMREAD_COLOR)
        px = (im[..., 2].astype(np.uint32) << 16) | (im[..., 1].astype(np.uint32) << 8) | im[..., 0].astype(np.uint32)
        np.savez_compressed(os.path.join(HERE, f"{name}_px.npz"), px=px, shape=np.array(im.shape[:2]))
        images[name] = px.ravel()

    # ---- reference outputs for the named configs ----
    ref = {}
    grid = grid_palette()
    ref["grid125"] = grid
    for name, px in images.items():
        out, pal = r.quant_recurse(px, 256, 0)
        ref[f"{name}_k256_palette"] = pal
        ref[f"{name}_k256_out_hash"] = np.array([o.hash_words(out)], np.uint64)
        ref[f"{name}_grid125_out_hash"] = np.array([o.hash_words(r.map_colors_mps(px, grid))], np.uint64)
        for k in (4, 64, 125):
            out, pal = r.quant_recurse(px, k, 0)
            ref[f"{name}_k{k}_palette"] = pal
            ref[f"{name}_k{k}_out_hash"] = np.array([o.hash_words(out)], np.uint64)
    for tag, kind, w, h, k in (("g1_1080_k256", 1, 1920, 1080, 256), ("g1_1080_k64", 1, 1920, 1080, 64),
                               ("g1_4k_k256", 1, 3840, 2160, 256), ("g2_640x360_k256", 2, 640, 360, 256)):
        px = o.generate(kind, w, h)
        out, pal = r.quant_recurse(px, k, 0)
        ref[f"{tag}_in_hash"] = np.array([o.hash_words(px)], np.uint64)
        ref[f"{tag}_palette"] = pal
        ref[f"{tag}_out_hash"] = np.array([o.hash_words(out)], np.uint64)
        ref[f"{tag}_unique"] = np.array([np.unique(px & 0xFFFFFF).size], np.uint64)

    # ---- seeded small cases with full outputs (both weighted and uniform paths) ----
    rng = np.random.default_rng(20261018)
    for i, (px, k) in enumerate(small_cases(rng)):
        ref[f"small{i}_in"] = px
        ref[f"small{i}_k"] = np.array([k], np.uint32)
        for uq in (0, 1):
            out, pal = r.quant_recurse(px, k, uq)
            ref[f"small{i}_u{uq}_palette"] = pal
            ref[f"small{i}_u{uq}_out"] = out
        pal_r = rng.integers(0, 1 << 24, int(rng.integers(1, 300)), dtype=np.uint32)
        ref[f"small{i}_mappal"] = pal_r
        ref[f"small{i}_mapout"] = r.map_colors_mps(px, pal_r)
        uq_c, uq_w = r.calc_color_table(px)
        ref[f"small{i}_hist_colours"] = uq_c
        ref[f"small{i}_hist_weights"] = uq_w
    # ---- seeded crops of the fixture images / G1 with more unique colours than the sequential-order kernel takes
    #      (U > 4096): palettes of the compiled reference, for the tolerance test of the exact-integer path ----
    crops = crop_specs()
    ref["crop_specs"] = np.array(crops, np.int64)
    shaped = {"batman": images["batman"].reshape(1000, 1778), "cookie": images["cookie"].reshape(1000, 1000),
              "g1": o.generate(1, 1920, 1080, 99).reshape(1080, 1920)}
    for i, spec in enumerate(crops):
        px = crop_pixels(shaped, spec)
        out, pal = r.quant_recurse(px, int(spec[6]), 0)
        ref[f"crop{i}_palette"] = pal
        ref[f"crop{i}_out_hash"] = np.array([o.hash_words(out)], np.uint64)
        ref[f"crop{i}_unique"] = np.array([np.unique(px & 0xFFFFFF).size], np.uint64)
    # ---- SRM front half: the sorted edge list the unmodified reference builds (SRM/srm.c), as fingerprints ----
    from oracle import ReferenceSRM
    rs = ReferenceSRM()
    for name in ("batman", "cookie"):
        pairs, _ = rs.sorted_edges(srm_image(shaped[name]))
        ref[f"srm_{name}_pairs_hash"] = np.array([o.hash_words(pairs.reshape(-1))], np.uint64)
        ref[f"srm_{name}_num_pairs"] = np.array([pairs.shape[0]], np.uint64)
    small = srm_image(shaped["cookie"][100:137, 200:251])          # full list for a small crop (ragged 37 x 51)
    ref["srm_small_pairs"] = rs.sorted_edges(small)[0]
    np.savez_compressed(os.path.join(HERE, "reference_outputs.npz"), **ref)
    print("golden fixtures written:", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
