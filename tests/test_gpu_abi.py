"""The drop-in boundary EXECUTED on the GPU: the reference's own unit tests for the hot path (Test/DivQuantTest.m:31-316)
as a C++ program compiled against the reference-named headers and linked with -ldivquant_b200 (tests/abi/divquant_kat.cpp),
i.e. what a maintainer who relinks the reference gets.  Also checks the two timing lines quant_recurse prints on stdout
(quant_util.cpp:62-66, 141-145)."""
import json
import os
import subprocess

import pytest

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu


def _kat_header(path):
    with open(os.path.join(GOLDEN, "divquant_kat.json")) as f:
        cases = json.load(f)["cases"]
    lines = ["#include <stdint.h>", "struct Kat { const char *name; int num_pixels, k, num_palette; const uint32_t *pixels, *palette, *out_pixels; };"]
    for i, c in enumerate(cases):
        for key in ("pixels", "palette", "out_pixels"):
            vals = ", ".join(f"0x{int(v):06X}u" for v in c[key])
            lines.append(f"static const uint32_t kat{i}_{key}[] = {{{vals}}};")
    rows = [f'  {{"{c["name"]}", {len(c["pixels"])}, {c["k"]}, {len(c["palette"])}, kat{i}_pixels, kat{i}_palette, kat{i}_out_pixels}}'
            for i, c in enumerate(cases)]
    lines.append("static const Kat kKats[] = {\n" + ",\n".join(rows) + "\n};")
    lines.append(f"static const int kNumKats = {len(cases)};")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    return len(cases)


def test_reference_kats_through_the_reference_symbols(pkg, tmp_path):
    n = _kat_header(tmp_path / "kat_data.h")
    assert n == 7
    exe = tmp_path / "divquant_kat"
    libdir = os.path.dirname(pkg.LIB_PATH)
    subprocess.run(["g++", "-std=c++11", "-O1", "-I", pkg.INCLUDE_DIR, "-I", str(tmp_path), os.path.join(ROOT, "tests", "abi", "divquant_kat.cpp"),
                    "-o", str(exe), "-L", libdir, "-ldivquant_b200", f"-Wl,-rpath,{libdir}"], check=True)
    env = dict(os.environ)
    env.pop("DIVQUANT_B200_TIMINGS", None)
    run = subprocess.run([str(exe)], capture_output=True, text=True, env=env, timeout=300)
    assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-2000:]
    assert "divquant_kat: 0 failure(s)" in run.stdout
    # 7 cases x 2 weight paths, two lines each, like the reference
    assert run.stdout.count("quant_varpart_fast() elapsed:") == 14
    assert run.stdout.count("map_colors_mps() elapsed:") == 14
    assert "# empty clusters: 1" in run.stderr  # testQuantN3N2 (DivQuantTest.m:147): 2 colours, K = 3
