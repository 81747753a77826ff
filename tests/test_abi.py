"""CPU tests of the drop-in boundary: the shared library loads, exports every symbol that include/*.h
declares plus the reference's own mangled names, and its host-only palette logic matches the oracle.
No device work is done here (the library aborts without a GPU by design)."""
import os
import re
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _exported(lib_path):
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], check=True, capture_output=True, text=True).stdout
    return {line.split()[-1] for line in out.splitlines() if line.strip()}


def test_library_exports_every_declared_symbol(pkg):
    header = open(os.path.join(ROOT, "include", "divquant_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(dq_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    syms = _exported(pkg.LIB_PATH)
    assert declared <= syms, declared - syms
    assert declared == set(pkg.EXPORTED_C_SYMBOLS)
    lib = pkg.load_library()  # dlopen + every prototype bound
    assert b"sm_100a" in lib.dq_version()


def test_library_exports_reference_symbol_names(pkg):
    # the reference's callers link against these exact names.
    syms = _exported(pkg.LIB_PATH)
    assert set(pkg.REFERENCE_SYMBOLS) <= syms, set(pkg.REFERENCE_SYMBOLS) - syms


def test_compat_headers_compile_reference_style_caller(pkg, tmp_path):
    # a caller written against the reference's headers compiles and links against the B200 library
    src = tmp_path / "caller.cpp"
    src.write_text('#include "DivQuantHeader.h"\n#include "quant_util.h"\n'
                   "int main(int argc, char**) { if (argc > 99) { uint32_t in[2] = {1, 2}, out[2], k = 2, ct[2];\n"
                   " quant_recurse(2, in, out, &k, ct, 1); map_colors_mps(in, 2, out, ct, 2);\n"
                   " int n; double *w = calc_color_table(in, 2, out, 1, 2, 1, &n); delete[] w; }\n"
                   " return validate_num_bits(8) ? 0 : 1; }\n")
    exe = tmp_path / "caller"
    libdir = os.path.dirname(pkg.LIB_PATH)
    subprocess.run(["g++", "-std=c++11", "-I", pkg.INCLUDE_DIR, str(src), "-o", str(exe), "-L", libdir,
                    "-ldivquant_b200", f"-Wl,-rpath,{libdir}"], check=True)
    assert subprocess.run([str(exe)]).returncode == 0  # no device call is made with argc == 1


def test_cuda_code_is_sm_100a_only(pkg):
    out = subprocess.run(["cuobjdump", "-lelf", pkg.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_host_palette_logic_matches_oracle(pkg, oracle):
    dq = pkg.DivQuant.__new__(pkg.DivQuant)
    dq.lib = pkg.load_library()
    rng = np.random.default_rng(2)
    for trial in range(200):
        k = int(rng.integers(1, 400))
        pal = rng.integers(0, 1 << 24, k, dtype=np.uint32)
        if trial % 2:
            pal[rng.integers(0, k, k // 2 + 1)] = pal[rng.integers(0, k)]
        if trial % 5 == 0:  # many equal sums: the std::sort permutation is observable
            pal = (rng.integers(0, 4, (k, 3)) * 64).astype(np.uint32)
            pal = (pal[:, 0] << 16) | (pal[:, 1] << 8) | pal[:, 2]
        s1, l1 = dq.host_build_search_tables(pal)
        s2, l2 = oracle.build_search_tables(pal)
        assert np.array_equal(s1, s2) and np.array_equal(l1, l2)
        d = dq.host_dedup_palette(pal)
        _, first = np.unique(pal, return_index=True)
        assert np.array_equal(d, pal[np.sort(first)])
