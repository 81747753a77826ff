"""The committed bench lines (profiles/r02_bench_*.json, written by bench.py on the B200) carry every key the
driver's contract names; a change to bench.py that drops one shows up here without a GPU."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e"}


def _line(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not committed")
    with open(path) as f:
        lines = [ln for ln in f.read().splitlines() if ln.strip()]
    assert len(lines) == 1, "bench.py prints exactly one JSON line"
    return json.loads(lines[0])


@pytest.mark.parametrize("name,gpus", [("r02_bench_n1.json", 1), ("r02_bench_n2.json", 2), ("r02_bench_n8.json", 8)])
def test_our_arm_line(name, gpus):
    d = _line(name)
    assert BASE_KEYS <= set(d) and {"clocks", "gpu_launches", "roofline"} <= set(d)
    assert d["n_gpus"] == gpus and d["unit"] == "Mpixels/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert d["gpu_launches"] > 0 and d["value"] > 0
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] != d["value"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"]) and d["roofline"]["bound"] in ("hbm", "tensor")
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["parity"] is True and d["pipeline_matches_single_call"] is True
    assert d["parity_detail"]["frames_checked"] >= 7           # every distinct frame of the ring, not frame 0 only
    assert d["ms_per_step"] * d["steps"] >= 500.0              # the timed region is at least half a second
    b = d["batch1080"]                                         # BASELINE config 4, host-fed, at every N
    assert b["every_distinct_frame_matches_reference"] is True and b["value"] > 0 and "1024 frames" in b["workload"]
    assert b["device_resident"]["value"] > b["value"] and b["device_resident"]["last_palette_matches_reference"] is True
    if gpus == 1:
        assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"]) and d["cpu_baseline"]["kind"] in ("reference", "port")
        assert len(d["small_inputs"]["rows"]) == 6 and d["e2e"]["pageable_single_call_ms"] > 0
        assert d["g2_stress"]["parity"] is True and d["g2_stress"]["unique_colours"] > 1_000_000   # SURVEY 8c stress input
        assert d["single_call"]["calls_timed"] >= 32                                                # every ring frame once
    else:
        r = d["row_sharded"]
        assert r["matches_single_gpu_call"] is True and len(r["sizes"]) == 2


def test_reference_arm_line():
    d = _line("r02_bench_n1_reference_arm.json")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    ours = _line("r02_bench_n1.json")
    assert d["config"] == ours["config"] and d["metric"] == ours["metric"] and d["unit"] == ours["unit"]
