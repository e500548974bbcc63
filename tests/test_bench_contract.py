"""bench.py's CPU side: the reference arm prints ONE JSON line with the keys the driver reads, the standalone-executable
baseline runs, and the helper that bounds the pinned e2e buffers returns something sensible.  No GPU."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

HAVE_REF = (ROOT / "oracle" / "_ref" / "liblgar_ref.so").exists() or (ROOT / "oracle" / "liblgar_oracle.so").exists()


@pytest.mark.skipif(not HAVE_REF, reason="no CPU oracle built")
def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-cols-per-core", "2",
                        "--hours-per-step", "24", "--spinup-hours", "24"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "column-timesteps/sec" and line["unit"] == "column-timesteps/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["steps"] == 1 and line["warmup"] == 0
    assert line["config"]["workload"] == "config5-shard" and line["dtype"] == "f64"
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.skipif(not (ROOT / "oracle" / "_ref" / "lasam_standalone").exists(), reason="reference executable not built")
def test_standalone_executable_baseline_runs():
    v = bench.run_standalone_exe(2, 2, 72)
    assert v is not None and v > 0
    e = bench.standalone_exe_entry(1, 1, 48)
    assert e["kind"] == "reference" and e["cores"] == 1 and e["value"] > 0


def test_host_memory_helper():
    m = bench.host_memory_bytes()
    assert 1 << 28 < m < 1 << 50


@pytest.mark.skipif(not (ROOT / "oracle" / "_ref" / "liblgar_ref.so").exists(), reason="reference harness not built")
def test_parity_sampling_tool_selftest(tmp_path):
    """tools/parity_sampling_full.py with the unmodified reference standing in for the engine: the bookkeeping of the
    full-duration parity report (forcing of configs 3/4/5 by global column id, per-step diff, final fronts) must find nothing."""
    out = tmp_path / "pf.json"
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "parity_sampling_full.py"), "--selftest", "--sample", "3", "--steps", "150",
                        "--cores", "2", "--configs", "3,4,5", "--out", str(out)], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    rep = json.loads(out.read_text())
    for name in ("config3", "config4", "config5"):
        c = rep["configs"][name]
        assert c["column_steps_diffed"] == 450 and c["front_count_mismatch_columns"] == 0 and c["failure_mismatch_columns"] == 0
        assert c["final_front_flag_mismatch_columns"] == 0 and c["final_front_field_mismatch_columns"] == 0
        assert all(v["columns_outside_bar"] == 0 for v in c["scalars"].values())


def test_committed_ncu_capture_belongs_to_the_sources():
    """bench.py prints the measured DRAM traffic of profiles/r2_window_profile.json only when the capture was taken with the
    sources liblgarb.so is built from (kernel_source_sha256): a kernel change without a new capture must show as `traffic: null`,
    never as stale numbers.  Here: the committed capture is the one of the committed sources, and it carries what the line needs."""
    import bench
    p = json.loads((ROOT / "profiles" / "r2_window_profile.json").read_text())
    for key in ("source_sha256", "dram_bytes_per_column_timestep", "stage_time_share", "launches", "columns", "hours"):
        assert key in p, key
    assert abs(sum(p["stage_time_share"].values()) - 1.0) < 1e-9
    traffic, share, source = bench.committed_window_profile()
    assert p["source_sha256"] == bench.kernel_source_sha256(), "kernel sources changed after the ncu capture: re-capture (tools/prof_window.sh, tools/make_window_profile.py)"
    assert traffic == p["dram_bytes_per_column_timestep"] and share and "r2_window_profile.json" in source
