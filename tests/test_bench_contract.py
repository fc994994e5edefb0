"""bench.py's reference arm runs on the CPU: check the JSON-line contract (keys, types) without a GPU."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, env=dict(os.environ, RANK="0", WORLD_SIZE="1"))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1 and d["gpu_launches"] == 0
    # "reference" when oracle/_ref holds the staged, unmodified reference files (build() stages them where /root/reference exists;
    # the directory travels to the GPU box), else the labelled port
    from oracle import ref_runner
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_runner.staged() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert "host_numa_bind" in d["config"]
    assert d["e2e"] == {"value": d["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=60, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_clock_sampler_degrades_without_a_gpu_and_aggregates_rows():
    """bench.ClockSampler: NVML polling in-process, `nvidia-smi` as the fallback, a labelled empty record when neither exists;
    the aggregation (median clock, max clock, union of throttle reasons) on hand-made samples."""
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ClockSampler(0)
    s.start()
    out = s.stop()
    assert "sm_mhz" in out and "reasons" in out
    if out.get("sm_mhz") is None:                              # no GPU here: nothing was sampled, and the record says why or is empty
        assert out.get("samples", 0) == 0
    t = bench.ClockSampler(0)
    t.nvml, t.source = object(), "nvml"                         # pretend a poller ran: stop() only joins the thread and aggregates

    class _Done:
        def join(self, timeout=None):
            return None
    t.thread = _Done()
    t.rows = [[1965.0, 1965.0, 300.0, False, False, False, False], [1950.0, 1965.0, 410.0, False, False, False, True],
              [1965.0, 1965.0, 390.0, False, False, False, False]]
    agg = t.stop()
    assert agg["sm_mhz"] == 1965.0 and agg["sm_max_mhz"] == 1965.0 and agg["power_w_max"] == 410.0
    assert agg["samples"] == 3 and agg["reasons"] == ["sw_power_cap"] and agg["source"] == "nvml"
