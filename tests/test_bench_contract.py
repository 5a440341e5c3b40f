"""bench.py's contract pieces that need no GPU: the `--impl reference` JSON line (one line on stdout, the keys the driver
reads), the loud failure of the CUDA arm without a device, and the clock sampler's time-window summary."""
import importlib.util
import json
import os
import stat
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--cpu-sample-rows", "1500"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    cbl = d["cpu_baseline"]
    # ms_per_step is what was really timed (one query on the bounded sample); value = that rate scaled to the full corpus
    assert d["value"] > 0 and abs(d["value"] * d["ms_per_step"] * cbl["extrapolation_factor"] / 1e3 - 1.0) < 1e-9
    assert cbl["sample_rows"] == 1500 and abs(cbl["extrapolation_factor"] - 10_000_000 / 1500) < 1e-6
    assert cbl["measured_10k"]["scaled"] is False and cbl["measured_10k"]["queries_per_s"] > 0
    assert cbl["measured_dense_1m"]["scaled"] is False and cbl["measured_dense_1m"]["rows"] == 1_000_000   # config 2 context
    assert d["steps"] * d["ms_per_step"] / 1e3 < 60, "the timed region must fit inside the run"
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "1500-row slice" in cb["sample"]
    assert d["config"]["corpus_rows"] == 10_000_000 and d["config"]["top_k"] == 10 and "workload" in d["config"]
    # the reference arm reports on the CUDA arm's config: same keys, same values
    spec = importlib.util.spec_from_file_location("bench_mod_cfg", BENCH)
    bench = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(bench)
        assert d["config"] == bench.workload(bench.parse(), 1)
    finally:
        sys.argv = argv


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_cuda_arm_fails_loudly_without_a_gpu(built_lib):
    from b200rag import _ffi
    if _ffi.device_count() > 0:
        return                                                        # on a GPU box the arm runs (the driver does that)
    r = subprocess.run([sys.executable, BENCH, "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "no CPU fallback" in r.stderr


def test_clock_sampler_summarises_only_the_timed_window(tmp_path, monkeypatch):
    fake = tmp_path / "nvidia-smi"
    fake.write_text("#!/bin/bash\nsleep 0.1\ni=0\nwhile true; do i=$((i+1));\n"
                    "if [ $i -le 4 ]; then echo \"300, 1965, 80.0, 0x0, Not Active, Not Active, Not Active, Not Active\";\n"
                    "else echo \"1900, 1965, 700.5, 0x4, Not Active, Not Active, Not Active, Active\"; fi; sleep 0.05; done\n")
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", f"{tmp_path}:{os.environ['PATH']}")
    spec = importlib.util.spec_from_file_location("bench_mod", BENCH)
    bench = importlib.util.module_from_spec(spec)
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    spec.loader.exec_module(bench)
    s = bench.ClockSampler("GPU-fake")
    s.start()
    time.sleep(0.6)                      # idle samples (300 MHz) arrive first and must not be summarised
    t0 = time.perf_counter()
    time.sleep(0.4)
    out = s.stop(t0, time.perf_counter())
    assert out["samples"] >= 3 and out["sm_mhz"] == 1900.0 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"] and out["power_w_max"] == 700.5
    assert s.proc.poll() is not None     # the sampler process is gone
    none = bench.ClockSampler("GPU-fake")
    assert none.stop()["reasons"] == ["nvidia-smi unavailable"]
