"""The driver-facing contract of bench.py that can be checked without a GPU: the reference arm runs on the host cores
and prints ONE JSON line with the agreed keys; the B200 arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          cwd=ROOT, timeout=600)


def test_reference_arm_prints_one_json_line():
    r = _run("--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["value"] > 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps",
                        "1", "--warmup", "1"], capture_output=True, text=True, cwd=ROOT, timeout=120, env=env)
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_b200_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return  # meaningful only in the CPU container
    r = _run("--steps", "1", "--warmup", "1")
    assert r.returncode != 0 and "CUDA" in (r.stderr + r.stdout)


def test_experiment_hooks_stay_out_of_the_product_path():
    """NRSE_EXPERIMENT (kernels that skip work for timing) and NRSE_B200_LIB (another build of the library) exist for
    scripts/ only: the package never sets them, bench.py and smoke() refuse to run with them, and only _lib.py reads the
    library override."""
    import glob
    pkg = os.path.join(ROOT, "noise-robust-speech-embedding_b200")
    for path in glob.glob(os.path.join(pkg, "**", "*.py"), recursive=True):
        src = open(path).read()
        if not path.endswith(os.path.join("csrc", "build.py")):  # build.py names the -DNRSE_EXPERIMENTS switch of its scripts-only variant
            assert "NRSE_EXPERIMENT" not in src, path
        if not path.endswith("_lib.py") and not path.endswith(os.path.join("csrc", "build.py")):
            assert "NRSE_B200_LIB" not in src, path
    for name in ("bench.py", "__graft_entry__.py"):
        src = open(os.path.join(ROOT, name)).read()
        assert 'for var in ("NRSE_EXPERIMENT", "NRSE_B200_LIB")' in src, name
        assert "nrse_experiments_build" in src, name
    # the shipped binary itself: the hooks are compiled in only with -DNRSE_EXPERIMENTS (csrc/build.py --experiments,
    # a separate scripts-only library); the product library never reads the environment variable
    lib = os.path.join(pkg, "csrc", "libnrse_b200.so")
    if os.path.exists(lib):
        assert b"NRSE_EXPERIMENT" not in open(lib, "rb").read()


def test_clock_sampler_parses_time_stamped_samples():
    """bench.py samples nvidia-smi during the timed region (the sampler is started before the warm-up and its samples are
    time-stamped, so its start-up neither perturbs nor pollutes the region): the line parser on a captured line."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("nrse_bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    line = "2026/10/18 17:03:21.123, 0, 1485, 1965, 615.03, 0x0000000000000004, Not Active, Not Active, Not Active, Active"
    ts, sm, smax, power, reasons = bench.ClockSampler.parse_line(line)
    assert (sm, smax, power, reasons) == (1485.0, 1965.0, 615.03, ["sw_power_cap"]) and ts is not None
    assert bench.ClockSampler.parse_line("N/A, 0, [N/A]") is None
    s = bench.ClockSampler(0)
    assert s.stop()["reasons"] == ["nvidia-smi unavailable"]   # never started: reported, not raised
