"""CPU check of bench.py's reference arm (the oracle port on host cores): it must
print ONE JSON line with the contract's keys, on a bounded sample."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--steps", "3", "--warmup", "1", "--size", "48"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "cg_iterations_per_sec" and d["unit"] == "it/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["vs_baseline"] is None and d["dtype"] == "f64" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and d["value"] > 0
    # the arm runs the workload it names (no extrapolation from a smaller grid) ...
    assert d["config"]["extrapolated"] is False and "48^3" in d["config"]["sample"]
    assert d["steps"] == 3


def test_reference_arm_does_not_map_the_cuda_library():
    """... and touches nothing of the product: krylov_b200 (whose import loads
    libkrylov_b200.so) must not be imported by the CPU arm."""
    code = (
        "import sys, runpy\n"
        "sys.argv = ['bench.py', '--impl', 'reference', '--size', '24', '--steps', '2', '--warmup', '1']\n"
        "try:\n"
        f"    runpy.run_path({os.path.join(ROOT, 'bench.py')!r}, run_name='__main__')\n"
        "except SystemExit:\n"
        "    pass\n"
        "maps = open('/proc/self/maps').read()\n"
        "print('MAPPED' if 'libkrylov_b200' in maps else 'CLEAN', 'krylov_b200' in sys.modules)\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip().splitlines()[-1] == "CLEAN False"


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--gpus", "2", "--steps", "2", "--size", "32"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
