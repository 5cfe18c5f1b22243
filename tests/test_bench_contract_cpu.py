"""The driver's bench contract, as far as it can be checked without a GPU: `bench.py --impl reference` (the oracle
port timed on the host cores) prints ONE JSON line with the keys the contract names, and exits 0."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--ref-rows", "3000",
                          "--steps", "1", "--warmup", "3"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "samples/sec FM/HOFM predict+grad" and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["vs_baseline"] is None and d["dtype"] == "f64" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_bench_sections_that_do_not_shard_are_guarded():
    """the per-sample solvers refuse to run while a communicator is up (replicas only), so bench.py must not call
    that section under torchrun -- it once aborted every N > 1 line"""
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert "if world == 1 and B.on(\"seq\"):" in src
    assert "if world == 1 and B.on(\"uniform\"):" in src
