"""bench.py contract, CPU side: the reference arm (the reference's CPU path, timed through oracle/cpu_port.py) prints one
JSON line with the agreed keys.  The GPU arm is exercised on the B200 by the driver."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["n_gpus"] == 1
    assert line["unit"] == "date*stocks/s" and line["value"] > 0 and line["vs_baseline"] is None
    for key in ("metric", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["e2e"]["value"] == line["value"]


def test_bench_source_has_no_rank_local_loops_around_collectives():
    """Regression guard for the multi-GPU deadlock: no time-based `while` loop may drive steps that carry an all-reduce."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert "while time.perf_counter()" not in src
