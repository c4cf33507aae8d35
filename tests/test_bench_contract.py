"""bench.py contract checks that need no GPU: the reference arm prints exactly ONE JSON line on stdout (library chatter goes to
stderr), carries the keys the driver reads, and names the same `config` the product arm would print for the same flags."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(300)
def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    env = dict(os.environ, PYTHONPATH=ROOT)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--gpus", "1"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, cwd=ROOT, timeout=280)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    lines = [l for l in p.stdout.decode().splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "decompressed_GB_per_s" and d["unit"] == "GB/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the product arm builds its `config` with the same function from the same flags
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")
    assert d["config"] == bench.workload_config(bench.N_UNITS, bench.DISTINCT, 1)


def test_rank_other_than_zero_of_the_reference_arm_does_nothing():
    env = dict(os.environ, PYTHONPATH=ROOT, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--gpus", "2"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, cwd=ROOT, timeout=120)
    assert p.returncode == 0 and p.stdout.decode().strip() == ""


def test_host_memory_budget_is_positive_and_bounded_by_meminfo():
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")
    b = bench.host_memory_budget()
    total = int([l for l in open("/proc/meminfo") if l.startswith("MemTotal:")][0].split()[1]) * 1024
    assert 0 < b <= total
