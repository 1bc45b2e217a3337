"""bench.py's reference arm on CPU (a small city): one JSON line with the keys the driver reads, rank 0 alone working
under a multi-rank launch."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ["--persons", "6000", "--steps", "2", "--warmup", "1", "--preroll", "1"]


def run_bench(extra, env=None):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *SMALL, *extra],
                         capture_output=True, text=True, timeout=300, env={**os.environ, **(env or {})})
    assert out.returncode == 0, out.stderr[-2000:]
    return [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_prints_one_contract_line():
    lines = run_bench(["--ref-cores", "2"])
    assert len(lines) == 1
    line = lines[0]
    assert line["impl"] == "reference" and line["metric"] == "simulated person-hours/sec" and line["unit"] == "person-hours/s"
    assert line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 1 and line["higher_is_better"] is True
    assert line["dtype"] == "f64" and line["data"] == "synthetic" and line["vs_baseline"] is None and line["gpu_launches"] == 0
    assert "workload" in line["config"] and "N=6000" in line["config"]["workload"]
    base = line["cpu_baseline"]
    assert base["kind"] == "port" and base["cores"] == 2 and base["value"] == line["value"] and "sample" in base
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # value = persons x 24 h x days x replications / wall time of the days; two replications with different seeds
    assert abs(line["value"] - 2 * 6000 * 24.0 * 2 / (line["ms_per_step"] * 2e-3)) < 1e-6 * line["value"]
    assert len(line["infections_per_replication"]) == 2
    assert 0 < base["one_replication"]["value"] <= line["value"]


def test_reference_arm_other_ranks_do_nothing():
    assert run_bench(["--gpus", "2", "--ref-cores", "1"], env={"RANK": "1", "WORLD_SIZE": "2"}) == []
    lines = run_bench(["--gpus", "2", "--ref-cores", "1"], env={"RANK": "0", "WORLD_SIZE": "2"})
    assert len(lines) == 1 and lines[0]["n_gpus"] == 2 and "2 geographic shards" in lines[0]["config"]["workload"]
