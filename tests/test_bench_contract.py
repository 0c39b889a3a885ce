"""bench.py contract checks that need no GPU: the reference arm (--impl reference) runs the unmodified reference
CPU function (oracle/_ref/libref.so, or the oracle port where the reference sources are absent) and prints one
JSON line with the keys the driver reads."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "GCUPS" and line["higher_is_better"] is True
    assert line["config"]["workload"] == "cfg2" and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["value"] > 0 and abs(line["value"] - line["e2e"]["value"]) < 1e-9
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"]


def test_workload_table_names_the_baseline_configs():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", ROOT / "bench.py")
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    base = json.load(open(ROOT / "BASELINE.json"))
    assert b.METRIC and "GCUPS" in b.METRIC
    for k in ("cfg1", "cfg2", "cfg3", "cfg4", "cfg5"):
        assert k in b.WORKLOADS or k in getattr(b, "BATCH_WORKLOADS", {}), k
    assert "metric" in base
