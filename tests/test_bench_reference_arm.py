"""The reference arm of bench.py (the CPU restatement timed on the host cores) runs without a GPU and keeps the bench contract:
one JSON line with impl / metric / unit / cpu_baseline / e2e; at N > 1 it runs N independent sequences on N host threads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(gpus):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", str(gpus), "--steps", "3", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT, stdin=subprocess.DEVNULL)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("gpus", [1, 2])
def test_reference_arm_line(gpus):
    d = _run(gpus)
    assert d["impl"] == "reference" and d["metric"] == "scan_to_map_scans_per_s" and d["unit"] == "scans/s"
    assert d["n_gpus"] == gpus and d["steps"] == 3 and d["warmup"] == 1 and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] == min(gpus, os.cpu_count())
    assert d["config"]["sequences"] == gpus
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
