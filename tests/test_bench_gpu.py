"""bench.py's own arm at a reduced size on the GPU box: one JSON line, the contract keys, parity
with the oracle reported by the bench itself."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_bench_line_contract_small():
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--n", "50000", "--queries", "2000", "--steps", "3",
                        "--warmup", "3", "--recall-queries", "500"], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in d, key
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["value"] > 0 and d["gpu_launches"] >= 3
    assert d["roofline"]["bound"] == "hbm" and 0 < d["roofline"]["frac"] < 1.5 and d["roofline"]["unit"] == "GB/s"
    assert d["e2e"]["h2d_bytes_per_step"] == 2000 * 128 * 4 and d["e2e"]["d2h_bytes_per_step"] == 2000 * 10 * 12 + 2000 * 4
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert d["config"]["parity"] == {"id_rows_identical_to_oracle": 1.0, "dists_bit_identical": True}
    assert d["config"]["recall_at_10"] >= 0.95
