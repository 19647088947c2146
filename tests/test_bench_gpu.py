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
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--points", "50000", "--queries", "2000", "--steps", "3",
                        "--warmup", "3", "--recall-queries", "500", "--extra", "c5a,c3,c5b",
                        "--extra-n", "c5a=40000,c3=30000,c5b=60000"], capture_output=True, text=True, timeout=900)
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
    par = d["config"]["parity"]
    assert par["id_rows_identical_to_oracle"] == 1.0 and par["dists_bit_identical"] is True and par["queries"] == 2000
    assert d["config"]["recall_at_10"] >= 0.95
    assert d["roofline"]["kernel_ms"] > 0 and d["roofline"]["kernel_ms"] <= d["ms_per_step"] * 1.05
    # the other BASELINE configs measured in the same run (reduced sizes here), each with its own
    # oracle parity probe on the GPU-built graph
    ex = d["extra_configs"]
    fl = ex["flat"]  # K5 on the headline shard: tensor-core path, lists equal to the exact scan's
    assert "error" not in fl, fl
    assert fl["path"] == "tcgen05 candidate pass" and fl["lists_identical_to_exact_scan"]["identical"] is True
    assert fl["ms_per_batch"] > 0 and fl["roofline"]["bound"] == "tensor" and fl["candidates_per_query"] >= 10
    assert ex["c5a"]["points_per_s"] > 0 and ex["c5a"]["recall_at_10_of_built_graph"] >= 0.95
    assert ex["c5a"]["insert_stats"]["points"] == 40000 and 0 < ex["c5a"]["roofline"]["frac"] < 1.5
    for name in ("c3", "c5b"):
        e = ex[name]
        assert "error" not in e and "skipped" not in e, e
        assert e["parity"]["id_rows_identical_to_oracle"] == 1.0 and e["parity"]["dists_bit_identical"] is True, e["parity"]
        assert e["recall_at_10_merged_tie_aware"] >= 0.95 and e["shard_searches_per_s"] > 0
        assert 0 < e["roofline"]["frac"] < 1.5
