"""bench.py contract checks that need no GPU: the reference arm runs and prints ONE well-formed JSON line; the
algorithmic byte counts match SURVEY.md section 8(d) (+8 B for the per-env step counter)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_algorithmic_bytes_match_survey():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.algorithmic_bytes(0, 1, 23) == 1209 + 8        # grid-only
    assert bench.algorithmic_bytes(1, 0, 23) == 489 + 8         # genset-only
    assert bench.algorithmic_bytes(1, 1, 23) == 1265 + 8        # genset + grid
    assert bench.algorithmic_bytes(0, 1, 24, discrete=True) == 1245 + 8
    assert bench.algorithmic_bytes(1, 1, 24, discrete=True) == 1285 + 8
    mean = (7 * bench.algorithmic_bytes(0, 1, 23) + 10 * bench.algorithmic_bytes(1, 0, 23) + 8 * bench.algorithmic_bytes(1, 1, 23)) / 25
    assert abs(mean - (938.9 + 8)) < 0.05


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["metric"].startswith("microgrid env-steps/sec") and d["vs_baseline"] is None and d["dtype"] == "f64"


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
