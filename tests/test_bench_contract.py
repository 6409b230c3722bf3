"""CPU: the reference arm of bench.py honours the output contract (exactly one JSON line on
stdout, the required keys) -- the part of the contract that can be checked without a GPU."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*extra, env=None):
  e = dict(os.environ)
  e.update(env or {})
  return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                         "--warmup", "1", "--cpu-sample-n", "320", "--d", "32", *extra],
                        capture_output=True, text=True, cwd=ROOT, env=e, timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
  res = run_bench("--cpu-stagewise-n", "256")
  assert res.returncode == 0, res.stderr[-2000:]
  lines = [l for l in res.stdout.split("\n") if l.strip()]
  assert len(lines) == 1, res.stdout
  line = json.loads(lines[0])
  for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
              "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
    assert key in line, key
  assert line["impl"] == "reference" and line["steps"] == 2 and line["warmup"] == 1
  assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
  assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
  assert set(line["cpu_stagewise"]["seconds"]) >= {"affinity", "gaussian_blur", "row_threshold", "diffuse"}
  assert line["vs_baseline"] is None and line["higher_is_better"] is True


def test_reference_arm_other_ranks_exit_quietly():
  """Under torchrun (N > 1) rank 0 alone runs the CPU arm; the other ranks print nothing."""
  res = run_bench("--gpus", "2", "--cpu-stagewise-n", "0", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
  assert res.returncode == 0, res.stderr[-2000:]
  assert res.stdout.strip() == ""
