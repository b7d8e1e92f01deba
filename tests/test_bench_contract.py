"""bench.py pieces that run without a GPU: the roofline's algorithmic bytes are the figure
SURVEY.md section 8(d) states, and the reference arm (`--impl reference`: the C restatement of
the reference's arithmetic on the host cores) prints one JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

from conftest import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_algorithmic_bytes_match_the_survey():
  # config 2: 3.02 GB per sweep, 9.45 bytes per nnz*R*iter
  b = bench.algorithmic_bytes_per_sweep(1000000, 500000, 10000000, 32)
  assert abs(b / 1e9 - 3.024) < 0.01 and abs(b / (10000000 * 32) - 9.45) < 0.01
  # config 5: 8.54 bytes per nnz*R*iter
  b5 = bench.algorithmic_bytes_per_sweep(65000000, 1000000, 1800000000, 32)
  assert abs(b5 / (1800000000 * 32) - 8.54) < 0.01


def test_reference_arm_prints_the_contract_line(tmp_path):
  env = dict(os.environ, HGE_CACHE_DIR=str(tmp_path))
  out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--workload", "mini", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=600)
  assert out.returncode == 0, out.stderr[-2000:]
  lines = [l for l in out.stdout.splitlines() if l.strip()]
  assert len(lines) == 1
  d = json.loads(lines[0])
  assert d["impl"] == "reference" and d["metric"] == "alg-dist incidence nnz*R*iters/sec"
  assert d["unit"] == "nnz*R*iters/s" and d["higher_is_better"] is True and d["value"] > 0
  assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
  assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
  assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
  assert d["config"]["workload"] and d["vs_baseline"] is None


def test_both_arms_print_the_same_config_object():
  """`config` names the workload and the cache policy only, built by one function for both arms
  (ours and --impl reference) at every N; run-specific figures live under `detail`."""
  import bench
  spec = bench.WORKLOADS["c2"]
  one = bench.workload_config(spec, 1, 1000000, 500000, 32, 20)
  assert set(one) == {"workload", "nodes", "edges", "R", "sweeps", "seed", "l2"}
  assert one["workload"] == spec["name"] and one["nodes"] == 1000000
  eight = bench.workload_config(spec, 8, 1000000, 500000, 32, 20)
  assert eight["nodes"] == 8000000 and eight["workload"].startswith("8 x [")
  src = open(os.path.join(ROOT, "bench.py")).read()
  assert src.count('"config": workload_config(') == 3      # reference arm, one GPU, N GPUs


def test_other_ranks_of_the_reference_arm_stay_silent(tmp_path):
  env = dict(os.environ, HGE_CACHE_DIR=str(tmp_path), RANK="1", WORLD_SIZE="2")
  out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--workload", "mini", "--gpus", "2"], capture_output=True, text=True, env=env,
                       timeout=120)
  assert out.returncode == 0 and out.stdout.strip() == ""
