#!/usr/bin/env python
"""Headline benchmark: algebraic-distance relaxation throughput (incidence nnz * R * iters / s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" is one full relaxation (load, `sweeps` x (node half, edge half, fused rescale),
store) of the BASELINE.json configs[1] workload: synthetic power-law hypergraph, 1M nodes /
500K edges / ~10M incidences, R = 32 trial columns, 20 sweeps.  Prints ONE JSON line.

`value`     device-resident: vectors and incidence already in HBM when the timed region starts.
`e2e`       the same step through the host-buffer C-ABI call (hge_algdist_run_csr with
            HGE_MEM_HOST: what EmbedAlgebraicDistance makes per hypergraph): pinned host -> device
            copies of the incidence arrays and the initial vectors and the device -> host copy of
            the result are inside the timed region.  N > 1: ShardedRelaxation from host buffers,
            with the per-step times, their median and the measured host-link rates beside it.
`parity`    the result of the timed configuration against the f64 C port of the reference's
            arithmetic (N = 1: full size; N > 1: an N-rank relaxation against oracle/port.py).
`roofline`  the half-sweep kernel: algorithmic bytes per launch / its mean launch
            duration measured with CUDA events on the launch stream, against the measured HBM
            copy peak in MEASURED_PEAKS.json; next to it the DRAM traffic of the committed ncu
            capture (while the kernel's sources still hash to it), the compulsory bytes, and the
            gather-only roof of the access pattern.  N > 1: the phase split of a sweep.
`cpu_baseline` the oracle port (oracle/algdist_ref.c: the reference's per-row f64 arithmetic, all
            host threads) on the same workload, timed on this box's host cores.  Reported, not
            the target.
`config`    the workload and the cache policy only -- the same object in both arms at every N;
            run-specific figures (incidence counts, exchange strategy) are under `detail`.
`hobe`, `hobe_scale`, `fobe_scale`, `hg2v_train`, `c1_end_to_end`, `pair_weighting`, `c5`
            the other half of BASELINE.json's metric (HOBE / FOBE samples/s on configs[0] and at
            100 000 - 1 000 000 nodes, with oracle digests), the consumer of the samples,
            configs[0] end to end, the 100M-pair weighting of configs[2] and configs[4]
            (strong scaling at N > 1), each with its own parity check.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]
    "c2": dict(kind="power_law", num_nodes=1000000, num_edges=500000, num_incidences=10000000,
               R=32, sweeps=20, seed=1234,
               name="synthetic power-law hypergraph 1M nodes / 500K edges / 10M incidences, R=32, 20 sweeps"),
    # BASELINE.json configs[4]: strong scaling, generated on the device shard by shard
    "c5": dict(kind="community", num_nodes=65000000, num_edges=1000000, num_incidences=1800000000,
               R=32, sweeps=20, seed=99,
               name="Friendster-community-shaped synthetic hypergraph 65M nodes / 1M edges / 1.8B incidences, R=32, 20 sweeps"),
    "c5mini": dict(kind="community", num_nodes=650000, num_edges=10000, num_incidences=18000000,
                   R=32, sweeps=20, seed=99, name="1/100-scale config 5 (debug)"),
    # BASELINE.json configs[3]: FOBE sample generation at 10M nodes (host RNG replay; side line)
    "c4": dict(kind="fobe", num_nodes=10000000, num_edges=5000000, seed=2024, k=5, num_samples=200,
               name="FOBE (HG2V_BOOLEAN) sample generation, synthetic Zipf hypergraph 10M nodes / 5M edges, "
                    "num_neighbors=5, num_samples=200"),
    "c4mini": dict(kind="fobe", num_nodes=1000000, num_edges=500000, seed=2024, k=5, num_samples=200,
                   name="1/10-scale config 4 (debug)"),
    # small variant for quick checks (not a bench line)
    "mini": dict(kind="power_law", num_nodes=100000, num_edges=50000, num_incidences=1000000,
                 R=32, sweeps=20, seed=1234, name="1/10-scale config 2 (debug)"),
}

FALLBACK_HBM_GBS = 6650.0   # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def log(*a):
  print(*a, file=sys.stderr, flush=True)


def build_workload(spec, cache=True):
  from hypergraphembedding_b200 import synthetic
  key = "hge_%s_%d_%d_%d_%d.npz" % (spec["kind"], spec["num_nodes"], spec["num_edges"],
                                   spec["num_incidences"], spec["seed"])
  path = os.path.join(os.environ.get("HGE_CACHE_DIR", "/tmp"), key)
  import scipy.sparse as sps
  if cache and os.path.exists(path):
    z = np.load(path)
    A = sps.csr_matrix((np.ones(len(z["indices"]), dtype=bool), z["indices"], z["indptr"]),
                       shape=tuple(z["shape"]))
  else:
    t = time.time()
    A = synthetic.power_law_hypergraph(spec["num_nodes"], spec["num_edges"], spec["num_incidences"],
                                       seed=spec["seed"])
    log("generated %s in %.1f s" % (spec["name"], time.time() - t))
    if cache:
      try:
        np.savez(path, indices=A.indices, indptr=A.indptr, shape=np.asarray(A.shape))
      except OSError:
        pass
  B = A.T.tocsr()
  B.sort_indices()
  return A, B


def algorithmic_bytes_per_sweep(N, E, nnz, R):
  """each half-sweep gathers one R-row (4R bytes) and reads one column
  id (4 bytes) per incidence and reads + writes each owned row once."""
  return 2 * nnz * (4 * R + 4) + 2 * (N + E) * 4 * R


def compulsory_bytes_per_launch(N, E, nnz, R):
  """DRAM bytes a half-sweep cannot avoid, averaged over the two halves:
  one 4-byte stream slot per incidence, every owned row read and written once, and the gathered
  table read from DRAM once (it was written by the previous launch and does not fit L2 next to
  the streamed data).  The algorithmic figure counts the gathered row once per INCIDENCE
  instead -- the difference is what L2 serves."""
  return (2 * nnz * 4 + 3 * (N + E) * 4 * R) / 2.0


# the sources that define the half-sweep kernel and the gather stream it reads
HALF_SWEEP_SOURCES = ("hge_sweep.cu", "hge_sweep.cuh", "hge_schedule.cu", "hge_incidence.cuh", "hge_common.cuh")


def kernel_source_sha(files=HALF_SWEEP_SOURCES):
  """Hash of the CUDA sources of a kernel: an ncu capture is only quoted next to the kernel it
  measured."""
  import hashlib
  h = hashlib.sha256()
  for name in sorted(files):
    h.update(open(os.path.join(ROOT, "hypergraphembedding_b200", "csrc", name), "rb").read())
  return h.hexdigest()[:16]


def parity_against(A, xn, xe, ref_xn, ref_xe, R):
  """The north-star bar on a relaxation result: distance of every stored incidence against the
  reference arithmetic (f64), |d - d_ref| <= 1e-5 d_ref + 1e-6 sqrt(R); HOBE incidence weights
  (sqrt(R) - d) / sqrt(R) alongside.  Returns the worst ratio to the bound (<= 1 passes)."""
  A = A.tocoo()
  worst = worst_abs = 0.0
  step = 1 << 20
  for i in range(0, A.nnz, step):
    r, c = A.row[i:i + step], A.col[i:i + step]
    d = np.sqrt(((xn[r].astype(np.float64) - xe[c].astype(np.float64))**2).sum(axis=1))
    dr = np.sqrt(((ref_xn[r].astype(np.float64) - ref_xe[c].astype(np.float64))**2).sum(axis=1))
    err = np.abs(d - dr)
    worst = max(worst, float((err / (1e-5 * dr + 1e-6 * np.sqrt(R))).max()))
    worst_abs = max(worst_abs, float(err.max()))
  coord = max(float(np.abs(xn - ref_xn).max()), float(np.abs(xe - ref_xe).max()))
  return {"max_dist_err_over_bound": worst, "max_dist_abs_err": worst_abs,
          "max_weight_err": worst_abs / float(np.sqrt(R)), "max_coord_abs_err": coord,
          "n_incidences": int(A.nnz), "bound": "|d - d_ref| <= 1e-5 d_ref + 1e-6 sqrt(R)",
          "ok": bool(worst <= 1.0)}


def hbm_peak():
  path = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(path):
    try:
      return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
      pass
  return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload, key="half_sweep"):
  """dram__bytes_read.sum + dram__bytes_write.sum per launch of the named kernel from the
  committed `ncu --set full` capture (profiles/<key>_traffic.json, written by
  tools/ncu_traffic.py).  None when there is no capture of this workload, or when the capture
  was taken from other kernel sources than the ones in this tree (the file carries their hash)."""
  path = os.path.join(ROOT, "profiles", key + "_traffic.json")
  try:
    d = json.load(open(path))
    if d.get("workload") != workload:
      return None, None
    if d.get("kernel_source_sha") != kernel_source_sha():
      return None, "stale: %s was captured from other kernel sources" % os.path.relpath(path, ROOT)
    return float(d["bytes_per_launch"]), d.get("source")
  except (OSError, ValueError, KeyError):
    pass
  return None, None


def init_process_group_quiet(rank, world, device):
  """torch.distributed over NCCL with stdout pointed at stderr while the communicator comes
  up: NCCL prints its version banner on stdout, and rank 0 must print exactly one JSON line."""
  import torch
  import torch.distributed as dist
  sys.stdout.flush()
  saved = os.dup(1)
  os.dup2(2, 1)
  try:
    dist.init_process_group(backend="nccl", rank=rank, world_size=world, device_id=device)
    probe = torch.zeros(1, device=device)
    dist.all_reduce(probe)
    torch.cuda.synchronize()
  finally:
    sys.stdout.flush()
    os.dup2(saved, 1)
    os.close(saved)


class ClockSampler(object):
  """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""
  QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
           "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
           "clocks_event_reasons.sw_power_cap")

  def __init__(self, index=0):
    self.index = index
    self.rows = []
    self.proc = None

  def start(self):
    try:
      self.proc = subprocess.Popen(
          ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
           "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
          stderr=subprocess.DEVNULL, text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
    except OSError:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append([c.strip() for c in line.split(",")])

  def stop(self):
    if not self.proc:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    time.sleep(0.15)
    self.proc.terminate()
    try:
      self.proc.wait(timeout=5)
    except subprocess.TimeoutExpired:
      self.proc.kill()
    sm, mx, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for r in self.rows:
      try:
        sm.append(float(r[0]))
        mx.append(float(r[1]))
      except (ValueError, IndexError):
        continue
      for name, flag in zip(names, r[3:7]):
        if flag.lower().startswith("active"):
          reasons.add(name)
    return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
            "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline_port(A, B, spec, sweeps, threads=0, keep=None):
  """The oracle port (oracle/algdist_ref.c: the reference's per-row f64 arithmetic, rows spread
  over all host cores the way the reference spreads them over a process pool) on `sweeps`
  sweeps of the full workload.  Returns (nnz*R*iters/s, seconds, threads); the relaxed vectors
  are appended to `keep` when given (bench.py checks the GPU result against them)."""
  from hypergraphembedding_b200 import synthetic
  from oracle import cport
  xn0, xe0 = synthetic.legacy_initial_vectors(A.shape[0], A.shape[1], spec["R"], seed=0)
  xn0, xe0 = xn0.astype(np.float64), xe0.astype(np.float64)
  cport.load()
  t = time.time()
  result = cport.algdist(A, B, xn0, xe0, sweeps, threads=threads)
  dt = time.time() - t
  if keep is not None:
    keep.extend(result)
  return A.nnz * spec["R"] * sweeps / dt, dt, (threads or cport.max_threads())


def cpu_baseline_scipy(A, B, spec, sweeps=2):
  """the vectorised scipy f64 restatement of the same arithmetic
  (oracle/port.py, validated against the reference to 3e-8) on one core -- scipy's sparse
  products are single-threaded -- on a bounded sample of the workload's sweeps."""
  from hypergraphembedding_b200 import synthetic
  from oracle import port
  xn0, xe0 = synthetic.legacy_initial_vectors(A.shape[0], A.shape[1], spec["R"], seed=0)
  t = time.time()
  port.algdist_vectorised(A, B, xn0, xe0, sweeps)
  dt = time.time() - t
  return {"value": A.nnz * spec["R"] * sweeps / dt, "unit": "nnz*R*iters/s", "cores": 1, "kind": "port",
          "sample": "%d of %d sweeps of the full workload, scipy CSR products in f64 (%.1f s)"
                    % (sweeps, spec["sweeps"], dt)}


def workload_config(spec, world, N, E, R, sweeps):
  """`config` of a bench line -- the workload and the cache policy of the measurement, nothing
  run-specific -- so that the two arms (ours, --impl reference) print the same object at every N.
  At N > 1 the workload is N config-2-shaped blocks of node rows (seeds seed .. seed + N - 1)
  over the same edges; per-run figures (incidence counts, exchange strategy) are top-level keys
  of our line (`detail`)."""
  if world == 1:
    return {"workload": spec["name"], "nodes": int(N), "edges": int(E), "R": int(R), "sweeps": int(sweeps),
            "seed": spec["seed"],
            "l2": "no flush: the per-step working set (vectors %d MB + incidence ~%d MB) exceeds the 126 MB L2"
                  % ((N + E) * R * 4 >> 20, 2 * spec["num_incidences"] * 4 >> 20)}
  return {"workload": "%d x [%s] node blocks over the same %d edges" % (world, spec["name"], E),
          "nodes": int(N) * world, "edges": int(E), "R": int(R), "sweeps": int(sweeps), "seed": spec["seed"],
          "l2": "no flush: per-GPU working set exceeds the 126 MB L2"}


def run_reference(args, spec):
  """--impl reference: the reference's CPU implementation of the path on this box's host cores.
  The reference is pure Python (there is nothing of it to compile into oracle/_ref) and its
  tree does not travel to the GPU box, so this is the oracle port of it: the same per-row f64
  arithmetic in C, one row block per host thread, every step the full workload."""
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  A, B = build_workload(spec)
  sweeps = spec["sweeps"]
  for _ in range(max(0, min(args.warmup, 1))):
    cpu_baseline_port(A, B, spec, 1)
  vals, secs, threads = [], [], 1
  for _ in range(max(1, args.steps)):
    v, dt, threads = cpu_baseline_port(A, B, spec, sweeps)
    vals.append(v)
    secs.append(dt)
  value = float(A.nnz * spec["R"] * sweeps * len(secs) / sum(secs))
  world = max(1, int(args.gpus))
  sample = "the full workload (%d sweeps) per step" % sweeps
  if world > 1:
    sample = ("one of the %d node blocks (the block of seed %d: %d incidences, %d sweeps) per step -- the CPU "
              "rate per incidence does not depend on how many blocks are stacked" % (world, spec["seed"], A.nnz, sweeps))
  out = {
      "impl": "reference", "metric": "alg-dist incidence nnz*R*iters/sec", "value": value,
      "unit": "nnz*R*iters/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
      "ms_per_step": 1e3 * float(np.mean(secs)), "higher_is_better": True, "scaling": "weak",
      "vs_baseline": None, "dtype": "f64", "data": "synthetic",
      "config": workload_config(spec, world, A.shape[0], A.shape[1], spec["R"], sweeps),
      "detail": {"nnz_of_the_sample": int(A.nnz)},
      "cpu_baseline": {"value": value, "unit": "nnz*R*iters/s", "cores": threads, "kind": "port",
                       "sample": sample, "host_cores": os.cpu_count()},
      "e2e": {"value": value, "unit": "nnz*R*iters/s", "h2d_bytes_per_step": 0,
              "d2h_bytes_per_step": 0},
  }
  print(json.dumps(out), flush=True)


def hobe_extra(ctx):
  """Second half of BASELINE.json's metric: HOBE weighted samples/s on configs[0] (the youtube
  fixture, 5 neighbours, 200 samples per row; reference: 1.47e4 samples/s on 8 processes)."""
  import hypergraphembedding_b200 as H
  path = os.path.join(ROOT, "tests", "golden", "algdist_youtube.npz")
  g = np.load(path)
  node_ids, edge_ids = g["node_ids"], g["edge_ids"]
  hg = H.Hypergraph()
  emb = H.HypergraphEmbedding()
  for n, e in zip(np.searchsorted(node_ids, g["pairs"][:, 0]).tolist(),
                  np.searchsorted(edge_ids, g["pairs"][:, 1]).tolist()):
    hg.node[n].edges.append(e)
    hg.edge[e].nodes.append(n)
  for i, v in enumerate(g["xn"]):
    emb.node[i].values.extend(v.tolist())
  for i, v in enumerate(g["xe"]):
    emb.edge[i].values.extend(v.tolist())
  np.random.seed(0)
  H.AlgebraicDistanceSamples(hg, emb, 5, 20)        # warm-up
  times = []
  for _ in range(3):
    np.random.seed(0)
    t = time.perf_counter()
    out = H.AlgebraicDistanceSamples(hg, emb, 5, 200)
    times.append(time.perf_counter() - t)
  np.random.seed(0)
  t = time.perf_counter()
  fobe = H.BooleanSamples(hg, 5, 200)
  fobe_s = time.perf_counter() - t
  return {"workload": "snap_youtube_tiny (3862 nodes / 50 edges / 4548 incidences), R=10, "
                      "num_neighbors=5, num_samples=200",
          "records": len(out), "seconds": min(times), "samples_per_s": len(out) / min(times),
          "unit": "weighted samples/s (proto in, columnar records out)",
          "fobe_records": len(fobe), "fobe_samples_per_s": len(fobe) / fobe_s}


def hobe_scale_extra(ctx):
  """HOBE weighted samples/s where the number means something: the 100 000-node member of the
  config-4 family (485 K incidences... see `workload`), num_neighbors 5, num_samples 20, R = 10.
  Split into the sequential replay of numpy's RNG stream on the host and the probability kernels
  on the device; pair sets, neighbour arrays and RNG state are checked against the committed
  digest of the scipy / numpy oracle (tests/golden/hobe_scale.npz), probabilities on a stride."""
  import hashlib
  import hypergraphembedding_b200 as H
  from hypergraphembedding_b200 import synthetic
  path = os.path.join(ROOT, "tests", "golden", "hobe_scale.npz")
  if not os.path.exists(path):
    return {"error": "tests/golden/hobe_scale.npz is missing"}
  g = np.load(path)
  A = synthetic.zipf_hypergraph(int(g["nodes"]), int(g["edges"]), seed=int(g["graph_seed"]))
  xn, xe = synthetic.legacy_initial_vectors(A.shape[0], A.shape[1], int(g["R"]), seed=int(g["vec_seed"]))
  k, num_samples = int(g["k"]), int(g["num_samples"])

  def sha(cols):
    h = hashlib.sha256()
    for c in cols:
      h.update(np.ascontiguousarray(c, dtype=np.int64).tobytes())
    return h.hexdigest()

  best, split, out = None, None, None
  for _ in range(3):
    np.random.seed(int(g["seed"]))
    timings = {}
    t = time.perf_counter()
    out = H.AlgebraicDistanceSamplesCsr(A, xn, xe, k, num_samples, timings=timings)
    dt = time.perf_counter() - t
    if best is None or dt < best:
      best, split = dt, timings
  prob = np.where(~np.isnan(out.nn_prob), out.nn_prob,
                  np.where(~np.isnan(out.ee_prob), out.ee_prob, out.ne_prob))
  want = g["prob_strided"]
  got = prob[::int(g["stride"])]
  return {"workload": "synthetic Zipf hypergraph %d nodes / %d edges / %d incidences (config-4 family), R=%d, "
                      "num_neighbors=%d, num_samples=%d" % (A.shape[0], A.shape[1], A.nnz, int(g["R"]), k,
                                                            num_samples),
          "records": len(out), "seconds": best, "samples_per_s": len(out) / best,
          "host_rng_replay_s": split["draw"], "gpu_probabilities_s": split["probabilities"],
          "incidence_weights_s": split["weights"],
          "pair_sets_bit_exact": bool(len(out) == int(g["count"]) and
                                      sha([out.left_node, out.left_edge, out.right_node, out.right_edge]) ==
                                      str(g["index_sha"])),
          "neighbours_bit_exact": bool(sha([out.neigh_node, out.neigh_edge]) == str(g["neigh_sha"])),
          "max_prob_err_over_bound": float((np.abs(got - want) / (1e-5 * np.abs(want) + 1e-6)).max()),
          "unit": "weighted samples/s (CSR + dense vectors in, columnar records out)"}


def fobe_scale_extra(ctx):
  """FOBE (BooleanSamples, hg2v_sample.py:125-242) on the config-4 family inside the default line:
  the 100 000-node member against the committed digest of the scipy / numpy oracle
  (tests/golden/boolean_c4.npz: 11.0 M records, index / neighbour columns and RNG state by
  SHA-256), then the 1 000 000-node member for throughput (the 10 M-node configs[3] itself is
  `--workload c4`: ~1e9 records).  Host work: FOBE probabilities are all 1."""
  import hashlib
  from hypergraphembedding_b200 import synthetic
  from hypergraphembedding_b200.hg2v_sample import BooleanSamplesCsr
  g = np.load(os.path.join(ROOT, "tests", "golden", "boolean_c4.npz"))

  def sha(arrays, keys):
    h = hashlib.sha256()
    for k in keys:
      h.update(np.ascontiguousarray(arrays[k], dtype=np.int64).tobytes())
    return h.hexdigest()

  A = synthetic.zipf_hypergraph(100000, 50000, seed=int(g["graph_seed"]))
  np.random.seed(int(g["seed"]))
  t = time.perf_counter()
  out = BooleanSamplesCsr(A, int(g["k"]), int(g["num_samples"]))
  small_s = time.perf_counter() - t
  arrays = out.arrays()
  state = np.random.get_state()
  exact = bool(len(out) == int(g["standalone_count"]) and
               sha(arrays, ("left_node", "left_edge", "right_node", "right_edge")) == str(g["standalone_index_sha"]) and
               sha(arrays, ("neigh_node", "neigh_edge")) == str(g["standalone_neigh_sha"]) and
               int(state[2]) == int(g["standalone_rng_pos"]) and
               hashlib.sha256(state[1].tobytes()).hexdigest() == str(g["standalone_rng_key_sha"]))
  del out, arrays
  A = synthetic.zipf_hypergraph(1000000, 500000, seed=int(g["graph_seed"]))
  np.random.seed(int(g["seed"]))
  t = time.perf_counter()
  out = BooleanSamplesCsr(A, int(g["k"]), int(g["num_samples"]))
  big_s = time.perf_counter() - t
  records = len(out)
  del out
  return {"workload": "config-4 family (Zipf degrees), num_neighbors %d, num_samples %d" % (int(g["k"]),
                                                                                           int(g["num_samples"])),
          "nodes_100k": {"records": int(g["standalone_count"]), "seconds": small_s,
                         "samples_per_s": int(g["standalone_count"]) / small_s,
                         "bit_exact_vs_oracle_digest": exact},
          "nodes_1m": {"incidences": int(A.nnz), "records": records, "seconds": big_s,
                       "samples_per_s": records / big_s},
          "unit": "samples/s (CSR in, columnar records out; host RNG replay)"}


def hg2v_train_extra(ctx, dimension=32, epochs=3):
  """The consumer of the HOBE sample columns: UnweightedFloatModel (hg2v_model.py:129-203) trained
  on the 755 267 configs[0] records with the reference's fit settings (batch 256, Adagrad;
  embedding.py:387-414); dimension 32 (runner.py asks for a dimension below the 50 edges)."""
  import hypergraphembedding_b200 as H
  path = os.path.join(ROOT, "tests", "golden", "algdist_youtube.npz")
  g = np.load(path)
  node_ids, edge_ids = g["node_ids"], g["edge_ids"]
  hg = H.Hypergraph()
  emb = H.HypergraphEmbedding()
  for n, e in zip(np.searchsorted(node_ids, g["pairs"][:, 0]).tolist(),
                  np.searchsorted(edge_ids, g["pairs"][:, 1]).tolist()):
    hg.node[n].edges.append(e)
    hg.edge[e].nodes.append(n)
  for i, v in enumerate(g["xn"]):
    emb.node[i].values.extend(v.tolist())
  for i, v in enumerate(g["xe"]):
    emb.edge[i].values.extend(v.tolist())
  np.random.seed(0)
  samples = H.AlgebraicDistanceSamples(hg, emb, 5, 200)
  feats, targets = H.SamplesToModelInput(samples, 5, weighted=False)
  model = H.UnweightedFloatModel(hg, dimension, 5)
  model.set_samples(feats, targets)
  m = len(samples)
  model.fit_epoch(np.random.permutation(m), 256)          # warm-up epoch
  secs, losses = [], []
  for _ in range(epochs):
    order = np.random.permutation(m)
    t = time.perf_counter()
    losses.append(model.fit_epoch(order, 256))
    secs.append(time.perf_counter() - t)
  # the same records in batches of 4 096: one cluster per 256 samples, global barrier between phases
  big = {}
  for cap, key in ((0, "many_clusters"), (1, "one_cluster")):
    ctx.set_trainer_clusters(cap)
    model.fit_epoch(np.random.permutation(m), 4096)
    t = time.perf_counter()
    model.fit_epoch(np.random.permutation(m), 4096)
    dt = time.perf_counter() - t
    big[key] = {"clusters": model.last_clusters, "epoch_s": dt, "samples_per_s": m / dt,
                "us_per_batch": 1e6 * dt / ((m + 4095) // 4096)}
  ctx.reset_tuning()
  model.close()
  batches = (m + 255) // 256
  return {"workload": "UnweightedFloatModel on the %d HOBE records of snap_youtube_tiny, dimension %d, "
                      "num_neighbors 5, batch 256, Adagrad" % (m, dimension),
          "samples_per_s": m / min(secs), "epoch_s": min(secs), "us_per_batch": 1e6 * min(secs) / batches,
          "batches_per_epoch": batches, "epoch_losses": losses, "batch_4096": big,
          "kernel": "k_hg2v_epoch: one launch per epoch, one 8-CTA cluster per 256 samples of a batch, "
                    "2 barriers per batch (hardware cluster barrier; + one global arrival per cluster "
                    "when a batch spans several)"}


def c1_end_to_end_extra(ctx, dimension=32):
  """BASELINE.json configs[0] end to end: EmbedHg2vAlgDist (embedding.py:387-414) on the
  snap_youtube_tiny hypergraph with the reference's defaults -- compress, algebraic-distance
  embedding (R = 10, 20 sweeps), 755 267 HOBE samples, model input, up to 10 epochs of
  UnweightedFloatModel with EarlyStopping, embedding proto out.  (The reference needs 51 s for
  the sampling stage alone on 8 processes)"""
  import hypergraphembedding_b200 as H
  g = np.load(os.path.join(ROOT, "tests", "golden", "algdist_youtube.npz"))
  hg = H.Hypergraph()
  for n, e in g["pairs"].tolist():
    hg.node[n].edges.append(e)
    hg.edge[e].nodes.append(n)
  np.random.seed(0)
  H.EmbedHg2vAlgDist(hg, dimension, num_samples=5, epochs=1, disable_pbar=True)     # warm-up
  secs = []
  for _ in range(2):
    np.random.seed(0)
    t = time.perf_counter()
    emb = H.EmbedHg2vAlgDist(hg, dimension, disable_pbar=True)
    secs.append(time.perf_counter() - t)
  return {"workload": "EmbedHg2vAlgDist on snap_youtube_tiny (3862 nodes / 50 edges / 4548 incidences), "
                      "dimension %d, reference defaults (num_neighbors 5, num_samples 200, batch 256, "
                      "epochs <= 10 with early stopping)" % dimension,
          "seconds": min(secs), "method_name": emb.method_name, "nodes_embedded": len(emb.node),
          "edges_embedded": len(emb.edge)}


def pair_weighting_extra(ctx, num_pairs=100000000):
  """BASELINE.json configs[2]: AMiner-shaped bipartite hypergraph, R=64, distance + HOBE weight
  transform of 100M sampled (node, edge) pairs, everything resident on the device."""
  import torch
  from hypergraphembedding_b200 import _native, synthetic
  from hypergraphembedding_b200 import algebraic_distance as ad
  t = time.time()
  A = synthetic.bipartite_author_paper()
  gen_s = time.time() - t
  N, E = A.shape
  R, sweeps = 64, 20
  xn0, xe0 = synthetic.legacy_initial_vectors(N, E, R, seed=0)
  inc = ad.make_incidence(A, ctx=ctx)
  xn, xe = torch.from_numpy(xn0).cuda(), torch.from_numpy(xe0).cuda()
  init_n, init_e = xn.clone(), xe.clone()
  ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
  _native.algdist_run(ctx, inc, xn, xe, sweeps)
  xn.copy_(init_n)
  xe.copy_(init_e)
  ev[0].record()
  _native.algdist_run(ctx, inc, xn, xe, sweeps)
  ev[1].record()
  coo = A.tocoo()
  rows = torch.from_numpy(coo.row.astype(np.int32)).cuda()
  cols = torch.from_numpy(coo.col.astype(np.int32)).cuda()
  gen = torch.Generator(device="cuda")
  gen.manual_seed(4321)
  pick = torch.randint(0, A.nnz, (num_pairs,), device="cuda", generator=gen)
  ia, ib = rows[pick].contiguous(), cols[pick].contiguous()
  del pick
  d = _native.pair_l2(ctx, xn, xe, ia, ib)           # warm-up
  _native.scale_transform(ctx, d, 0.0)
  torch.cuda.synchronize()
  ev[2].record()
  d = _native.pair_l2(ctx, xn, xe, ia, ib)
  _native.scale_transform(ctx, d, 0.0)
  ev[3].record()
  torch.cuda.synchronize()
  relax_ms, pair_ms = ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3])
  # parity at full size: the L2 distances of every 997th pair against numpy f64 on the same
  # vectors (|d - d_ref| <= 1e-5 d_ref + 1e-6 sqrt(R)), and the transform
  # 1 - (d - min) / (max - min) of ALL pairs recomputed from those distances with the extrema
  # taken over all 100 M of them (hg2v_weighting.py:94-95, 301-333)
  d_raw = _native.pair_l2(ctx, xn, xe, ia, ib)
  w_all = d_raw.clone()
  _native.scale_transform(ctx, w_all, 0.0)
  sel = torch.arange(0, num_pairs, 997, device="cuda")
  hn, he = xn.cpu().numpy().astype(np.float64), xe.cpu().numpy().astype(np.float64)
  sa, sb = ia[sel].cpu().numpy(), ib[sel].cpu().numpy()
  d_ref = np.sqrt(((hn[sa] - he[sb])**2).sum(axis=1))
  d_got = d_raw[sel].cpu().numpy().astype(np.float64)
  dist_ratio = float((np.abs(d_got - d_ref) / (1e-5 * d_ref + 1e-6 * np.sqrt(R))).max())
  lo, hi = float(d_raw.min()), float(d_raw.max())
  w_ref = 1.0 - (d_got - lo) / (hi - lo)
  w_got = w_all[sel].cpu().numpy().astype(np.float64)
  weight_ratio = float((np.abs(w_got - w_ref) / (1e-5 * np.abs(w_ref) + 1e-6)).max())
  parity = {"pairs_checked": int(len(sel)), "max_dist_err_over_bound": dist_ratio,
            "max_weight_err_over_bound": weight_ratio,
            "weights_in_unit_interval": bool(float(w_all.min()) == 0.0 and float(w_all.max()) == 1.0),
            "ok": bool(dist_ratio <= 1.0 and weight_ratio <= 1.0),
            "against": "numpy f64 on the same vectors, every 997th of the %d pairs" % num_pairs}
  del d_raw, w_all
  inc.close()
  bytes_pair = 8 * R + 12
  peak, _ = hbm_peak()
  return {"workload": "AMiner-shaped synthetic bipartite hypergraph %d author nodes / %d paper "
                      "edges / %d incidences, R=%d" % (N, E, A.nnz, R),
          "pairs": num_pairs, "pairs_per_s": num_pairs / (pair_ms * 1e-3), "ms": pair_ms,
          "algorithmic_GBps": num_pairs * bytes_pair / (pair_ms * 1e-3) / 1e9,
          "frac_of_measured_hbm_peak": num_pairs * bytes_pair / (pair_ms * 1e-3) / 1e9 / peak,
          "relaxation_ms_20_sweeps": relax_ms,
          "relaxation_nnz_R_iters_per_s": A.nnz * R * sweeps / (relax_ms * 1e-3),
          "parity": parity, "generate_s": gen_s}


def run_fobe(args, spec):
  """--workload c4: BooleanSamples on the 10M-node hypergraph, streamed in row chunks (the full
  result is ~1e9 records).  First the sub-hypergraph induced by the first 100 000 nodes is
  sampled and compared with the committed digest of the scipy / numpy oracle
  (tests/golden/boolean_c4.npz), then the full size is timed.  FOBE probabilities are all 1, so
  this workload has no device arithmetic: the cost is the sequential replay of numpy's MT19937
  stream (one draw per candidate of every product row) on one host thread, with the candidate
  rows built ahead by worker threads."""
  import hashlib
  from hypergraphembedding_b200 import synthetic
  from hypergraphembedding_b200.hg2v_sample import BooleanSamplesCsr, _Graph
  if int(os.environ.get("RANK", "0")) != 0:
    return
  t = time.time()
  A = synthetic.zipf_hypergraph(spec["num_nodes"], spec["num_edges"], seed=spec["seed"])
  gen_s = time.time() - t
  log("generated %s in %.1f s: %d incidences" % (spec["name"], gen_s, A.nnz))
  k, num_samples = spec["k"], spec["num_samples"]

  def digest(cols, keys):
    h = hashlib.sha256()
    for key in keys:
      h.update(np.ascontiguousarray(getattr(cols, key), dtype=np.int64).tobytes())
    return h.hexdigest()

  parity = None
  golden_path = os.path.join(ROOT, "tests", "golden", "boolean_c4.npz")
  if spec["num_nodes"] == 10000000 and os.path.exists(golden_path):
    g = np.load(golden_path)
    if "induced_count" in g.files:
      sub = synthetic.induced_on_first_nodes(A, 100000)
      np.random.seed(int(g["seed"]))
      out = BooleanSamplesCsr(sub, k, num_samples)
      parity = {"case": "sub-hypergraph induced by the first 100000 nodes vs scipy/numpy oracle digest",
                "records": len(out),
                "bit_exact": bool(len(out) == int(g["induced_count"]) and
                                  digest(out, ("left_node", "left_edge", "right_node", "right_edge")) ==
                                  str(g["induced_index_sha"]) and
                                  digest(out, ("neigh_node", "neigh_edge")) == str(g["induced_neigh_sha"]))}
      del out
  t = time.time()
  graph = _Graph.from_csr(A)
  prep_s = time.time() - t
  secs, records, check = [], 0, 0
  for step in range(max(1, args.steps)):
    np.random.seed(0)
    t = time.time()
    records, check = 0, 0
    for part in BooleanSamplesCsr(graph, k, num_samples, chunk_rows=262144, stream=True):
      records += len(part)
      check ^= int(np.bitwise_xor.reduce(part.left_node.astype(np.int64) * 1000003 +
                                         part.right_node.astype(np.int64) * 10007 +
                                         part.left_edge.astype(np.int64) * 101 +
                                         part.right_edge.astype(np.int64))) if len(part) else 0
    secs.append(time.time() - t)
    log("step %d: %d records in %.1f s" % (step, records, secs[-1]))
  best = min(secs)
  out = {"metric": "FOBE samples/sec", "value": records / best, "unit": "samples/s", "n_gpus": args.gpus,
         "steps": len(secs), "warmup": 0, "ms_per_step": best * 1e3, "higher_is_better": True,
         "scaling": "replicas only", "vs_baseline": None, "dtype": "int32/uint32", "data": "synthetic",
         "config": {"workload": spec["name"], "nodes": int(A.shape[0]), "edges": int(A.shape[1]),
                    "nnz": int(A.nnz), "seed": spec["seed"], "records": records,
                    "xor_checksum": check, "generate_s": gen_s, "transpose_s": prep_s,
                    "host_threads": os.cpu_count(), "parity": parity},
         "roofline": None, "cpu_baseline": None, "e2e": None, "gpu_launches": 0,
         "note": "host-only workload: every FOBE probability is 1, the work is numpy's sequential "
                 "MT19937 stream replayed bit-exactly (DESIGN.md section 4)"}
  print(json.dumps(out), flush=True)


def run_ours(args, spec):
  import torch
  import torch.distributed as dist
  from hypergraphembedding_b200 import _native, synthetic
  from hypergraphembedding_b200 import algebraic_distance as ad

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  if spec["kind"] == "community":
    return run_community(args, spec, world, rank, local_rank)
  if world > 1:
    return run_sharded(args, spec, world, rank, local_rank)
  torch.cuda.set_device(local_rank)
  ctx = _native.default_context(local_rank)

  A, B = build_workload(spec)
  N, E = A.shape
  nnz, R, sweeps = int(A.nnz), spec["R"], spec["sweeps"]
  xn0, xe0 = synthetic.legacy_initial_vectors(N, E, R, seed=0)

  # ---- device-resident arm ---------------------------------------------------------------
  a_ptr = torch.from_numpy(A.indptr.astype(np.int64)).cuda()
  a_idx = torch.from_numpy(A.indices.astype(np.int32)).cuda()
  b_ptr = torch.from_numpy(B.indptr.astype(np.int64)).cuda()
  b_idx = torch.from_numpy(B.indices.astype(np.int32)).cuda()
  inc = _native.Incidence(ctx, N, E, a_ptr, a_idx, b_ptr, b_idx)
  xn_init = torch.from_numpy(xn0).cuda()
  xe_init = torch.from_numpy(xe0).cuda()
  xn = torch.empty_like(xn_init)
  xe = torch.empty_like(xe_init)

  def step_device():
    xn.copy_(xn_init)
    xe.copy_(xe_init)
    _native.algdist_run(ctx, inc, xn, xe, sweeps)

  # clocks / throttle reasons are sampled from the warm-up to the end of the per-launch timing
  # pass below, so that the short timed region (~0.1 s) is inside a longer window under load
  sampler = ClockSampler(local_rank)
  sampler.start()
  for _ in range(args.warmup):
    step_device()
  torch.cuda.synchronize()
  launches0 = ctx.launch_count
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  torch.cuda.synchronize()
  ev0.record()
  for _ in range(args.steps):
    step_device()
  ev1.record()
  torch.cuda.synchronize()
  total_ms = ev0.elapsed_time(ev1)
  launches = ctx.launch_count - launches0
  ms_per_step = total_ms / args.steps
  value = nnz * R * sweeps / (ms_per_step * 1e-3)
  gpu_xn, gpu_xe = xn.cpu().numpy(), xe.cpu().numpy()      # result of the last timed step

  # ---- per-launch duration of the half-sweep kernel (stepwise API, same stream) -------------
  st = _native.AlgDistState(ctx, inc, R, sweeps)
  half_ms = []
  for rep in range(max(1, min(args.steps, 3))):
    xn.copy_(xn_init)
    xe.copy_(xe_init)
    st.load(xn, xe)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(2 * sweeps + 1)]
    evs[0].record()
    for t in range(sweeps):
      st.node_half(t)
      evs[2 * t + 1].record()
      st.edge_half(t)
      evs[2 * t + 2].record()
    torch.cuda.synchronize()
    half_ms.extend(evs[i].elapsed_time(evs[i + 1]) for i in range(2 * sweeps))
  st.store(sweeps, xn, xe)
  st.close()
  clocks = sampler.stop()
  half_ms = np.asarray(half_ms)
  node_ms = float(half_ms[0::2].mean())
  edge_ms = float(half_ms[1::2].mean())
  mean_half_ms = float(half_ms.mean())
  bytes_per_launch = algorithmic_bytes_per_sweep(N, E, nnz, R) / 2.0
  achieved = bytes_per_launch / (mean_half_ms * 1e-3) / 1e9
  peak, peak_src = hbm_peak()
  traffic, traffic_src = ncu_traffic(args.workload)

  # ---- end-to-end arm: host buffers through the C ABI --------------------------------------
  pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
  h_aptr, h_aidx = pin(A.indptr.astype(np.int64)), pin(A.indices.astype(np.int32))
  h_xn0, h_xe0 = pin(xn0), pin(xe0)
  h_xn, h_xe = pin(np.empty_like(xn0)), pin(np.empty_like(xe0))

  def step_host():
    # the vectors are relaxed in place; the next step starts from this step's output, which is
    # again a valid input in [0, 1] (no host-to-host reset copy inside the timed region)
    # one orientation is uploaded; the library transposes it on the device.  One call of the
    # C ABI (hge_algdist_run_csr: what EmbedAlgebraicDistance makes per hypergraph)
    _native.algdist_run_csr(ctx, N, E, h_aptr.numpy(), h_aidx.numpy(), h_xn.numpy(), h_xe.numpy(), sweeps)

  e2e_steps = max(1, min(args.steps, 5))
  h_xn.copy_(h_xn0)
  h_xe.copy_(h_xe0)
  step_host()
  # parity spot check: the host-buffer arm must reproduce the device arm bit for bit
  same = bool(np.array_equal(xn.cpu().numpy(), h_xn.numpy()))
  for _ in range(min(args.warmup, 2)):
    step_host()
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  for _ in range(e2e_steps):
    step_host()
  torch.cuda.synchronize()
  e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
  h2d = h_aptr.numel() * 8 + h_aidx.numel() * 4 + (h_xn0.numel() + h_xe0.numel()) * 4
  d2h = (h_xn.numel() + h_xe.numel()) * 4


  # the C port runs the full workload from the same seed-0 vectors: it is the CPU baseline AND the
  # checker of the device arm's result (full-size parity on the headline configuration)
  cpu_sweeps = sweeps
  ref = []
  cpu_value, cpu_dt, cpu_threads = cpu_baseline_port(A, B, spec, cpu_sweeps, keep=ref)
  parity = parity_against(A, gpu_xn, gpu_xe, ref[0], ref[1], R)
  parity["against"] = "oracle/algdist_ref.c (f64, the reference's per-row arithmetic), full workload, %d sweeps" % sweeps
  del ref
  try:
    scipy_line = cpu_baseline_scipy(A, B, spec)
  except Exception as exc:
    scipy_line = {"error": "%s: %s" % (type(exc).__name__, exc)}
  extras = {}
  if not args.no_extras:
    inc.close()
    del a_ptr, a_idx, b_ptr, b_idx, xn, xe, xn_init, xe_init
    torch.cuda.empty_cache()
    for key, fn in (("hobe", hobe_extra), ("hobe_scale", hobe_scale_extra), ("fobe_scale", fobe_scale_extra),
                    ("hg2v_train", hg2v_train_extra),
                    ("c1_end_to_end", c1_end_to_end_extra),
                    ("pair_weighting", pair_weighting_extra),
                    ("c5", lambda c: c5_extra(args, 1, 0, local_rank, c))):
      try:
        extras[key] = fn(ctx)
      except Exception as exc:   # the headline line must survive a failing side measurement
        extras[key] = {"error": "%s: %s" % (type(exc).__name__, exc)}

  out = {
      "metric": "alg-dist incidence nnz*R*iters/sec", "value": value, "unit": "nnz*R*iters/s",
      "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
      "data": "synthetic",
      "config": workload_config(spec, 1, N, E, R, sweeps),
      "detail": {"nnz": nnz, "device_vs_host_arm_identical": same},
      "parity": parity,
      "roofline": {"bound": "hbm", "kernel": "k_sweep<8, node | edge>", "achieved": achieved, "peak": peak,
                   "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                   "traffic_source": traffic_src, "peak_source": peak_src,
                   "bytes_per_launch": bytes_per_launch, "ms_per_launch": mean_half_ms,
                   "node_half_ms": node_ms, "edge_half_ms": edge_ms,
                   "frac_of_nominal_8TBs": achieved / 8000.0,
                   # `achieved` counts one gathered row per INCIDENCE, most of which L2
                   # serves; the DRAM side of the same launch:
                   "dram_GBps": traffic / (mean_half_ms * 1e-3) / 1e9 if traffic else None,
                   "dram_frac": traffic / (mean_half_ms * 1e-3) / 1e9 / peak if traffic else None,
                   "compulsory_bytes_per_launch": compulsory_bytes_per_launch(N, E, nnz, R),
                   "compulsory_GBps": compulsory_bytes_per_launch(N, E, nnz, R) / (mean_half_ms * 1e-3) / 1e9,
                   "compulsory_frac": compulsory_bytes_per_launch(N, E, nnz, R) / (mean_half_ms * 1e-3) / 1e9 / peak,
                   "kernel_source_sha": kernel_source_sha(),
                   # what a kernel that does nothing but this access pattern reaches on this part
                   # (committed microbenchmark, NOT measured in this run: `tools/roof/gather_roof c2`,
                   # profiles/r2_gather_roof_c2.log -- config 2's row counts and gathers per row,
                   # uniform-random ids, no padding, no row finish)
                   "gather_roof": ({"node_half_ms": 0.110, "edge_half_ms": 0.132,
                                    "frac_node_half": 0.110 / node_ms, "frac_edge_half": 0.132 / edge_ms,
                                    "source": "profiles/r2_half_sweep.md section 3"}
                                   if args.workload == "c2" else None)},
      "cpu_baseline": {"value": cpu_value, "unit": "nnz*R*iters/s", "cores": cpu_threads,
                       "kind": "port",
                       "sample": "the full workload, %d of %d sweeps, C restatement of the reference's "
                                 "per-row f64 arithmetic on all host threads (%.1f s)"
                                 % (cpu_sweeps, sweeps, cpu_dt),
                       "host_cores": os.cpu_count(), "scipy_restatement_one_core": scipy_line},
      "e2e": {"value": nnz * R * sweeps / (e2e_ms * 1e-3), "unit": "nnz*R*iters/s",
              "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
      "gpu_launches": int(launches),
      "clocks": clocks,
  }
  out.update(extras)
  print(json.dumps(out), flush=True)
  if not parity["ok"]:
    log("PARITY FAILED: %s" % json.dumps(parity))
    sys.exit(1)


def community_shard_device(spec, rank, world, device):
  """One node block of the config-5 family, generated on the GPU (the host generators would
  need minutes and > 100 GB for 1.8B incidences).  Global shape: edge sizes ~ power law
  (exponent 1.5) on [3, 1e6] scaled to the incidence budget, members uniform; every node
  additionally joins one uniformly random edge, so no node is isolated.  Rank r holds nodes
  [r N / P, (r+1) N / P) and of every edge the share  size // P (+1 for the first size % P
  ranks)  of its members, drawn uniformly from the block.  Returns device CSR tensors of the
  block in both orientations (sorted, unique column ids)."""
  import torch
  N, E, nnz = spec["num_nodes"], spec["num_edges"], spec["num_incidences"]
  n0, n1 = N * rank // world, N * (rank + 1) // world
  n_loc = n1 - n0
  rng = np.random.Generator(np.random.PCG64(spec["seed"]))          # same sizes on every rank
  u = rng.random(E)
  a, lo, hi = 1.0 - 1.5, 3.0, 1.0e6
  raw = ((hi**a - lo**a) * u + lo**a)**(1.0 / a)
  sizes = np.clip(raw * ((nnz - N) / raw.sum()), lo, min(hi, N // 2)).astype(np.int64)
  local = sizes // world + (rank < (sizes % world)).astype(np.int64)
  gen = torch.Generator(device=device)
  gen.manual_seed(spec["seed"] * 1000 + rank)
  sizes_t = torch.from_numpy(local).to(device)
  edge_of = torch.repeat_interleave(torch.arange(E, device=device, dtype=torch.int64), sizes_t)
  node_of = torch.randint(0, n_loc, (edge_of.numel(),), device=device, generator=gen)
  keys = edge_of * n_loc + node_of
  del edge_of, node_of
  extra_e = torch.randint(0, E, (n_loc,), device=device, generator=gen)
  keys = torch.cat([keys, extra_e * n_loc + torch.arange(n_loc, device=device, dtype=torch.int64)])
  del extra_e
  keys = torch.unique(keys)                                          # sorted, duplicates removed
  e_idx = torch.div(keys, n_loc, rounding_mode="floor")
  n_idx = keys - e_idx * n_loc
  del keys
  b_idx = n_idx.to(torch.int32)
  b_ptr = torch.zeros(E + 1, dtype=torch.int64, device=device)
  b_ptr[1:] = torch.cumsum(torch.bincount(e_idx, minlength=E), 0)
  keys2 = n_idx * E + e_idx
  del n_idx, e_idx
  keys2, _ = torch.sort(keys2)
  n_row = torch.div(keys2, E, rounding_mode="floor")
  a_idx = (keys2 - n_row * E).to(torch.int32)
  del keys2
  a_ptr = torch.zeros(n_loc + 1, dtype=torch.int64, device=device)
  a_ptr[1:] = torch.cumsum(torch.bincount(n_row, minlength=n_loc), 0)
  del n_row
  torch.cuda.empty_cache()
  return dict(n_loc=n_loc, E=E, a_ptr=a_ptr, a_idx=a_idx, b_ptr=b_ptr, b_idx=b_idx,
              nnz=int(a_idx.numel()))


def community_measure(spec, world, rank, local_rank, ctx, steps, warmup, comm="auto", slices=1,
                      oracle_check=False, force_tiles=False):
  """One strong-scaling measurement of a config-5-family hypergraph over `world` ranks (1: the
  plain single-GPU path).  The process group must be up when world > 1.  Returns a dict on every
  rank.  oracle_check: the blocks and the result are gathered on rank 0 and compared with the C
  port of the reference arithmetic (small members of the family only)."""
  import torch
  import torch.distributed as dist
  from hypergraphembedding_b200 import _native
  from hypergraphembedding_b200 import distributed as hd

  device = torch.device("cuda", local_rank)
  t = time.time()
  g = community_shard_device(spec, rank, world, device)
  torch.cuda.synchronize()
  gen_s = time.time() - t
  n_loc, E, R, sweeps = g["n_loc"], g["E"], spec["R"], spec["sweeps"]
  nnz_t = torch.tensor([g["nnz"]], dtype=torch.int64, device=device)
  if world > 1:
    dist.all_reduce(nnz_t)
  nnz_global = int(nnz_t.item())
  gen = torch.Generator(device=device)
  gen.manual_seed(7 + rank)
  xn_init = torch.rand((n_loc, R), device=device, generator=gen)
  gen.manual_seed(10**6)
  xe_init = torch.rand((E, R), device=device, generator=gen)       # identical on every rank
  xn, xe = torch.empty_like(xn_init), torch.empty_like(xe_init)
  if force_tiles:
    ctx.set_tile_mb(8, 0)     # node-range tiles of 8 MB whatever the sizes (the config-5 code path)

  if world == 1:
    inc = _native.Incidence(ctx, n_loc, E, g["a_ptr"], g["a_idx"], g["b_ptr"], g["b_idx"])
    relax, exchange = None, "single GPU"

    def step():
      xn.copy_(xn_init)
      xe.copy_(xe_init)
      _native.algdist_run(ctx, inc, xn, xe, sweeps)
  else:
    inc = None
    relax = hd.ShardedRelaxation(None, R, sweeps, comm=comm, num_slices=slices, ctx=ctx,
                                 shape=(n_loc, E),
                                 csr_device=(g["a_ptr"], g["a_idx"], g["b_ptr"], g["b_idx"]))
    exchange = "p2p" if relax.use_p2p else "nccl"

    def step():
      xn.copy_(xn_init)
      xe.copy_(xe_init)
      relax.run(xn, xe)

  for _ in range(warmup):
    step()
  launches0 = ctx.launch_count
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  if world > 1:
    dist.barrier()
  torch.cuda.synchronize()
  ev0.record()
  for _ in range(steps):
    step()
  ev1.record()
  torch.cuda.synchronize()
  ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=device)
  if world > 1:
    dist.barrier()
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
  ms_per_step = float(ms.item()) / steps
  launches = ctx.launch_count - launches0
  finite = bool(torch.isfinite(xn).all().item() and torch.isfinite(xe).all().item())
  span_ok = bool((xe.min() >= -1e-6).item() and (xe.max() <= 1 + 1e-6).item())
  out = {"ms_per_step": ms_per_step, "value": nnz_global * R * sweeps / (ms_per_step * 1e-3),
         "nnz": nnz_global, "nnz_per_gpu": g["nnz"], "n_loc": n_loc, "E": E, "exchange": exchange,
         "generate_s": gen_s, "launches": int(launches) * world, "result_finite": finite,
         "result_in_unit_cube": span_ok,
         "tiled": bool(force_tiles or n_loc * R * 4 > (512 << 20)), "parity": None}

  if oracle_check:
    block = [t.cpu().numpy() for t in (g["a_ptr"], g["a_idx"], xn_init, xn)]
    blocks = [None] * world
    if world > 1:
      dist.gather_object(block, blocks if rank == 0 else None, dst=0)
    else:
      blocks = [block]
    if rank == 0:
      import scipy.sparse as sps
      from oracle import cport
      A = sps.vstack([sps.csr_matrix((np.ones(len(b[1]), dtype=bool), b[1], b[0]),
                                     shape=(len(b[0]) - 1, E)) for b in blocks]).tocsr()
      A.sort_indices()
      B = A.T.tocsr()
      B.sort_indices()
      x0 = np.concatenate([b[2] for b in blocks]).astype(np.float64)
      ref_xn, ref_xe = cport.algdist(A, B, x0, xe_init.cpu().numpy().astype(np.float64), sweeps)
      out["parity"] = parity_against(A, np.concatenate([b[3] for b in blocks]), xe.cpu().numpy(),
                                     ref_xn, ref_xe, R)
      out["parity"]["against"] = "oracle/algdist_ref.c (f64) on the gathered blocks, %d sweeps" % sweeps

  if relax is not None:
    relax.close()
  if inc is not None:
    inc.close()
  if force_tiles:
    ctx.set_tile_mb(64, 512)
  del g, xn_init, xe_init, xn, xe
  torch.cuda.empty_cache()
  return out


def c5_extra(args, world, rank, local_rank, ctx):
  """BASELINE.json configs[4] inside the driver-run lines: (1) the 1/100-scale member of the family
  sharded over the ranks and checked against the oracle, untiled and with node-range tiles forced
  (the code path the full size takes); (2) the full 65M-node / 1.8B-incidence hypergraph
  strong-scaled over the ranks; (3) at N > 1, the same hypergraph on rank 0's GPU alone, so the
  speed-up is measured inside one run on one box."""
  import torch
  import torch.distributed as dist
  from hypergraphembedding_b200 import distributed as hd
  mini, full = WORKLOADS["c5mini"], WORKLOADS["c5"]
  out = {"workload": full["name"]}
  checks = []
  for tiles in (False, True):
    m = community_measure(mini, world, rank, local_rank, ctx, 1, 1, comm=args.comm, slices=args.slices,
                          oracle_check=True, force_tiles=tiles)
    if world > 1:
      hd.release_peer_arenas(dist)
    if rank == 0:
      checks.append(dict(m["parity"], tiled=tiles, nnz=m["nnz"], exchange=m["exchange"]))
  out["c5mini_parity"] = checks
  m = community_measure(full, world, rank, local_rank, ctx, max(1, min(args.steps, 3)), 1, comm=args.comm,
                        slices=args.slices)
  if world > 1:
    hd.release_peer_arenas(dist)
  out.update({k: m[k] for k in ("ms_per_step", "value", "nnz", "nnz_per_gpu", "exchange", "tiled",
                                "result_finite", "result_in_unit_cube", "generate_s")})
  out["n_gpus"] = world
  if world > 1:
    one = None
    if rank == 0:
      one = community_measure(full, 1, 0, local_rank, ctx, 2, 1)
    dist.barrier()
    if rank == 0:
      out["n1_ms_per_step"] = one["ms_per_step"]
      out["n1_value"] = one["value"]
      out["speedup_vs_c5_n1"] = one["ms_per_step"] / m["ms_per_step"]
      out["n1_note"] = "the same hypergraph family member on rank 0's GPU alone, measured in this run"
  return out if rank == 0 else None


def run_community(args, spec, world, rank, local_rank):
  """--workload c5 / c5mini: strong scaling of one fixed hypergraph over the ranks (N = 1: the
  plain single-GPU path).  Device-resident arm only."""
  import torch
  import torch.distributed as dist
  from hypergraphembedding_b200 import _native
  from hypergraphembedding_b200 import distributed as hd

  torch.cuda.set_device(local_rank)
  device = torch.device("cuda", local_rank)
  if world > 1:
    init_process_group_quiet(rank, world, device)
  ctx = _native.default_context(local_rank)
  sampler = ClockSampler(local_rank)
  if rank == 0:
    sampler.start()
  m = community_measure(spec, world, rank, local_rank, ctx, args.steps, args.warmup, comm=args.comm,
                        slices=args.slices, oracle_check=spec["num_nodes"] <= 1000000)
  clocks = sampler.stop() if rank == 0 else None
  N, E, R, sweeps = spec["num_nodes"], m["E"], spec["R"], spec["sweeps"]
  bytes_step = algorithmic_bytes_per_sweep(N, E, m["nnz"], R) * sweeps
  achieved = bytes_step / (m["ms_per_step"] * 1e-3) / 1e9 / world      # per GPU, whole step
  peak, peak_src = hbm_peak()
  if rank == 0:
    out = {
        "metric": "alg-dist incidence nnz*R*iters/sec", "value": m["value"], "unit": "nnz*R*iters/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": spec["name"], "nodes": N, "edges": E, "nnz": m["nnz"],
                   "nnz_per_gpu": m["nnz_per_gpu"], "R": R, "sweeps": sweeps, "seed": spec["seed"],
                   "exchange": m["exchange"], "generate_s": m["generate_s"],
                   "result_finite": m["result_finite"], "result_in_unit_cube": m["result_in_unit_cube"],
                   "l2": "no flush: per-GPU working set exceeds the 126 MB L2"},
        "parity": m["parity"],
        "roofline": {"bound": "hbm", "kernel": "whole step (load + %d sweeps + store), per GPU" % sweeps,
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": peak_src,
                     "bytes_per_launch": bytes_step / world, "ms_per_launch": m["ms_per_step"]},
        "cpu_baseline": None, "e2e": None,
        "gpu_launches": m["launches"], "clocks": clocks,
    }
    print(json.dumps(out), flush=True)
  if world > 1:
    hd.release_peer_arenas(dist)
    dist.destroy_process_group()


def host_link_probe(world, rank, mb=256, reps=3):
  """What the host side of the box gives N ranks at once: every rank copies `mb` MB of pinned
  memory to its GPU and back, first rank 0 alone, then all ranks together (CUDA events, max over
  ranks).  The end-to-end arm moves ~0.5 GB per rank and step through these links; when the
  aggregate does not grow with N the host-buffer step cannot scale, whatever the kernels do."""
  import torch
  import torch.distributed as dist
  n = mb << 18
  host = torch.empty(n, dtype=torch.float32).pin_memory()
  dev = torch.empty(n, dtype=torch.float32, device="cuda")

  def timed(h2d):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(reps):
      if h2d:
        dev.copy_(host, non_blocking=True)
      else:
        host.copy_(dev, non_blocking=True)
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / reps

  out = {}
  for name, h2d in (("h2d", True), ("d2h", False)):
    timed(h2d)                                   # warm-up
    dist.barrier()
    alone = torch.tensor([timed(h2d) if rank == 0 else 0.0], dtype=torch.float64, device="cuda")
    dist.barrier()
    dist.all_reduce(alone, op=dist.ReduceOp.MAX)
    dist.barrier()
    together = torch.tensor([timed(h2d)], dtype=torch.float64, device="cuda")
    dist.all_reduce(together, op=dist.ReduceOp.MAX)
    gb = mb * (1 << 20) / 1e9
    out[name + "_GBps_one_rank_alone"] = gb / (float(alone.item()) * 1e-3)
    out[name + "_GBps_per_rank_all_ranks_at_once"] = gb / (float(together.item()) * 1e-3)
    out[name + "_GBps_aggregate"] = world * gb / (float(together.item()) * 1e-3)
  out["note"] = "%d MB pinned copies, CUDA events, slowest rank" % mb
  del host, dev
  return out


def sharded_parity_check(world, rank, ctx, comm, slices, R=32, iters=10):
  """The multi-rank relaxation against the oracle before anything is timed: a 80 000-node /
  3 000-edge / ~0.98M-incidence hypergraph (the case of tests/check_multi_gpu_parity.py), node
  rows sharded over the ranks by incidence count, `iters` sweeps; the blocks are gathered on
  rank 0 and compared with oracle/port.py (scipy f64)."""
  import torch
  import torch.distributed as dist
  import scipy.sparse as sps
  from hypergraphembedding_b200 import distributed as hd
  n, e, nnz = 80000, 3001, 900000
  rng = np.random.default_rng(21)
  rows = np.concatenate([rng.integers(0, n, nnz), np.arange(n), rng.integers(0, n, e)])
  cols = np.concatenate([rng.integers(0, e, nnz), rng.integers(0, e, n), np.arange(e)])
  A = sps.csr_matrix((np.ones(len(rows), dtype=bool), (rows, cols)), shape=(n, e), dtype=bool)
  A.sum_duplicates()
  A.sort_indices()
  rng = np.random.default_rng(123)
  xn0 = rng.random((n, R)).astype(np.float32)
  xe0 = rng.random((e, R)).astype(np.float32)
  A_loc, r0, r1 = hd.local_shard(A, rank, world)
  relax = hd.ShardedRelaxation(A_loc, R, iters, num_slices=slices, comm=comm, ctx=ctx)
  xn = torch.from_numpy(xn0[r0:r1].copy()).cuda()
  xe = torch.from_numpy(xe0.copy()).cuda()
  relax.run(xn, xe)
  torch.cuda.synchronize()
  exchange = "p2p" if relax.use_p2p else "nccl"
  relax.close()
  hd.release_peer_arenas(dist)
  blocks = [None] * world
  dist.gather_object((r0, r1, xn.cpu().numpy()), blocks if rank == 0 else None, dst=0)
  if rank != 0:
    return None
  from oracle import port
  ref_xn, ref_xe = port.algdist_vectorised(A, A.T.tocsr(), xn0, xe0, iters)
  got = np.zeros_like(xn0)
  for b0, b1, blk in blocks:
    got[b0:b1] = blk
  res = parity_against(A, got, xe.cpu().numpy(), ref_xn, ref_xe, R)
  res.update(case="80000 nodes / 3001 edges / %d incidences, R=%d, %d sweeps, %d ranks (%s) vs oracle/port.py"
                  % (A.nnz, R, iters, world, exchange))
  return res


def run_sharded(args, spec, world, rank, local_rank):
  """N > 1: weak scaling.  Every rank owns one config-2-shaped block of node rows (its own
  seed) over the same 500K edges; the global hypergraph is the stack of the blocks."""
  import torch
  import torch.distributed as dist
  from hypergraphembedding_b200 import _native, synthetic
  from hypergraphembedding_b200 import distributed as hd

  torch.cuda.set_device(local_rank)
  init_process_group_quiet(rank, world, torch.device("cuda", local_rank))
  ctx = _native.default_context(local_rank)
  if os.environ.get("HGE_BLOCKS_PER_SM"):
    ctx.set_tuning(0, 0, int(os.environ["HGE_BLOCKS_PER_SM"]))
  parity = sharded_parity_check(world, rank, ctx, args.comm, args.slices)
  if rank == 0:
    log("multi-rank parity: %s" % json.dumps(parity))
  shard_spec = dict(spec, seed=spec["seed"] + rank)
  A, B = build_workload(shard_spec)
  n_loc, E = A.shape
  R, sweeps = spec["R"], spec["sweeps"]
  nnz_local = int(A.nnz)
  nnz_t = torch.tensor([nnz_local], dtype=torch.int64, device="cuda")
  dist.all_reduce(nnz_t)
  nnz_global = int(nnz_t.item())
  xn0, xe0 = synthetic.legacy_initial_vectors(n_loc, E, R, seed=rank)
  xe0 = synthetic.legacy_initial_vectors(1, E, R, seed=10**6)[1]     # identical on every rank

  relax = hd.ShardedRelaxation(A, R, sweeps, num_slices=args.slices, comm=args.comm, ctx=ctx,
                               B_local=B)
  comm = "p2p" if relax.use_p2p else "nccl"
  xn_init, xe_init = torch.from_numpy(xn0).cuda(), torch.from_numpy(xe0).cuda()
  xn, xe = torch.empty_like(xn_init), torch.empty_like(xe_init)

  def step_device():
    xn.copy_(xn_init)
    xe.copy_(xe_init)
    relax.run(xn, xe)

  for _ in range(args.warmup):
    step_device()
  sampler = ClockSampler(local_rank)
  if rank == 0:
    sampler.start()
  launches0 = ctx.launch_count
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  dist.barrier()
  torch.cuda.synchronize()
  ev0.record()
  for _ in range(args.steps):
    step_device()
  ev1.record()
  torch.cuda.synchronize()
  dist.barrier()
  ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
  dist.all_reduce(ms, op=dist.ReduceOp.MAX)
  total_ms = float(ms.item())
  launches = ctx.launch_count - launches0
  clocks = sampler.stop() if rank == 0 else None
  ms_per_step = total_ms / args.steps
  value = nnz_global * R * sweeps / (ms_per_step * 1e-3)

  # per-sweep device time on this rank; with the NCCL exchange also the node half-sweep launch
  # alone (the same kernel and shard shape as at N = 1)
  evs = [torch.cuda.Event(enable_timing=True) for _ in range(2 * sweeps + 1)]
  relax.ops.load(xn_init, xe_init)
  arena = getattr(relax.ops, "arena", None) if relax.use_p2p else None
  if arena is not None:
    arena.set_timing(True)        # phase marks for this pass only, not in the timed region above
  for t in range(sweeps):
    evs[2 * t].record()
    if relax.use_p2p:
      relax.sweep(t)
    else:
      relax.ops.node_half(t)
      evs[2 * t + 1].record()
      relax.sweep_after_node_half(t)
  evs[2 * sweeps].record()
  torch.cuda.synchronize()
  sweep_ms = float(np.mean([evs[2 * t].elapsed_time(evs[2 * t + 2]) for t in range(sweeps)]))
  phases = None
  if arena is not None:
    ms5, n_timed = arena.phase_ms()
    arena.set_timing(False)
    worst = torch.tensor(ms5, dtype=torch.float64, device="cuda")
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    phases = dict(zip(("node_half", "edge_gather_push", "barrier_a", "owner_reduce_all_gather",
                       "barrier_b_minmax"), [float(v) for v in worst.tolist()]))
    phases["note"] = "ms per sweep, max over ranks, CUDA events between the launches of %d sweeps" % n_timed
  peak, peak_src = hbm_peak()
  if relax.use_p2p:
    kernel = "fused sweep: k_sweep<8> node half + sliced edge gather with peer push, per-slice barrier + k_edge_reduce_push on a second stream, min/max barrier (rank 0)"
    bytes_launch = algorithmic_bytes_per_sweep(n_loc, E, nnz_local, R)
    launch_ms = sweep_ms
  else:
    kernel = "k_sweep<8, node> (rank 0)"
    bytes_launch = nnz_local * (4 * R + 4) + 2 * n_loc * 4 * R
    launch_ms = float(np.mean([evs[2 * t].elapsed_time(evs[2 * t + 1]) for t in range(sweeps)]))
  achieved = bytes_launch / (launch_ms * 1e-3) / 1e9
  relax.close()      # its exchange arena goes back to the pool: the host-buffer steps re-use it

  # end to end: host buffers in, host buffers out, incidence upload and set-up collectives inside
  pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
  h_xn0, h_xe0 = pin(xn0), pin(xe0)
  h_xn, h_xe = pin(np.empty_like(xn0)), pin(np.empty_like(xe0))
  h_csr = [pin(A.indptr.astype(np.int64)), pin(A.indices.astype(np.int32))]   # one orientation

  def step_host():
    t_a = time.perf_counter()
    r = hd.ShardedRelaxation(None, R, sweeps, num_slices=args.slices, comm=args.comm, ctx=ctx,
                             shape=(n_loc, E), csr_host=[t.numpy() for t in h_csr])
    t_b = time.perf_counter()
    r.run(h_xn.numpy(), h_xe.numpy())
    t_c = time.perf_counter()
    r.close()
    return (t_b - t_a) * 1e3, (t_c - t_b) * 1e3, (time.perf_counter() - t_c) * 1e3

  e2e_steps = 20 if args.steps >= 10 else max(1, args.steps)   # enough steps that one host hiccup does not decide the mean
  h_xn.copy_(h_xn0)
  h_xe.copy_(h_xe0)
  step_host()
  same = bool(np.array_equal(xn.cpu().numpy(), h_xn.numpy()))
  dist.barrier()
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  step_ms, split = [], []
  for _ in range(e2e_steps):
    t1 = time.perf_counter()
    split.append(step_host())    # returns when this rank's results are in its host buffers
    step_ms.append((time.perf_counter() - t1) * 1e3)
  torch.cuda.synchronize()
  dist.barrier()
  e2e = torch.tensor([(time.perf_counter() - t0) * 1e3 / e2e_steps] + step_ms, dtype=torch.float64,
                     device="cuda")
  dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
  e2e_ms = float(e2e[0].item())
  slowest = int(np.argmax(step_ms))       # this rank's slowest step: set-up / run / tear-down
  slowest_split = dict(zip(("construct_ms", "run_ms", "close_ms"), [round(v, 2) for v in split[slowest]]))
  step_ms = [float(v) for v in e2e[1:].tolist()]
  h2d = A.indptr.size * 8 + A.indices.size * 4 + (xn0.size + xe0.size) * 4
  d2h = (xn0.size + xe0.size) * 4
  host_link = host_link_probe(world, rank)
  hd.release_peer_arenas(dist)
  del xn, xe, xn_init, xe_init
  torch.cuda.empty_cache()
  c5 = None
  if not args.no_extras:
    try:
      c5 = c5_extra(args, world, rank, local_rank, ctx)
    except Exception as exc:       # every rank fails or none does (same shapes, same memory)
      c5 = {"error": "%s: %s" % (type(exc).__name__, exc)}

  if rank == 0:
    out = {
        "metric": "alg-dist incidence nnz*R*iters/sec", "value": value, "unit": "nnz*R*iters/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(spec, world, n_loc, E, R, sweeps),
        "detail": {"nnz": nnz_global, "nnz_per_gpu": nnz_local, "slices": args.slices, "exchange": comm,
                   "partition": "nodes row-partitioned, edge block replicated; per sweep: reduce-scatter of "
                                "the E x R partial sums to the owning GPU, all-gather of the updated edge "
                                "rows, all-reduce(min/max) of 2R bounds -- " +
                                ("stores to peer memory from inside the kernels + flag barriers"
                                 if comm == "p2p" else "NCCL all-reduce between the kernels"),
                   "device_vs_host_arm_identical": same},
        "roofline": {"bound": "hbm", "kernel": kernel,
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": peak_src, "bytes_per_launch": bytes_launch,
                     "ms_per_launch": launch_ms, "ms_per_sweep": sweep_ms, "sweep_phases_ms": phases},
        "cpu_baseline": None,
        "e2e": {"value": nnz_global * R * sweeps / (e2e_ms * 1e-3), "unit": "nnz*R*iters/s",
                "ms_per_step": e2e_ms, "step_ms": step_ms, "median_step_ms": float(np.median(step_ms)),
                "slowest_step_rank0": slowest_split,
                "h2d_bytes_per_step": int(h2d) * world,
                "d2h_bytes_per_step": int(d2h) * world, "host_link_GBps": host_link,
                "copy_floor_ms": 1e3 * (h2d / (host_link["h2d_GBps_per_rank_all_ranks_at_once"] * 1e9) +
                                        d2h / (host_link["d2h_GBps_per_rank_all_ranks_at_once"] * 1e9))},
        "gpu_launches": int(launches) * world,
        "clocks": clocks,
        "parity": parity,
        "c5": c5,
    }
    print(json.dumps(out), flush=True)
  ok = parity is None or parity["ok"]
  hd.release_peer_arenas(dist)
  dist.destroy_process_group()
  if not ok:
    sys.exit(1)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=10)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
  ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
  ap.add_argument("--no-extras", action="store_true",
                  help="skip the HOBE samples/s and 100M-pair weighting side measurements")
  ap.add_argument("--comm", default="auto", choices=["auto", "p2p", "nccl"],
                  help="exchange of the sharded edge half: peer-memory stores inside the kernels, or NCCL")
  ap.add_argument("--slices", type=int, default=1,
                  help="edge slices of the sharded edge half (overlap of all-reduce and gather)")
  args = ap.parse_args()
  spec = WORKLOADS[args.workload]
  if spec["kind"] == "fobe":
    run_fobe(args, spec)
  elif args.impl == "reference":
    run_reference(args, spec)
  else:
    run_ours(args, spec)


if __name__ == "__main__":
  main()
