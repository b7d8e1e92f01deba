"""The multi-rank orchestration (node-row sharding, one-time degree collectives, sliced
all-reduce of the edge partial sums, MIN/MAX all-reduce of the rescale bounds) over gloo with
world_size 2 on CPU.  The per-rank kernels are replaced by tests/dist_helpers.NumpyOps; the
collectives, the partition and the protocol are the product's."""
import socket

import numpy as np
import pytest
import scipy.sparse as sps
import torch.multiprocessing as mp

import dist_helpers
from hypergraphembedding_b200 import distributed as hd
from oracle import port


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def test_partition_is_contiguous_and_balanced():
  A = dist_helpers.make_graph(0, 5000, 300, 40000)
  for parts in (1, 2, 3, 8):
    b = hd.partition_rows_by_nnz(A.indptr, parts)
    assert b[0] == 0 and b[-1] == A.shape[0] and np.all(np.diff(b) >= 0)
    nnz = np.diff(A.indptr[b])
    assert nnz.sum() == A.nnz
    assert nnz.max() - nnz.min() <= np.diff(A.indptr).max() + 1
  blocks = [hd.local_shard(A, r, 3) for r in range(3)]
  assert sps.vstack([b[0] for b in blocks]).nnz == A.nnz
  assert [b[1] for b in blocks[1:]] == [b[2] for b in blocks[:-1]]


@pytest.mark.parametrize("world,slices", [(2, 1), (2, 3)])
def test_two_rank_relaxation_matches_single_process_oracle(tmp_path, world, slices):
  graph_args = (7, 900, 60, 5000)
  R, iters = 8, 6
  mp.spawn(dist_helpers.worker,
           args=(world, _free_port(), "gloo", graph_args, R, iters, slices, str(tmp_path), False),
           nprocs=world, join=True)
  A = dist_helpers.make_graph(*graph_args)
  rng = np.random.default_rng(123)
  xn0 = rng.random((A.shape[0], R)).astype(np.float32)
  xe0 = rng.random((A.shape[1], R)).astype(np.float32)
  ref_xn, ref_xe = port.algdist_vectorised(A, A.T.tocsr(), xn0, xe0, iters)
  got_xn = np.zeros_like(ref_xn)
  for r in range(world):
    z = np.load(tmp_path / ("rank%d.npz" % r))
    got_xn[int(z["r0"]):int(z["r1"])] = z["xn"]
    assert np.abs(z["xe"] - ref_xe).max() < 2e-6        # replicated edge block, every rank
  assert np.abs(got_xn - ref_xn).max() < 2e-6


def test_sharded_pair_weights_share_one_min_max(tmp_path):
  """Pairs sharded over 2 ranks (one of them empty-handed): the weights equal the single-process
  transform over the pairs that were weighted, because (min, max) are all-reduced."""
  world, num_pairs, alpha = 2, 4000, 0.25
  mp.spawn(dist_helpers.pair_worker,
           args=(world, _free_port(), "gloo", 5, num_pairs, alpha, str(tmp_path), False),
           nprocs=world, join=True)
  rng = np.random.default_rng(5)
  xa = rng.random((500, 12)).astype(np.float32)
  xb = rng.random((300, 12)).astype(np.float32)
  ia = rng.integers(0, 500, num_pairs).astype(np.int32)
  ib = rng.integers(0, 300, num_pairs).astype(np.int32)
  parts = [np.load(tmp_path / ("pairs%d.npz" % r)) for r in range(world)]
  done = np.concatenate([np.arange(int(z["lo"]), int(z["hi"])) for z in parts])
  got = np.concatenate([z["w"] for z in parts])
  dist64 = np.sqrt(((xa.astype(np.float64)[ia[done]] - xb.astype(np.float64)[ib[done]])**2).sum(1))
  want = alpha + (1 - alpha) * (1 - (dist64 - dist64.min()) / (dist64.max() - dist64.min()))
  assert len(got) == len(done) == num_pairs // 2 and len(parts[1]["w"]) == 0
  assert np.abs(got - want).max() < 1e-5
  assert got.min() >= alpha - 1e-6 and abs(got.max() - 1.0) < 1e-6


def test_rank_local_failure_is_raised_on_every_rank(tmp_path):
  """An isolated node is only visible to the rank that holds it (the edge check is all-reduced,
  the node check is not): the status word agreed after the local set-up makes both ranks raise
  the reference's ZeroDivisionError (algebraic_distance.py:49) -- nobody is left in a collective."""
  world = 2
  mp.spawn(dist_helpers.failure_worker, args=(world, _free_port(), "gloo", (3, 400, 30, 2500), str(tmp_path)),
           nprocs=world, join=True)
  outcomes = [(tmp_path / ("outcome%d.txt" % r)).read_text() for r in range(world)]
  assert all(o.startswith("ZeroDivisionError") for o in outcomes), outcomes
  assert "another rank" in outcomes[0] and "another rank" not in outcomes[1]
