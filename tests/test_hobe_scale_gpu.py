"""HOBE sampling at 100 000 nodes (tests/golden/hobe_scale.npz, written by
oracle/make_golden_hobe_scale.py from the scipy / numpy oracle): pair sets, neighbour arrays and
the final RNG state bit for bit, probabilities within the 1e-5 bar on every 499th record."""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, load_golden

pytestmark = pytest.mark.gpu


def _sha(cols):
  h = hashlib.sha256()
  for c in cols:
    h.update(np.ascontiguousarray(c, dtype=np.int64).tobytes())
  return h.hexdigest()


@pytest.mark.skipif(not os.path.exists(os.path.join(GOLDEN_DIR, "hobe_scale.npz")),
                    reason="golden digest not generated")
def test_hobe_samples_at_100k_nodes_match_the_oracle_digest():
  from hypergraphembedding_b200 import AlgebraicDistanceSamplesCsr, synthetic
  g = load_golden("hobe_scale")
  A = synthetic.zipf_hypergraph(int(g["nodes"]), int(g["edges"]), seed=int(g["graph_seed"]))
  assert hashlib.sha256(A.indptr.astype(np.int64).tobytes() +
                        A.indices.astype(np.int32).tobytes()).hexdigest() == str(g["csr_sha"])
  xn, xe = synthetic.legacy_initial_vectors(A.shape[0], A.shape[1], int(g["R"]), seed=int(g["vec_seed"]))
  np.random.seed(int(g["seed"]))
  timings = {}
  out = AlgebraicDistanceSamplesCsr(A, xn, xe, int(g["k"]), int(g["num_samples"]), timings=timings)
  state = np.random.get_state()
  assert len(out) == int(g["count"])
  assert _sha([out.left_node, out.left_edge, out.right_node, out.right_edge]) == str(g["index_sha"])
  assert _sha([out.neigh_node, out.neigh_edge]) == str(g["neigh_sha"])
  assert int(state[2]) == int(g["rng_pos"])
  assert hashlib.sha256(state[1].tobytes()).hexdigest() == str(g["rng_key_sha"])
  prob = np.where(~np.isnan(out.nn_prob), out.nn_prob,
                  np.where(~np.isnan(out.ee_prob), out.ee_prob, out.ne_prob))
  got, want = prob[::int(g["stride"])], g["prob_strided"]
  assert np.all(np.abs(got - want) <= 1e-5 * np.abs(want) + 1e-6)
  assert set(timings) == {"draw", "probabilities", "weights"}
