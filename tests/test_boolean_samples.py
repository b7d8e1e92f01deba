"""FOBE sampling (BooleanSamples) is bit-exact against the committed outputs of the unmodified
reference: index columns, neighbour arrays, record count and the final global RNG state.  The
sampler is host code (sequential MT19937 replay), so this runs without a GPU."""
import hashlib

import numpy as np
import pytest

from conftest import hypergraph_from_pairs, load_golden
from hypergraphembedding_b200 import BooleanSamples, SamplesToModelInput
from oracle import port

INDEX_KEYS = ("left_node", "left_edge", "right_node", "right_edge")
NEIGH_KEYS = ("neigh_node", "neigh_edge")


def _sha(arrays, keys):
  h = hashlib.sha256()
  for k in keys:
    h.update(np.ascontiguousarray(arrays[k], dtype=np.int64).tobytes())
  return h.hexdigest()


@pytest.mark.parametrize("name", ["tiny", "rand25", "youtube_s10", "youtube_s200",
                                  "youtube_s200_neg"])
def test_boolean_samples_bit_exact(name):
  g = load_golden("boolean_" + name)
  hg = hypergraph_from_pairs(g["pairs"])
  assert list(hg.node) == g["node_rows"].tolist() and list(hg.edge) == g["edge_rows"].tolist()
  np.random.seed(int(g["seed"]))
  out = BooleanSamples(hg, int(g["k"]), int(g["num_samples"]), neg_samples=int(g["neg"]),
                       disable_pbar=True)
  assert len(out) == int(g["count"])
  arrays = out.arrays()
  assert _sha(arrays, INDEX_KEYS) == str(g["index_sha"])
  assert _sha(arrays, NEIGH_KEYS) == str(g["neigh_sha"])
  state = np.random.get_state()
  assert int(state[2]) == int(g["rng_pos"])
  assert hashlib.sha256(state[1].tobytes()).hexdigest() == str(g["rng_key_sha"])
  if "col_left_node" in g:
    for k in INDEX_KEYS + NEIGH_KEYS:
      assert np.array_equal(arrays[k], g["col_" + k]), k
    for k in ("nn_prob", "ee_prob", "ne_prob"):
      assert np.array_equal(np.isnan(arrays[k]), np.isnan(g["col_" + k]))
      assert np.array_equal(np.nan_to_num(arrays[k]), np.nan_to_num(g["col_" + k]))


def test_per_row_counts_follow_the_weights():
  g = load_golden("boolean_rand25")
  hg = hypergraph_from_pairs(g["pairs"])
  for n in hg.node:
    hg.node[n].weight = 0.5 if n % 2 else 2.0
  A = port.to_csr(hg)
  np.random.seed(3)
  out = BooleanSamples(hg, 2, 4, disable_pbar=True)
  np.random.seed(3)
  want = port.boolean_samples(A, A.T.tocsr(), list(hg.node), list(hg.edge),
                              [d.weight for _, d in hg.node.items()],
                              [d.weight for _, d in hg.edge.items()], 2, 4)
  arrays = out.arrays()
  for k in INDEX_KEYS + NEIGH_KEYS:
    assert np.array_equal(arrays[k], want[k]), k


def test_records_and_model_input_from_fobe_samples():
  g = load_golden("boolean_tiny")
  hg = hypergraph_from_pairs(g["pairs"])
  np.random.seed(int(g["seed"]))
  out = BooleanSamples(hg, int(g["k"]), int(g["num_samples"]), neg_samples=int(g["neg"]))
  feats, targets = SamplesToModelInput(out, int(g["k"]), weighted=False)
  want = port.samples_to_model_input({k[4:]: g[k] for k in g if k.startswith("col_")},
                                     int(g["k"]), weighted=False)
  assert [np.asarray(c).tolist() for c in feats] == want[0]
  assert all(np.allclose(a, b) for a, b in zip(targets, want[1]))
  rec = out[len(out) - 1]
  assert rec.node_edge_prob is None and rec.neighbor_node_indices is not None
