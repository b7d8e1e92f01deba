"""HOBE sampling (AlgebraicDistanceSamples): pair sets and neighbour arrays bit-exact, weighted
probabilities within 1e-5 of the unmodified reference (committed golden outputs)."""
import hashlib

import numpy as np
import pytest

from conftest import hypergraph_from_pairs, load_golden

pytestmark = pytest.mark.gpu

INDEX_KEYS = ("left_node", "left_edge", "right_node", "right_edge")
NEIGH_KEYS = ("neigh_node", "neigh_edge")
RTOL, ATOL = 1e-5, 1e-6


def _sha(arrays, keys):
  h = hashlib.sha256()
  for k in keys:
    h.update(np.ascontiguousarray(arrays[k], dtype=np.int64).tobytes())
  return h.hexdigest()


def _embedding(xn, xe):
  from hypergraphembedding_b200 import HypergraphEmbedding
  emb = HypergraphEmbedding()
  emb.dim = xn.shape[1]
  for i in range(xn.shape[0]):
    emb.node[i].values.extend(xn[i].tolist())
  for i in range(xe.shape[0]):
    emb.edge[i].values.extend(xe[i].tolist())
  return emb


def _run(g, parallel):
  from hypergraphembedding_b200 import AlgebraicDistanceSamples
  hg = hypergraph_from_pairs(g["pairs"])
  assert list(hg.node) == g["node_rows"].tolist() and list(hg.edge) == g["edge_rows"].tolist()
  np.random.seed(int(g["seed"]))
  out = AlgebraicDistanceSamples(hg, _embedding(g["xn"], g["xe"]), int(g["k"]),
                                 int(g["num_samples"]), run_in_parallel=parallel, disable_pbar=True)
  state = np.random.get_state()
  assert int(state[2]) == int(g["rng_pos"])
  assert hashlib.sha256(state[1].tobytes()).hexdigest() == str(g["rng_key_sha"])
  assert len(out) == int(g["count"])
  return out


@pytest.mark.parametrize("name", ["tiny", "rand25", "youtube_s2"])
def test_hobe_samples_match_reference(name):
  g = load_golden("hobe_" + name)
  out = _run(g, parallel=False)
  arrays = out.arrays()
  for k in INDEX_KEYS + NEIGH_KEYS:
    assert np.array_equal(arrays[k], g["col_" + k]), k          # bit-exact sample sets
  assert _sha(arrays, INDEX_KEYS) == str(g["index_sha"])
  assert _sha(arrays, NEIGH_KEYS) == str(g["neigh_sha"])
  for k in ("nn_prob", "ee_prob", "ne_prob"):
    assert np.array_equal(np.isnan(arrays[k]), np.isnan(g["col_" + k])), k
    got, want = np.nan_to_num(arrays[k]), np.nan_to_num(g["col_" + k])
    assert np.all(np.abs(got - want) <= RTOL * np.abs(want) + ATOL), k
    assert np.all((got >= 0) & (got <= 1))


def test_hobe_default_config_on_the_fixture():
  """BASELINE.json configs[0]: 5 neighbours, 200 samples per row -> 755 267 records."""
  g = load_golden("hobe_youtube_s200")
  out = _run(g, parallel=True)
  arrays = out.arrays()
  assert _sha(arrays, INDEX_KEYS) == str(g["index_sha"])
  kinds = np.where(~np.isnan(arrays["nn_prob"]), 0, np.where(~np.isnan(arrays["ee_prob"]), 1, 2))
  assert np.bincount(kinds, minlength=3).tolist() == g["kind_counts"].tolist()
  prob = np.where(kinds == 0, arrays["nn_prob"], np.where(kinds == 1, arrays["ee_prob"],
                                                          arrays["ne_prob"]))
  want = g["prob_strided"]
  got = prob[::int(g["stride"])]
  assert np.all(np.abs(got - want) <= RTOL * np.abs(want) + ATOL)


def test_model_input_from_hobe_samples():
  from hypergraphembedding_b200 import SamplesToModelInput
  from oracle import port
  g = load_golden("hobe_rand25")
  out = _run(g, parallel=False)
  feats, targets = SamplesToModelInput(out, int(g["k"]), weighted=False)
  want = port.samples_to_model_input({k[4:]: g[k] for k in g if k.startswith("col_")}, int(g["k"]))
  assert [np.asarray(c).tolist() for c in feats] == want[0]
  for a, b in zip(targets, want[1]):
    assert np.allclose(a, b, rtol=RTOL, atol=ATOL)
  assert len(feats) == 4 + 2 * int(g["k"]) and all(len(c) == len(out) for c in feats)
