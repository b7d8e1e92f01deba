"""Known-answer vectors the reference's own tests pin for this path, re-run against this
package's host-side mirror (no GPU needed):
  tests/test_hg2v_samples.py:15-189   SamplesToModelInput (6 vectors)
  tests/test_hg2v_weights.py:23-56    UniformWeight, DictToSparseRow
  tests/test_hypergraph_util.py       CSR construction and CompressRange invariants
The expected values below are the reference tests' literals, restated."""
import random

import numpy as np
import pytest
import scipy.sparse as sps
from scipy.sparse import csr_matrix

from hypergraphembedding_b200 import (AddNodeToEdge, CompressRange, DictToSparseRow, Hypergraph,
                                      SampleColumns, SamplesToModelInput, SimilarityRecord,
                                      ToCsrMatrix, ToEdgeCsrMatrix, UniformWeight)
from hypergraphembedding_b200.hg2v_weighting import (AlphaScaleValues, OneMinusValues,
                                                     ZeroOneScaleValues)

CASES = [
    # (record kwargs, num_neighbors, weighted, expected features, expected targets)
    (dict(left_node_idx=0, right_node_idx=1, node_node_prob=0.5), 2, False,
     [[1], [0], [2], [0], [0], [0], [0], [0]], [[0.5], [0], [0]]),
    (dict(left_edge_idx=0, right_edge_idx=1, edge_edge_prob=0.5), 2, False,
     [[0], [1], [0], [2], [0], [0], [0], [0]], [[0], [0.5], [0]]),
    (dict(left_node_idx=0, right_edge_idx=1, neighbor_node_indices=[2],
          neighbor_edge_indices=[3, 4], node_edge_prob=0.5), 2, False,
     [[1], [0], [0], [2], [3], [0], [4], [5]], [[0], [0], [0.5]]),
    (dict(left_node_idx=0, right_node_idx=1, left_weight=0.3, right_weight=0.6,
          node_node_prob=0.5), 2, True,
     [[1], [0], [2], [0], [0.3], [0.6], [0], [0], [0], [0], [0], [0], [0], [0]],
     [[0.5], [0], [0]]),
    (dict(left_edge_idx=0, right_edge_idx=1, left_weight=0.3, right_weight=0.6,
          neighbor_node_indices=[2], neighbor_node_weights=[0.25], neighbor_edge_indices=[3, 4],
          neighbor_edge_weights=[0.5, 0.75], node_edge_prob=0.5), 2, True,
     [[0], [1], [0], [2], [0.3], [0.6], [3], [0], [0.25], [0], [4], [5], [0.5], [.75]],
     [[0], [0], [0.5]]),
    (dict(left_edge_idx=0, right_edge_idx=1, left_weight=0.3, right_weight=0.6,
          edge_edge_prob=0.5), 2, True,
     [[0], [1], [0], [2], [0.3], [0.6], [0], [0], [0], [0], [0], [0], [0], [0]],
     [[0], [0.5], [0]]),
]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_samples_to_model_input_golden_vectors(case):
  kwargs, k, weighted, features, targets = CASES[case]
  actual = SamplesToModelInput([SimilarityRecord(**kwargs)], num_neighbors=k, weighted=weighted)
  assert actual == (features, targets)


def test_samples_to_model_input_default_is_weighted():
  rec = SimilarityRecord(left_node_idx=3, right_node_idx=4, node_node_prob=1)
  assert len(SamplesToModelInput([rec], num_neighbors=1)[0]) == 4 + 2 + 4


def test_columnar_packing_equals_record_packing():
  cols = SampleColumns.concatenate([
      SampleColumns.build(3, 2, left_node=[0, 5], right_node=[1, 6], nn_prob=[0.5, 0.25]),
      SampleColumns.build(3, 1, left_edge=[2], right_edge=[7], ee_prob=[1.0]),
      SampleColumns.build(3, 2, left_node=[4, 0], right_edge=[1, 1], neigh_node=[[1, 2, 3], [0, 0, 9]],
                          neigh_edge=[[4, 4, 4], [5, 6, 7]], ne_prob=[0.75, 0.0]),
      SampleColumns.build(3, 1, left_node=[8], right_node=[9]),      # a negative: no probability
  ])
  for weighted in (False, True):
    for k in (2, 3, 4):
      a = SamplesToModelInput(cols, k, weighted=weighted)
      b = SamplesToModelInput(list(cols), k, weighted=weighted)
      assert len(a[0]) == len(b[0])
      for x, y in zip(a[0] + a[1], b[0] + b[1]):
        assert np.asarray(x).tolist() == [float(v) if isinstance(v, float) else v for v in y]
  assert len(cols) == 6 and cols[2].left_edge_idx == 2 and cols[0].neighbor_node_indices is None
  assert cols[-1].node_node_prob is None and cols[3].neighbor_edge_indices.tolist() == [4, 4, 4]


def _sparse_close(a, b, tol=1e-5):
  assert a.shape == b.shape
  assert not np.max(np.abs(a - b) >= tol)


def test_uniform_weight_typical():
  hg = Hypergraph()
  AddNodeToEdge(hg, 0, 1)
  AddNodeToEdge(hg, 2, 2)
  AddNodeToEdge(hg, 3, 2)
  node2weight, edge2weight = UniformWeight(hg)
  _sparse_close(node2weight, csr_matrix([[0, 1, 0], [0, 0, 0], [0, 0, 1], [0, 0, 1]],
                                        dtype=np.float32))
  _sparse_close(edge2weight, csr_matrix([[0, 0, 0, 0], [1, 0, 0, 0], [0, 0, 1, 1]],
                                        dtype=np.float32))


def test_dict_to_sparse_row_typical():
  _sparse_close(DictToSparseRow({0: 1, 2: 4, 5: 100}),
                csr_matrix([1, 0, 4, 0, 0, 100], dtype=np.float32))


def test_scale_helpers():
  assert ZeroOneScaleValues({}) == {}
  assert ZeroOneScaleValues({3: 7.0, 9: 7.0}) == {3: 1, 9: 1}
  assert ZeroOneScaleValues({0: 1.0, 1: 3.0, 2: 2.0}) == {0: 0.0, 1: 1.0, 2: 0.5}
  assert OneMinusValues({0: 0.25}) == {0: 0.75}
  assert AlphaScaleValues({0: 0.5}, 0.5) == {0: 0.75}
  with pytest.raises(AssertionError):
    AlphaScaleValues({0: 0.5}, 1.5)
  with pytest.raises(AssertionError):
    AlphaScaleValues({0: 0.5}, -0.1)


def test_csr_construction_matches_dense():
  random.seed(3)
  hg = Hypergraph()
  dense = np.zeros((30, 20), dtype=bool)
  for i in range(30):
    for j in range(20):
      if random.random() < 0.2:
        AddNodeToEdge(hg, i, j)
        dense[i, j] = True
  a, b = ToCsrMatrix(hg), ToEdgeCsrMatrix(hg)
  assert np.array_equal(a.toarray(), dense[:a.shape[0], :a.shape[1]])
  assert np.array_equal(b.toarray(), dense.T[:b.shape[0], :b.shape[1]])
  assert a.has_sorted_indices and b.has_sorted_indices
  assert ToCsrMatrix(Hypergraph()).shape == sps.csr_matrix([]).shape


def test_compress_range_invariants():
  hg = Hypergraph()
  hg.name = "sparse ids"
  for n, e in [(10, 700), (10, 50), (999, 50), (4, 700), (4, 3)]:
    AddNodeToEdge(hg, n, e)
  hg.node[999].weight = 0.5
  compressed, inv_node, inv_edge = CompressRange(hg)
  assert compressed.name == "sparse ids"
  assert len(compressed.node) == max(compressed.node) + 1 == 3
  assert len(compressed.edge) == max(compressed.edge) + 1 == 3
  assert inv_node == {0: 4, 1: 10, 2: 999} and inv_edge == {0: 3, 1: 50, 2: 700}
  assert compressed.node[2].weight == 0.5 and compressed.node[0].weight == 1.0
  back = {(inv_node[n], inv_edge[e]) for n, d in compressed.node.items() for e in d.edges}
  assert back == {(10, 700), (10, 50), (999, 50), (4, 700), (4, 3)}
  back_e = {(inv_node[n], inv_edge[e]) for e, d in compressed.edge.items() for n in d.nodes}
  assert back_e == back
  assert not hg.node[4].HasField("weight")      # the input is not mutated


# ---- tests/test_hypergraph_util.py:39-86, 131-179 of the reference, restated -----------------


def test_add_node_to_edge_typical_dupl_and_names():
  actual = Hypergraph()
  AddNodeToEdge(actual, 1, 2)
  AddNodeToEdge(actual, 1, 2)                      # duplicate calls do not change the structure
  expected = Hypergraph()
  expected.node[1].edges.append(2)
  expected.edge[2].nodes.append(1)
  assert actual == expected

  actual = Hypergraph()
  AddNodeToEdge(actual, 0, 0, "A", "X")
  AddNodeToEdge(actual, 1, 0, node_name="B")
  AddNodeToEdge(actual, 1, 1, edge_name="Y")
  expected = Hypergraph()
  expected.node[0].edges.append(0)
  expected.node[0].name = "A"
  expected.node[1].edges.extend([0, 1])
  expected.node[1].name = "B"
  expected.edge[0].nodes.extend([0, 1])
  expected.edge[0].name = "X"
  expected.edge[1].nodes.append(1)
  expected.edge[1].name = "Y"
  assert actual == expected


def test_to_csr_matrix_one_multiple_and_empty():
  one = Hypergraph()
  AddNodeToEdge(one, 1, 2)
  assert np.array_equal(ToCsrMatrix(one).toarray(), [[0, 0, 0], [0, 0, 1]])
  assert np.array_equal(ToEdgeCsrMatrix(one).toarray(), [[0, 0], [0, 0], [0, 1]])
  multiple = Hypergraph()
  AddNodeToEdge(multiple, 1, 1)
  AddNodeToEdge(multiple, 1, 2)
  AddNodeToEdge(multiple, 2, 0)
  assert np.array_equal(ToCsrMatrix(multiple).toarray(), [[0, 0, 0], [0, 1, 1], [1, 0, 0]])
  assert ToCsrMatrix(multiple).dtype == bool
  empty = ToCsrMatrix(Hypergraph())
  assert empty.shape == sps.csr_matrix([]).shape and empty.nnz == 0
  assert ToEdgeCsrMatrix(Hypergraph()).nnz == 0
