"""Drop-in for the reference's ``hypergraph_embedding/hg2v_weighting.py``: the weighting schemes
that map a hypergraph (and a reference embedding) to node->edge / edge->node weight matrices.

``WeightByDistance``:67, ``WeightBySameTypeDistance``:34, ``WeightByDistanceCluster``:106,
``WeightByNeighborhood``:137, ``WeightByAlgebraicSpan``:170, ``UniformWeight``:195,
``ComputeSpans``:236, ``ZeroOneScaleValues``:301, ``OneMinusValues``:325,
``AlphaScaleValues``:329, ``DictToSparseRow``:336 -- same names, arguments and results.

Distances, the min/max scale transform and the spans run in libhge_b200.so
(csrc/hge_weighting.cu); assembling the scipy result matrices is host bookkeeping.  ``norm``
must be the L2 norm (``np.linalg.norm``): any other callable raises instead of silently
computing something else.
"""
import logging

import numpy as np
import scipy.sparse as sps
from scipy.sparse import csr_matrix, lil_matrix

from . import _native
from .algebraic_distance import EmbedAlgebraicDistance
from .hg2v_sample import _Graph, embedding_to_arrays
from .hypergraph_util import ToCsrMatrix, ToEdgeCsrMatrix

log = logging.getLogger()


def _require_l2(norm):
  if norm is np.linalg.norm or norm is None:
    return
  probe = np.asarray([3.0, -4.0, 12.0], dtype=np.float32)
  try:
    ok = abs(float(norm(probe)) - 13.0) < 1e-5 and abs(float(norm(-2 * probe)) - 26.0) < 1e-5
  except Exception:
    ok = False
  if not ok:
    raise NotImplementedError(
        "hypergraphembedding_b200 computes L2 distances on the GPU; norm=%r is not the L2 norm"
        % (norm,))


def _drop_zeros(ptr, idx, vals, shape):
  """scipy's lil_matrix does not store an assigned 0, so exact zeros vanish from the result
  (hg2v_weighting.py:98-103; with alpha = 0 the farthest pair gets weight exactly 0)."""
  m = csr_matrix((vals, idx, ptr), shape=shape, dtype=np.float32)
  m.eliminate_zeros()
  return m


def WeightByDistance(hypergraph, alpha, ref_embedding, norm, disable_pbar):
  """hg2v_weighting.py:67-103: every node-edge incidence weighted by
  alpha + (1 - alpha) * (1 - zero_one(||x_node - x_edge||)).  Returns (csr N x E, csr E x N)."""
  del disable_pbar
  _require_l2(norm)
  log.info("Getting largest indices")
  num_nodes = max(hypergraph.node) + 1
  num_edges = max(hypergraph.edge) + 1
  g = _Graph(hypergraph)
  assert g.num_nodes <= num_nodes and g.num_edges <= num_edges
  xn, xe = embedding_to_arrays(ref_embedding, g.num_nodes, g.num_edges)
  ctx = _native.default_context()
  inc = g.incidence(ctx)
  try:
    log.info("Getting distances")
    dist = _native.incidence_l2(ctx, inc, xn, xe, order=0)
    log.info("Scaling distances")
    _native.scale_transform(ctx, dist, alpha)
  finally:
    inc.close()
  log.info("Recording results in matrix")
  ptr = np.concatenate([g.a.ptr, np.full(num_nodes - g.num_nodes, g.a.ptr[-1], np.int64)])
  node2edge_dist = _drop_zeros(ptr, g.a.idx, dist, (num_nodes, num_edges))
  return node2edge_dist, csr_matrix(node2edge_dist.T)


def WeightBySameTypeDistance(hypergraph, alpha, ref_embedding, norm, disable_pbar):
  """hg2v_weighting.py:34-64: the same transform over every stored entry of A * A.T (node-node,
  diagonal included) and B * B.T (edge-edge).  Returns (csr N x N, csr E x E)."""
  del disable_pbar
  _require_l2(norm)
  g = _Graph(hypergraph)
  xn, xe = embedding_to_arrays(ref_embedding, g.num_nodes, g.num_edges)
  ctx = _native.default_context()

  def do_half(m, mt, x):
    rows = np.arange(m.shape[0], dtype=np.int32)
    ptr, idx = _native.spgemm_rows((m, mt), rows, sorted_rows=True)
    left = np.repeat(rows, np.diff(ptr)).astype(np.int32)
    log.info("Calculating distances")
    dist = _native.pair_l2(ctx, x, x, left, idx)
    log.info("Scaling")
    _native.scale_transform(ctx, dist, alpha)
    log.info("Converting")
    return _drop_zeros(ptr, idx, dist, (m.shape[0], m.shape[0]))

  log.info("Identifying all node-node relationships")
  node2node_dist = do_half(g.a, g.at, xn)
  log.info("Identifying all edge-edge relationships")
  edge2edge_dist = do_half(g.b, g.bt, xe)
  return node2node_dist, edge2edge_dist


def WeightByDistanceCluster(hypergraph, alpha, ref_embedding, norm, dim):
  """hg2v_weighting.py:106-134: WeightByDistance followed by sklearn's NMF of the weight
  matrix.  The factorisation itself is third-party and outside the GPU path."""
  node2edge_dist, _ = WeightByDistance(hypergraph, alpha, ref_embedding, norm, True)
  from sklearn.decomposition import NMF
  log.info("Clustering...")
  nmf_model = NMF(dim)
  W = nmf_model.fit_transform(node2edge_dist)
  H = nmf_model.components_
  return csr_matrix(W), csr_matrix(H.T)


def _scaled_row(idx2value, alpha):
  return DictToSparseRow(AlphaScaleValues(OneMinusValues(ZeroOneScaleValues(idx2value)), alpha))


def WeightByNeighborhood(hypergraph, alpha):
  """hg2v_weighting.py:137-167: larger neighbourhoods contribute less."""
  log.info("Getting neighboorhood sizes for all nodes / edges")
  node_neighborhood = {idx: len(node.edges) for idx, node in hypergraph.node.items()}
  edge_neighborhood = {idx: len(edge.nodes) for idx, edge in hypergraph.edge.items()}
  node_row = _scaled_row(node_neighborhood, alpha)
  edge_row = _scaled_row(edge_neighborhood, alpha)
  node2weight = ToCsrMatrix(hypergraph).astype(np.float32).multiply(edge_row)
  edge2weight = ToEdgeCsrMatrix(hypergraph).astype(np.float32).multiply(node_row)
  return node2weight, edge2weight


def WeightByAlgebraicSpan(hypergraph, alpha):
  """hg2v_weighting.py:170-192."""
  node_span, edge_span = ComputeSpans(hypergraph)
  node_row = _scaled_row(node_span, alpha)
  edge_row = _scaled_row(edge_span, alpha)
  node2weight = ToCsrMatrix(hypergraph).astype(np.float32).multiply(edge_row)
  edge2weight = ToEdgeCsrMatrix(hypergraph).astype(np.float32).multiply(node_row)
  return node2weight, edge2weight


def UniformWeight(hypergraph):
  """hg2v_weighting.py:195-198."""
  return (ToCsrMatrix(hypergraph).astype(np.float32), ToEdgeCsrMatrix(hypergraph).astype(np.float32))


################################################################################
# ComputeSpans                                                                 #
################################################################################


def _compute_span(idx, idx2neighbors=None, idx_emb=None, neigh_emb=None):
  """hg2v_weighting.py:214-233 for one row with explicit arguments (ComputeSpans does all rows
  in one launch): (idx, max(0, max diff) - min(0, min diff)) over the row's neighbours."""
  assert idx2neighbors is not None and idx_emb is not None and neigh_emb is not None
  m = csr_matrix(idx2neighbors)
  if idx not in idx_emb or idx >= m.shape[0]:
    return idx, 0
  cols = [int(c) for c in m[idx].indices if int(c) in neigh_emb]
  if not cols:
    return idx, 0
  dim = len(idx_emb[idx].values)
  x_self = np.asarray([idx_emb[idx].values], dtype=np.float32)
  x_other = np.asarray([neigh_emb[c].values for c in cols], dtype=np.float32).reshape(len(cols), dim)
  ctx = _native.default_context()
  ptr = np.asarray([0, len(cols)], np.int64)
  ids = np.arange(len(cols), dtype=np.int32)
  t_ptr = np.arange(len(cols) + 1, dtype=np.int64)
  inc = _native.Incidence(ctx, 1, len(cols), ptr, ids, t_ptr, np.zeros(len(cols), np.int32))
  try:
    span = _native.row_span(ctx, inc, x_self, x_other, 0)[0]
  finally:
    inc.close()
  return idx, span


def ComputeSpans(hypergraph, embedding=None, run_in_parallel=True, disable_pbar=False):
  """hg2v_weighting.py:236-293: for every node / edge the spread of its neighbours around it,
  max(0, max(x_neigh - x_self)) - min(0, min(x_neigh - x_self)) over all neighbours and
  components (``_compute_span``:214-233).  Without an embedding, algebraic distance in 5
  dimensions, 10 sweeps, is used (:257-262)."""
  if embedding is None:
    embedding = EmbedAlgebraicDistance(hypergraph, dimension=5, iterations=10,
                                       run_in_parallel=run_in_parallel, disable_pbar=disable_pbar)
  assert set(hypergraph.node) == set(embedding.node)
  assert set(hypergraph.edge) == set(embedding.edge)
  g = _Graph(hypergraph)
  xn, xe = embedding_to_arrays(embedding, g.num_nodes, g.num_edges)
  ctx = _native.default_context()
  inc = g.incidence(ctx)
  try:
    log.info("Computing span per node wrt edge %s", embedding.method_name)
    node_span = _native.row_span(ctx, inc, xn, xe, 0)
    log.info("Computing span per edge wrt node %s", embedding.method_name)
    edge_span = _native.row_span(ctx, inc, xn, xe, 1)
  finally:
    inc.close()
  # a key beyond the incidence matrix (an entity nothing refers to) has span 0, as in the
  # reference's _compute_span (`idx < idx2neighbors.shape[0]`, hg2v_weighting.py:219)
  node2span = {idx: float(node_span[idx]) if idx < len(node_span) else 0 for idx in hypergraph.node}
  edge2span = {idx: float(edge_span[idx]) if idx < len(edge_span) else 0 for idx in hypergraph.edge}
  return node2span, edge2span


################################################################################
# Scaling helpers on dictionaries (host-side, as in the reference)               #
################################################################################


def ZeroOneScaleValues(idx2value, disable_pbar=False):
  """hg2v_weighting.py:301-317: min-max scale to [0, 1]; one distinct value -> all 1; {} -> {}."""
  del disable_pbar
  if len(idx2value) == 0:
    return {}
  lo = min(idx2value.values())
  hi = max(idx2value.values())
  delta = hi - lo
  if delta == 0:
    return {idx: 1 for idx in idx2value}
  return {idx: (val - lo) / delta for idx, val in idx2value.items()}


def OneMinusValues(data):
  """hg2v_weighting.py:325-326."""
  return {k: 1 - v for k, v in data.items()}


def AlphaScaleValues(data, alpha):
  """hg2v_weighting.py:329-333: alpha is a minimum support for a value."""
  assert alpha >= 0
  assert alpha <= 1
  return {k: (alpha + (1 - alpha) * v) for k, v in data.items()}


def DictToSparseRow(idx2val):
  """hg2v_weighting.py:336-341: 1 x (max key + 1) fp32 CSR row."""
  num_cols = max(idx2val)
  tmp = lil_matrix((1, num_cols + 1), dtype=np.float32)
  for idx, val in idx2val.items():
    tmp[0, idx] = val
  return csr_matrix(tmp)
