"""Drop-in for the reference's ``hypergraph_embedding/algebraic_distance.py``.

``EmbedAlgebraicDistance`` keeps the reference signature and semantics
(algebraic_distance.py:126-175): ids are compressed by sorted order, the initial vectors are
drawn from the process-global legacy numpy RNG (nodes first, then edges, :140-141), the
relaxation runs for ``iterations`` sweeps, and the result is a ``HypergraphEmbedding`` keyed
by the original ids with ``method_name == "AlgebraicDistance"`` (:168).

The sweeps themselves run in libhge_b200.so (csrc/hge_algdist.cu) in fp32; there is no CPU
implementation in this package.
"""
import logging

import numpy as np

from . import _native
from .hypergraph_pb2 import HypergraphEmbedding
from .hypergraph_util import HypergraphArrays, csr_arrays, embedding_to_wire

log = logging.getLogger()


def relax(incidence, node_vectors, edge_vectors, iterations, ctx=None, lohi=None):
  """Runs the relaxation in place on fp32 [N, R] / [E, R] arrays (numpy on the host, or torch
  CUDA tensors that stay on the device).  ``incidence`` is a ``_native.Incidence``."""
  ctx = ctx or incidence.ctx
  return _native.algdist_run(ctx, incidence, node_vectors, edge_vectors, iterations, lohi=lohi)


def relax_csr(n2e_csr, node_vectors, edge_vectors, iterations, ctx=None, lohi=None):
  """Incidence set-up + relaxation in one library call (hge_algdist_run_csr), in place on fp32
  numpy arrays: what a caller that relaxes one hypergraph once wants -- the vectors' upload
  overlaps the set-up kernels.  The edge -> node orientation is built on the device."""
  ctx = ctx or _native.default_context()
  a_ptr, a_idx = csr_arrays(n2e_csr)
  return _native.algdist_run_csr(ctx, n2e_csr.shape[0], n2e_csr.shape[1], a_ptr, a_idx, node_vectors,
                                 edge_vectors, iterations, lohi=lohi)


def make_incidence(n2e_csr, e2n_csr=None, ctx=None):
  """Uploads a scipy N x E incidence matrix (and its E x N counterpart, default: transpose)."""
  ctx = ctx or _native.default_context()
  if e2n_csr is None:
    e2n_csr = n2e_csr.T.tocsr()
  a_ptr, a_idx = csr_arrays(n2e_csr)
  b_ptr, b_idx = csr_arrays(e2n_csr)
  return _native.Incidence(ctx, n2e_csr.shape[0], n2e_csr.shape[1], a_ptr, a_idx, b_ptr, b_idx)


def _helper_update_embeddings(hypergraph, node_embeddings, edge_embeddings, node2edges, edge2nodes,
                              workers=None, disable_pbar=True):
  """algebraic_distance.py:54-91: one un-rescaled sweep -- nodes placed with respect to the old
  edge vectors, then edges with respect to the new node vectors.  Returns new arrays."""
  del hypergraph, workers, disable_pbar
  ctx = _native.default_context()
  a_ptr, a_idx = csr_arrays(node2edges)
  b_ptr, b_idx = csr_arrays(edge2nodes)
  xn = np.ascontiguousarray(node_embeddings, dtype=np.float32).copy()
  xe = np.ascontiguousarray(edge_embeddings, dtype=np.float32).copy()
  inc = _native.Incidence(ctx, xn.shape[0], xe.shape[0], a_ptr, a_idx, b_ptr, b_idx)
  try:
    state = _native.AlgDistState(ctx, inc, xn.shape[1], 1)
    try:
      state.load(xn, xe)
      state.node_half(0)
      state.edge_half(0)
      state.store(0, xn, xe)       # sweeps_done = 0: no rescale applied
    finally:
      state.close()
  finally:
    inc.close()
  return xn, xe


def _helper_scale_embeddings(hypergraph, node_embeddings, edge_embeddings, workers=None,
                             disable_pbar=True):
  """algebraic_distance.py:97-123: joint per-column min-max rescale to the unit hypercube, in
  place on fp32 arrays (a copy is returned for other dtypes)."""
  del hypergraph, workers, disable_pbar
  ctx = _native.default_context()
  xn = np.ascontiguousarray(node_embeddings, dtype=np.float32)
  xe = np.ascontiguousarray(edge_embeddings, dtype=np.float32)
  assert xn.shape[1] == xe.shape[1]
  _native.check(ctx.lib.hge_column_rescale(ctx.handle, _native.ptr(xn), xn.shape[0], _native.ptr(xe),
                                           xe.shape[0], xn.shape[1], _native.MEM_HOST),
                "hge_column_rescale")
  return xn, xe


def EmbedAlgebraicDistance(hypergraph,
                           dimension,
                           iterations=20,
                           run_in_parallel=True,
                           disable_pbar=False):
  """algebraic_distance.py:126-175.  ``run_in_parallel`` / ``disable_pbar`` are accepted for
  compatibility; the work is one GPU call either way."""
  del run_in_parallel, disable_pbar
  # proto -> arrays through the wire format (csrc/hge_proto.cpp): CompressRange, ToCsrMatrix and
  # the transpose without a Python loop over incidences
  arrays = HypergraphArrays(hypergraph)
  try:
    node_ids, edge_ids, a_ptr, a_idx, b_ptr, b_idx = arrays.compress()
  finally:
    arrays.close()
  num_nodes = len(node_ids)   # == max(compressed ids) + 1, algebraic_distance.py:135-136
  num_edges = len(edge_ids)

  log.info("Random Initialization")
  # all embeddings are in 0-1 interval; same draws as the reference (f64, nodes then edges)
  node_embeddings = np.random.random((num_nodes, dimension)).astype(np.float32)
  edge_embeddings = np.random.random((num_edges, dimension)).astype(np.float32)

  log.info("Performing iterations of Algebraic Distance Calculations")
  # incidence upload, schedules and the sweeps in one library call (hge_algdist_run_csr)
  _native.algdist_run_csr(_native.default_context(), num_nodes, num_edges, a_ptr, a_idx, node_embeddings,
                          edge_embeddings, iterations, e2n_ptr=b_ptr, e2n_idx=b_idx)

  # arrays -> proto through the wire format: no per-row packing loop (algebraic_distance.py:169-174)
  embedding = HypergraphEmbedding()
  embedding.ParseFromString(embedding_to_wire(node_ids, node_embeddings, edge_ids, edge_embeddings,
                                              dimension, "AlgebraicDistance"))
  return embedding
