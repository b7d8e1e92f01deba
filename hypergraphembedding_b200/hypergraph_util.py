"""Host-side graph utilities on the HOBE hot path: the proto <-> incidence-array boundary.

Mirrors the part of the reference's ``hypergraph_util.py`` that the path calls
(``AddNodeToEdge``:13, ``ToCsrMatrix``:96, ``ToEdgeCsrMatrix``:117, ``Relabel``:198,
``CompressRange``:223), with the same names, arguments and results, and adds the array
forms the CUDA kernels consume (int64 row pointers + int32 sorted column ids).
"""
import ctypes
import logging

import numpy as np
import scipy.sparse as sps

from . import _native
from .hypergraph_pb2 import Hypergraph

log = logging.getLogger()


def AddNodeToEdge(hypergraph, node_idx, edge_idx, node_name=None, edge_name=None):
  """hypergraph_util.py:13-44: records node_idx in edge_idx (both directions, no
  duplicates) and optionally names them."""
  assert node_idx >= 0
  assert edge_idx >= 0
  node = hypergraph.node[node_idx]
  edge = hypergraph.edge[edge_idx]
  if edge_idx not in node.edges:
    node.edges.append(edge_idx)
  if node_idx not in edge.nodes:
    edge.nodes.append(node_idx)
  if node_name is not None:
    if node.HasField("name") and node.name != node_name:
      log.warning("Overwriting Node #{} name from {} to {}".format(node_idx, node.name, node_name))
    node.name = node_name
  if edge_name is not None:
    if edge.HasField("name") and edge.name != edge_name:
      log.warning("Overwriting Edge #{} name from {} to {}".format(edge_idx, edge.name, edge_name))
    edge.name = edge_name
  return hypergraph


def IsEmpty(hypergraph):
  """hypergraph_util.py:91-93."""
  return len(hypergraph.node) == 0 or len(hypergraph.edge) == 0


def _coo_to_csr(rows, cols, shape=None):
  rows = np.asarray(rows, dtype=np.int64)
  cols = np.asarray(cols, dtype=np.int64)
  vals = np.ones(len(rows), dtype=bool)
  if shape is None:
    return sps.csr_matrix((vals, (rows, cols)), dtype=bool)
  return sps.csr_matrix((vals, (rows, cols)), shape=shape, dtype=bool)


def ToCsrMatrix(hypergraph):
  """hypergraph_util.py:96-114: N x E bool CSR from ``node.edges``; shape = max id + 1;
  duplicates collapse; column ids sorted.  The incidences are read from the serialized
  message (csrc/hge_proto.cpp), not by a Python loop over them."""
  if IsEmpty(hypergraph):
    return sps.csr_matrix([])
  arrays = HypergraphArrays(hypergraph)
  rows = np.repeat(arrays.node_ids, np.diff(arrays.node_ptr))
  cols = arrays.node_edges
  arrays.close()
  m = _coo_to_csr(rows, cols)
  m.sum_duplicates()
  return m


def ToEdgeCsrMatrix(hypergraph):
  """hypergraph_util.py:117-135: E x N bool CSR from ``edge.nodes``."""
  if IsEmpty(hypergraph):
    return sps.csr_matrix([])
  arrays = HypergraphArrays(hypergraph)
  rows = np.repeat(arrays.edge_ids, np.diff(arrays.edge_ptr))
  cols = arrays.edge_nodes
  arrays.close()
  m = _coo_to_csr(rows, cols)
  m.sum_duplicates()
  return m


def Relabel(original_hg, node_map, edge_map):
  """hypergraph_util.py:198-220: new hypergraph with ids mapped through node_map / edge_map.
  Only ``node.edges`` drives the connections; weights of every node and edge are copied
  (which also materialises entries for unreferenced edges)."""
  relabeled = Hypergraph()
  if original_hg.HasField("name"):
    relabeled.name = original_hg.name
  # Same result as calling AddNodeToEdge per incidence, without its O(deg) membership scans.
  node_seen = {}
  edge_seen = {}
  for node_idx, src in original_hg.node.items():
    for edge_idx in src.edges:
      assert node_idx in node_map
      assert edge_idx in edge_map
      n, e = node_map[node_idx], edge_map[edge_idx]
      assert n >= 0
      assert e >= 0
      node = relabeled.node[n]   # map access creates the entry, as AddNodeToEdge does
      edge = relabeled.edge[e]
      seen = node_seen.setdefault(n, set())
      if e not in seen:
        seen.add(e)
        node.edges.append(e)
      seen = edge_seen.setdefault(e, set())
      if n not in seen:
        seen.add(n)
        edge.nodes.append(n)
  for node_idx, node in original_hg.node.items():
    relabeled.node[node_map[node_idx]].weight = node.weight
  for edge_idx, edge in original_hg.edge.items():
    relabeled.edge[edge_map[edge_idx]].weight = edge.weight
  return relabeled


def CompressRange(original_hg):
  """hypergraph_util.py:223-244: ids moved into 0..n-1 by sorted order; returns the new
  hypergraph and the INVERSE node / edge maps (compressed -> original)."""
  node_indices = sorted(original_hg.node)
  edge_indices = sorted(original_hg.edge)
  node_map = {n: i for i, n in enumerate(node_indices)}
  edge_map = {e: i for i, e in enumerate(edge_indices)}
  compressed = Relabel(original_hg, node_map, edge_map)
  inv_node_map = {y: x for x, y in node_map.items()}
  inv_edge_map = {y: x for x, y in edge_map.items()}
  return compressed, inv_node_map, inv_edge_map


# ---------------------------------------------------------------------------------------
# array forms consumed by the CUDA path
# ---------------------------------------------------------------------------------------


def csr_arrays(matrix):
  """(int64 row pointers, int32 sorted unique column ids) of a scipy sparse matrix."""
  m = sps.csr_matrix(matrix)
  if not m.has_canonical_format:
    m.sum_duplicates()
  return (np.ascontiguousarray(m.indptr, dtype=np.int64),
          np.ascontiguousarray(m.indices, dtype=np.int32))


def incidence_arrays(hypergraph):
  """(num_nodes, num_edges, n2e_ptr, n2e_idx, e2n_ptr, e2n_idx) exactly as
  ToCsrMatrix / ToEdgeCsrMatrix would lay them out, padded to a common shape
  (max node id + 1) x (max edge id + 1) so row counts agree between the two."""
  a = ToCsrMatrix(hypergraph)
  b = ToEdgeCsrMatrix(hypergraph)
  num_nodes = max(a.shape[0], b.shape[1])
  num_edges = max(a.shape[1], b.shape[0])
  a = sps.csr_matrix((a.data, a.indices, _pad_ptr(a.indptr, num_nodes)),
                     shape=(num_nodes, num_edges))
  b = sps.csr_matrix((b.data, b.indices, _pad_ptr(b.indptr, num_edges)),
                     shape=(num_edges, num_nodes))
  return (num_nodes, num_edges) + csr_arrays(a) + csr_arrays(b)


def _pad_ptr(indptr, rows):
  indptr = np.asarray(indptr)
  if len(indptr) - 1 >= rows:
    return indptr
  return np.concatenate([indptr, np.full(rows - (len(indptr) - 1), indptr[-1], indptr.dtype)])


def compressed_incidence(hypergraph):
  """Array form of ``CompressRange`` followed by ``ToCsrMatrix`` / ``ToEdgeCsrMatrix``
  (algebraic_distance.py:133-146) without materialising the relabelled proto.

  Returns (node_ids, edge_ids, A) with node_ids / edge_ids the sorted original ids
  (= the inverse maps) and A the N x E bool CSR; after Relabel the edge->node matrix is
  always A.T (only ``node.edges`` drives the connections)."""
  node_ids = np.asarray(sorted(hypergraph.node), dtype=np.int64)
  edge_ids = np.asarray(sorted(hypergraph.edge), dtype=np.int64)
  rows, cols = [], []
  for node_idx, node in hypergraph.node.items():
    edges = node.edges
    rows.extend([node_idx] * len(edges))
    cols.extend(edges)
  rows = np.asarray(rows, dtype=np.int64)
  cols = np.asarray(cols, dtype=np.int64)
  r = np.searchsorted(node_ids, rows)
  c = np.searchsorted(edge_ids, cols)
  # Relabel asserts that every referenced edge id is a key of hypergraph.edge
  assert len(edge_ids) > 0 and np.all(c < len(edge_ids))
  assert np.all(edge_ids[c] == cols)
  a = _coo_to_csr(r, c, shape=(len(node_ids), len(edge_ids)))
  a.sum_duplicates()
  return node_ids, edge_ids, a


# ---------------------------------------------------------------------------------------
# wire-format fast path (csrc/hge_proto.cpp): serialized proto <-> arrays without a Python
# loop over incidences or rows
# ---------------------------------------------------------------------------------------


def _wire_bytes(message_or_bytes):
  if isinstance(message_or_bytes, (bytes, bytearray, memoryview)):
    return bytes(message_or_bytes)
  return message_or_bytes.SerializeToString()


class HypergraphArrays(object):
  """A serialized ``Hypergraph`` read into arrays.  Per side: ids in wire order (whatever order
  the producer's serializer emitted its map in -- not necessarily the order Python iterates the
  map; take ``list(hypergraph.node)`` where that order matters), int64 row pointers, members in stored order (duplicates kept, as
  in the proto), fp32 weights (default 1)."""

  def __init__(self, message_or_bytes):
    lib = _native.load_library()
    data = _wire_bytes(message_or_bytes)
    handle = _native.c_vp()
    _native.check(lib.hge_hypergraph_parse(data, len(data), ctypes.byref(handle)),
                  "hge_hypergraph_parse")
    self._lib, self._handle = lib, handle
    sizes = np.zeros(4, dtype=np.int64)
    _native.check(lib.hge_hypergraph_sizes(handle, _native.ptr(sizes)))
    self.num_node_entries, self.num_edge_entries, n_members, e_members = (int(v) for v in sizes)
    self.node_ids = np.empty(self.num_node_entries, np.int32)
    self.node_ptr = np.empty(self.num_node_entries + 1, np.int64)
    self.node_edges = np.empty(n_members, np.int32)
    self.node_weight = np.empty(self.num_node_entries, np.float32)
    self.edge_ids = np.empty(self.num_edge_entries, np.int32)
    self.edge_ptr = np.empty(self.num_edge_entries + 1, np.int64)
    self.edge_nodes = np.empty(e_members, np.int32)
    self.edge_weight = np.empty(self.num_edge_entries, np.float32)
    _native.check(lib.hge_hypergraph_arrays(
        handle, *[_native.ptr(a) for a in (self.node_ids, self.node_ptr, self.node_edges,
                                           self.node_weight, self.edge_ids, self.edge_ptr,
                                           self.edge_nodes, self.edge_weight)]))

  def compress(self):
    """CompressRange + ToCsrMatrix + transpose (algebraic_distance.py:133-146) on the arrays:
    (sorted node ids, sorted edge ids, n2e_ptr, n2e_idx, e2n_ptr, e2n_idx).  Raises
    AssertionError when a node lists an edge that is not a key of the edge map (Relabel
    asserts the same, hypergraph_util.py:207)."""
    n, e = self.num_node_entries, self.num_edge_entries
    cap = max(1, len(self.node_edges))
    node_ids, edge_ids = np.empty(n, np.int32), np.empty(e, np.int32)
    a_ptr, a_idx = np.empty(n + 1, np.int64), np.empty(cap, np.int32)
    b_ptr, b_idx = np.empty(e + 1, np.int64), np.empty(cap, np.int32)
    nnz = ctypes.c_int64(0)
    _native.check(self._lib.hge_hypergraph_compress(
        self._handle, _native.ptr(node_ids), _native.ptr(edge_ids), _native.ptr(a_ptr),
        _native.ptr(a_idx), _native.ptr(b_ptr), _native.ptr(b_idx), ctypes.byref(nnz)),
                  "hge_hypergraph_compress")
    return node_ids, edge_ids, a_ptr, a_idx[:nnz.value], b_ptr, b_idx[:nnz.value]

  def close(self):
    if getattr(self, "_handle", None):
      self._lib.hge_hypergraph_destroy(self._handle)
      self._handle = None

  def __del__(self):
    try:
      self.close()
    except Exception:
      pass


def embedding_to_wire(node_ids, node_vectors, edge_ids, edge_vectors, dim, method_name):
  """Serialized ``HypergraphEmbedding`` from dense fp32 rows keyed by strictly ascending ids:
  field for field what the protobuf runtime emits for the reference's message (map entries in
  ascending key order)."""
  lib = _native.load_library()
  node_ids = np.ascontiguousarray(node_ids, dtype=np.int32)
  edge_ids = np.ascontiguousarray(edge_ids, dtype=np.int32)
  xn = np.ascontiguousarray(node_vectors, dtype=np.float32)
  xe = np.ascontiguousarray(edge_vectors, dtype=np.float32)
  R = int(xn.shape[1]) if xn.ndim == 2 else 0
  assert xn.shape == (len(node_ids), R) and xe.shape == (len(edge_ids), R)
  name = method_name.encode("utf-8") if method_name is not None else None
  size = ctypes.c_size_t(0)
  _native.check(lib.hge_embedding_wire_size(_native.ptr(node_ids), len(node_ids),
                                            _native.ptr(edge_ids), len(edge_ids), R, int(dim), name,
                                            ctypes.byref(size)), "hge_embedding_wire_size")
  out = np.empty(size.value, dtype=np.uint8)
  written = ctypes.c_size_t(0)
  _native.check(lib.hge_embedding_write(_native.ptr(node_ids), len(node_ids), _native.ptr(xn),
                                        _native.ptr(edge_ids), len(edge_ids), _native.ptr(xe), R,
                                        int(dim), name, _native.ptr(out), out.size,
                                        ctypes.byref(written)), "hge_embedding_write")
  assert written.value == size.value
  return out.tobytes()


def embedding_from_wire(message_or_bytes):
  """A serialized ``HypergraphEmbedding`` as (node_ids, node_ptr, node_values, edge_ids, edge_ptr,
  edge_values, dim): ids in wire order, row pointers into the flat fp32 value arrays, dim = None
  when the field is absent."""
  lib = _native.load_library()
  data = _wire_bytes(message_or_bytes)
  handle = _native.c_vp()
  _native.check(lib.hge_embedding_parse(data, len(data), ctypes.byref(handle)), "hge_embedding_parse")
  try:
    sizes = np.zeros(4, dtype=np.int64)
    dim = ctypes.c_int32(0)
    _native.check(lib.hge_embedding_sizes(handle, _native.ptr(sizes), ctypes.byref(dim)))
    n, e, nv, ev = (int(v) for v in sizes)
    out = (np.empty(n, np.int32), np.empty(n + 1, np.int64), np.empty(nv, np.float32),
           np.empty(e, np.int32), np.empty(e + 1, np.int64), np.empty(ev, np.float32))
    _native.check(lib.hge_embedding_arrays(handle, *[_native.ptr(a) for a in out]))
  finally:
    lib.hge_embedding_destroy(handle)
  return out + (None if dim.value < 0 else int(dim.value),)
