"""The wire-format reader / writer (csrc/hge_proto.cpp) against the protobuf runtime itself:
a serialized Hypergraph read into arrays equals what walking the parsed message gives
(hypergraph_util.py:96-135, 198-244), and the embedding bytes we write are the runtime's
own serialization of the message the reference builds (algebraic_distance.py:166-174), field for
field, up to the (undefined) order of the map entries.  Host code: no GPU needed."""
import os
import random

import numpy as np
import pytest

from conftest import ROOT, hypergraph_from_pairs, load_golden
from hypergraphembedding_b200 import (AddNodeToEdge, CompressRange, Hypergraph, HypergraphEmbedding,
                                      ToCsrMatrix, ToEdgeCsrMatrix)
from hypergraphembedding_b200.hypergraph_util import (HypergraphArrays, compressed_incidence,
                                                      embedding_from_wire, embedding_to_wire)


def random_hypergraph(seed, nodes=60, edges=40, p=0.1, id_scale=7, named=True):
  rng = random.Random(seed)
  hg = Hypergraph()
  hg.name = "random-%d" % seed
  for n in range(nodes):
    for e in range(edges):
      if rng.random() < p:
        AddNodeToEdge(hg, n * id_scale + 3, e * 5 + 1, "n%d" % n if named else None)
  for n in list(hg.node)[::3]:
    hg.node[n].weight = rng.random() * 3
  for e in list(hg.edge)[::4]:
    hg.edge[e].weight = rng.random() * 3
  return hg


def check_against_message(hg, arrays):
  # same key sets; the wire order of a map is the serializer's business, not the iteration order
  assert sorted(arrays.node_ids.tolist()) == sorted(hg.node)
  assert sorted(arrays.edge_ids.tolist()) == sorted(hg.edge)
  for i, n in enumerate(arrays.node_ids.tolist()):
    assert arrays.node_edges[arrays.node_ptr[i]:arrays.node_ptr[i + 1]].tolist() == list(hg.node[n].edges)
    assert arrays.node_weight[i] == np.float32(hg.node[n].weight)
  for i, e in enumerate(arrays.edge_ids.tolist()):
    assert arrays.edge_nodes[arrays.edge_ptr[i]:arrays.edge_ptr[i + 1]].tolist() == list(hg.edge[e].nodes)
    assert arrays.edge_weight[i] == np.float32(hg.edge[e].weight)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_reader_matches_the_parsed_message(seed):
  hg = random_hypergraph(seed)
  check_against_message(hg, HypergraphArrays(hg))
  check_against_message(hg, HypergraphArrays(hg.SerializeToString(deterministic=True)))


def test_reader_on_the_reference_fixture_bytes():
  g = load_golden("algdist_youtube")
  hg = hypergraph_from_pairs(g["pairs"])
  arrays = HypergraphArrays(hg)
  check_against_message(hg, arrays)
  assert arrays.num_node_entries == 3862 and arrays.num_edge_entries == 50
  assert len(arrays.node_edges) == 4548


@pytest.mark.parametrize("seed", [3, 4])
def test_compress_equals_compress_range_then_to_csr(seed):
  hg = random_hypergraph(seed, id_scale=11)
  hg.node[5000].edges.extend([6, 6, 1])            # duplicates collapse; unsorted input
  hg.edge[6].nodes.append(5000)
  node_ids, edge_ids, a_ptr, a_idx, b_ptr, b_idx = HypergraphArrays(hg).compress()
  compressed, inv_node, inv_edge = CompressRange(hg)
  assert node_ids.tolist() == [inv_node[i] for i in range(len(inv_node))]
  assert edge_ids.tolist() == [inv_edge[i] for i in range(len(inv_edge))]
  A = ToCsrMatrix(compressed)
  A.sort_indices()
  assert np.array_equal(a_ptr[:A.shape[0] + 1], A.indptr) and np.all(a_ptr[A.shape[0]:] == A.nnz)
  assert np.array_equal(a_idx, A.indices)
  B = ToEdgeCsrMatrix(compressed)
  B.sort_indices()
  assert np.array_equal(b_ptr[:B.shape[0] + 1], B.indptr) and np.array_equal(b_idx, B.indices)
  want_ids, want_edges, want = compressed_incidence(hg)
  assert np.array_equal(want.indptr, a_ptr) and np.array_equal(want.indices, a_idx)


def test_compress_rejects_an_edge_that_is_not_in_the_edge_map():
  hg = random_hypergraph(5)
  hg.node[3].edges.append(999999)
  with pytest.raises(AssertionError):
    HypergraphArrays(hg).compress()
  with pytest.raises(AssertionError):
    CompressRange(hg)


def test_packed_ids_unknown_fields_duplicate_keys_and_truncation():
  def varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while v >= 0x80:
      out.append((v & 0x7f) | 0x80)
      v >>= 7
    out.append(v)
    return bytes(out)

  def ld(field, payload):
    return varint(field << 3 | 2) + varint(len(payload)) + payload

  packed = ld(1, varint(4) + varint(300) + varint(-2))          # repeated int32, packed
  data_a = packed + varint(9 << 3 | 0) + varint(77)             # + an unknown varint field
  data_b = varint(1 << 3 | 0) + varint(8) + varint(3 << 3 | 5) + np.float32(2.5).tobytes()
  entry = lambda key, data: ld(1, varint(1 << 3 | 0) + varint(key) + ld(2, data))
  msg = entry(12, data_a) + entry(-7, data_b) + entry(12, data_b) + ld(7, b"ignored")
  a = HypergraphArrays(msg)
  assert a.node_ids.tolist() == [12, -7]
  assert a.node_edges[a.node_ptr[0]:a.node_ptr[1]].tolist() == [8]      # last entry of key 12 wins
  assert a.node_weight.tolist() == [2.5, 2.5]
  hg = Hypergraph()
  hg.ParseFromString(msg)                                        # the runtime agrees
  assert list(hg.node[12].edges) == [8] and list(hg.node[-7].edges) == [8]
  first = HypergraphArrays(entry(12, data_a))
  assert first.node_edges.tolist() == [4, 300, -2] and first.node_weight.tolist() == [1.0]
  with pytest.raises(AssertionError):
    HypergraphArrays(msg[:-3])
  with pytest.raises(AssertionError):
    HypergraphArrays(b"\x0a\xff\xff\xff\xff\xff\xff\xff\xff\xff\xff\x01")


def _read_varint(buf, pos):
  v, shift = 0, 0
  while True:
    b = buf[pos]
    pos += 1
    v |= (b & 0x7f) << shift
    shift += 7
    if not b & 0x80:
      return v, pos


def _top_level_fields(buf):
  out, pos = [], 0
  while pos < len(buf):
    tag, pos = _read_varint(buf, pos)
    if tag & 7 == 2:
      n, pos = _read_varint(buf, pos)
      out.append((tag, bytes(buf[pos:pos + n])))
      pos += n
    else:
      assert tag & 7 == 0
      v, pos = _read_varint(buf, pos)
      out.append((tag, v))
  return out


def _entry_key(field):
  assert field[1][0] == 0x08
  return _read_varint(field[1], 1)[0]


@pytest.mark.parametrize("R", [1, 5, 32])
def test_embedding_bytes_are_the_runtimes_serialization(R):
  rng = np.random.default_rng(R)
  node_ids = np.sort(rng.choice(100000, 300, replace=False)).astype(np.int32)
  edge_ids = np.sort(rng.choice(5000, 40, replace=False)).astype(np.int32)
  edge_ids[0] = 0
  xn = rng.random((300, R)).astype(np.float32)
  xe = rng.random((40, R)).astype(np.float32)
  xn[0, 0] = 0.0
  wire = embedding_to_wire(node_ids, xn, edge_ids, xe, R, "AlgebraicDistance")
  want = HypergraphEmbedding()                                   # algebraic_distance.py:166-174
  want.dim = R
  want.method_name = "AlgebraicDistance"
  for i, n in enumerate(node_ids.tolist()):
    want.node[n].values.extend(xn[i].tolist())
  for i, e in enumerate(edge_ids.tolist()):
    want.edge[e].values.extend(xe[i].tolist())
  # byte-identical to the runtime's own serialization up to the order of the map entries
  # (which the runtime does not define): same multiset of top-level fields
  assert sorted(_top_level_fields(wire)) == sorted(_top_level_fields(want.SerializeToString()))
  entries = [f for f in _top_level_fields(wire) if f[0] == 0x0a]
  assert entries == sorted(entries, key=lambda f: _entry_key(f))   # ours: ascending keys
  got = HypergraphEmbedding()
  got.ParseFromString(wire)
  assert got == want
  ids_n, ptr_n, val_n, ids_e, ptr_e, val_e, dim = embedding_from_wire(got)
  assert dim == R and sorted(ids_n.tolist()) == node_ids.tolist()
  order = np.argsort(ids_n)
  assert np.array_equal(val_n.reshape(-1, R)[order], xn)
  assert np.array_equal(val_e.reshape(-1, R)[np.argsort(ids_e)], xe)
  assert np.array_equal(np.diff(ptr_n), np.full(300, R)) and np.array_equal(np.diff(ptr_e), np.full(40, R))


def test_embedding_writer_argument_checks():
  x = np.zeros((2, 3), np.float32)
  with pytest.raises(AssertionError):
    embedding_to_wire([5, 5], x, [1, 2], x, 3, "m")              # ids must ascend strictly
  wire = embedding_to_wire([], np.zeros((0, 3), np.float32), [], np.zeros((0, 3), np.float32), 3, None)
  emb = HypergraphEmbedding()
  emb.ParseFromString(wire)
  assert emb.dim == 3 and not emb.HasField("method_name") and len(emb.node) == 0
