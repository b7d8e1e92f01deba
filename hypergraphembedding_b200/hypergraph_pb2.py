"""Runtime-built protobuf classes for the reference's on-disk format.

The reference generates ``hypergraph_pb2.py`` with ``protoc`` from
``hypergraph_embedding/hypergraph.proto:1-69`` (``Makefile:10-11``).  There is no
``protoc`` in this image, so the identical message classes are built here from a
hand-written ``FileDescriptorProto``.  Field numbers, types, defaults and the
package name follow the reference schema exactly, so serialized bytes are
interchangeable with files written by the reference (``runner.py:347-364``).

Exports: ``Hypergraph``, ``HypergraphEmbedding``, ``EvaluationMetrics``,
``ExperimentalResult``.
"""
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

_F = descriptor_pb2.FieldDescriptorProto

_PACKAGE = "hypergraph_embedding"
# A private file name keeps this pool entry from colliding with a protoc-generated
# hypergraph_pb2 that a host application may also have imported.
_FILE_NAME = "hypergraphembedding_b200/hypergraph.proto"


def _field(msg, name, number, ftype, label=_F.LABEL_OPTIONAL, type_name=None,
           default=None):
  f = msg.field.add()
  f.name = name
  f.number = number
  f.type = ftype
  f.label = label
  if type_name is not None:
    f.type_name = type_name
  if default is not None:
    f.default_value = default
  return f


def _map_entry(parent, entry_name, value_type_name):
  entry = parent.nested_type.add()
  entry.name = entry_name
  entry.options.map_entry = True
  _field(entry, "key", 1, _F.TYPE_INT32)
  _field(entry, "value", 2, _F.TYPE_MESSAGE, type_name=value_type_name)
  return entry


def _build_file():
  fd = descriptor_pb2.FileDescriptorProto()
  fd.name = _FILE_NAME
  fd.package = _PACKAGE
  fd.syntax = "proto2"

  # message Hypergraph (hypergraph.proto:6-23)
  hg = fd.message_type.add()
  hg.name = "Hypergraph"
  node_data = hg.nested_type.add()
  node_data.name = "NodeData"
  _field(node_data, "edges", 1, _F.TYPE_INT32, _F.LABEL_REPEATED)
  _field(node_data, "name", 2, _F.TYPE_STRING)
  _field(node_data, "weight", 3, _F.TYPE_FLOAT, default="1")
  edge_data = hg.nested_type.add()
  edge_data.name = "EdgeData"
  _field(edge_data, "nodes", 1, _F.TYPE_INT32, _F.LABEL_REPEATED)
  _field(edge_data, "name", 2, _F.TYPE_STRING)
  _field(edge_data, "weight", 3, _F.TYPE_FLOAT, default="1")
  _map_entry(hg, "NodeEntry", ".%s.Hypergraph.NodeData" % _PACKAGE)
  _map_entry(hg, "EdgeEntry", ".%s.Hypergraph.EdgeData" % _PACKAGE)
  _field(hg, "node", 1, _F.TYPE_MESSAGE, _F.LABEL_REPEATED,
         ".%s.Hypergraph.NodeEntry" % _PACKAGE)
  _field(hg, "edge", 2, _F.TYPE_MESSAGE, _F.LABEL_REPEATED,
         ".%s.Hypergraph.EdgeEntry" % _PACKAGE)
  _field(hg, "name", 3, _F.TYPE_STRING)

  # message HypergraphEmbedding (hypergraph.proto:26-35)
  emb = fd.message_type.add()
  emb.name = "HypergraphEmbedding"
  vec = emb.nested_type.add()
  vec.name = "Embedding"
  _field(vec, "values", 1, _F.TYPE_FLOAT, _F.LABEL_REPEATED)
  _map_entry(emb, "NodeEntry", ".%s.HypergraphEmbedding.Embedding" % _PACKAGE)
  _map_entry(emb, "EdgeEntry", ".%s.HypergraphEmbedding.Embedding" % _PACKAGE)
  _field(emb, "node", 1, _F.TYPE_MESSAGE, _F.LABEL_REPEATED,
         ".%s.HypergraphEmbedding.NodeEntry" % _PACKAGE)
  _field(emb, "edge", 2, _F.TYPE_MESSAGE, _F.LABEL_REPEATED,
         ".%s.HypergraphEmbedding.EdgeEntry" % _PACKAGE)
  _field(emb, "dim", 3, _F.TYPE_INT32)
  _field(emb, "method_name", 4, _F.TYPE_STRING)

  # message EvaluationMetrics (hypergraph.proto:37-59)
  met = fd.message_type.add()
  met.name = "EvaluationMetrics"
  for i, n in enumerate(("accuracy", "precision", "recall", "f1"), start=1):
    _field(met, n, i, _F.TYPE_FLOAT)
  for i, n in enumerate(("num_true_pos", "num_true_neg", "num_false_pos",
                         "num_false_neg"), start=5):
    _field(met, n, i, _F.TYPE_INT32)
  _field(met, "experiment_name", 9, _F.TYPE_STRING)
  rec = met.nested_type.add()
  rec.name = "EvaluationRecord"
  _field(rec, "node_idx", 1, _F.TYPE_INT32)
  _field(rec, "edge_idx", 2, _F.TYPE_INT32)
  _field(rec, "label", 3, _F.TYPE_BOOL)
  _field(rec, "prediction", 4, _F.TYPE_BOOL)
  _field(met, "records", 10, _F.TYPE_MESSAGE, _F.LABEL_REPEATED,
         ".%s.EvaluationMetrics.EvaluationRecord" % _PACKAGE)

  # message ExperimentalResult (hypergraph.proto:61-69)
  res = fd.message_type.add()
  res.name = "ExperimentalResult"
  _field(res, "hypergraph", 1, _F.TYPE_MESSAGE,
         type_name=".%s.Hypergraph" % _PACKAGE)
  _field(res, "embedding", 2, _F.TYPE_MESSAGE,
         type_name=".%s.HypergraphEmbedding" % _PACKAGE)
  _field(res, "metrics", 3, _F.TYPE_MESSAGE, _F.LABEL_REPEATED,
         ".%s.EvaluationMetrics" % _PACKAGE)
  _field(res, "removal_probability", 4, _F.TYPE_FLOAT)
  return fd


_pool = descriptor_pool.DescriptorPool()
_pool.Add(_build_file())


def _cls(name):
  return message_factory.GetMessageClass(
      _pool.FindMessageTypeByName("%s.%s" % (_PACKAGE, name)))


Hypergraph = _cls("Hypergraph")
HypergraphEmbedding = _cls("HypergraphEmbedding")
EvaluationMetrics = _cls("EvaluationMetrics")
ExperimentalResult = _cls("ExperimentalResult")

__all__ = [
    "Hypergraph", "HypergraphEmbedding", "EvaluationMetrics",
    "ExperimentalResult"
]
