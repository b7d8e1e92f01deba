"""TEST INFRASTRUCTURE ONLY -- import shim that makes the *unmodified* reference
hot-path modules importable in the dev container.

Only ``tests/``, ``oracle/make_golden.py`` and ad-hoc validation scripts use this.
Nothing under ``hypergraphembedding_b200/`` may import it, and nothing that runs on
the GPU box may need it: ``/root/reference`` does not exist there.  The golden
vectors it produces are committed under ``tests/golden/``.

What is shimmed:
  * ``hypergraph_embedding.hypergraph_pb2`` is git-ignored upstream and ``protoc``
    is absent here, so the message classes are built at runtime from a
    ``FileDescriptorProto`` (same schema as ``hypergraph.proto:6-35``).
  * ``hypergraph_embedding/__init__.py:8-16`` pulls in keras / node2vec; a bare
    package object with ``__path__`` pointing at the reference tree is registered
    instead so sub-modules import without executing ``__init__``.
  * ``matplotlib`` (imported at module scope by ``hg2v_sample.py:21-23`` and
    ``hg2v_weighting.py:25-27``) is replaced by a no-op stub.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("HGE_REFERENCE_ROOT", "/root/reference")


def reference_available():
  return os.path.isdir(os.path.join(REFERENCE_ROOT, "hypergraph_embedding"))


def _proto_classes():
  # Private pool, same schema: the reference only needs attribute-compatible
  # message classes.  Re-using the product's builder keeps a single schema.
  here = os.path.dirname(os.path.abspath(__file__))
  path = os.path.join(here, "..", "hypergraphembedding_b200", "hypergraph_pb2.py")
  spec = importlib.util.spec_from_file_location("_hge_ref_shim_pb2", path)
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  return mod


_loaded = None


def load_reference():
  """Returns a namespace with the reference modules
  (hypergraph_util, algebraic_distance, hg2v_sample, hg2v_weighting) and the
  proto classes they use."""
  global _loaded
  if _loaded is not None:
    return _loaded
  if not reference_available():
    raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
  sys.dont_write_bytecode = True  # the reference tree is read-only

  pb = _proto_classes()
  pkg = types.ModuleType("hypergraph_embedding")
  pkg.__path__ = [os.path.join(REFERENCE_ROOT, "hypergraph_embedding")]
  pkg.Hypergraph = pb.Hypergraph
  pkg.HypergraphEmbedding = pb.HypergraphEmbedding
  pkg.EvaluationMetrics = pb.EvaluationMetrics
  pkg.ExperimentalResult = pb.ExperimentalResult
  pb2 = types.ModuleType("hypergraph_embedding.hypergraph_pb2")
  pb2.Hypergraph = pb.Hypergraph
  pb2.HypergraphEmbedding = pb.HypergraphEmbedding
  pb2.EvaluationMetrics = pb.EvaluationMetrics
  pb2.ExperimentalResult = pb.ExperimentalResult
  sys.modules["hypergraph_embedding"] = pkg
  sys.modules["hypergraph_embedding.hypergraph_pb2"] = pb2

  if "matplotlib" not in sys.modules:
    try:
      import matplotlib  # noqa: F401
    except ImportError:
      mpl = types.ModuleType("matplotlib")
      mpl.use = lambda *a, **k: None
      plt = types.ModuleType("matplotlib.pyplot")
      mpl.pyplot = plt
      sys.modules["matplotlib"] = mpl
      sys.modules["matplotlib.pyplot"] = plt

  ns = types.SimpleNamespace()
  ns.Hypergraph = pb.Hypergraph
  ns.HypergraphEmbedding = pb.HypergraphEmbedding
  for name in ("hypergraph_util", "algebraic_distance", "hg2v_sample",
               "hg2v_weighting"):
    setattr(ns, name, importlib.import_module("hypergraph_embedding." + name))
  _loaded = ns
  return ns


def load_reference_evaluation():
  """The reference's ``evaluation_util`` module (evaluation_util.py:1-590).  It imports keras at
  module scope for one predictor that nothing here calls; keras is absent, so attribute-free
  stub modules stand in for it."""
  load_reference()
  for name in ("keras", "keras.models", "keras.layers", "keras.callbacks"):
    if name not in sys.modules:
      try:
        importlib.import_module(name)
      except ImportError:
        stub = types.ModuleType(name)
        for attr in ("Model", "Input", "Dense", "EarlyStopping"):
          setattr(stub, attr, None)
        sys.modules[name] = stub
  return importlib.import_module("hypergraph_embedding.evaluation_util")
