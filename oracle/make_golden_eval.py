"""TEST INFRASTRUCTURE ONLY -- golden vectors of the link-prediction hooks
(evaluation_util.py:56-204), written by running the UNMODIFIED reference through
``oracle.ref_shim`` on a seeded random hypergraph with Python's ``random`` seeded:

    python -m oracle.make_golden_eval

Writes tests/golden/eval_hooks.npz: the hypergraph (pairs in proto insertion order), the
connections RemoveRandomConnections hid, the negatives SampleMissingConnections drew, the state of
``random`` afterwards (hashed), and the serialized EvaluationMetrics of RunLinkPredictionExperiment
with a deterministic predictor.
"""
import hashlib
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "..", "tests", "golden")
sys.path.insert(0, os.path.join(HERE, ".."))

from oracle import ref_shim  # noqa: E402

SEED, NODES, EDGES, PROB, REMOVAL = 7, 60, 40, 0.12, 0.35


def build(hg_cls, add):
  rng = np.random.RandomState(SEED)
  hg = hg_cls()
  pairs = []
  for n in range(NODES):
    for e in range(EDGES):
      if rng.rand() < PROB:
        add(hg, n, e)
        pairs.append((n, e))
  return hg, np.asarray(pairs, np.int64)


def predictor(hypergraph, embedding, links):
  # deterministic stand-in for a classifier: keeps links whose ids sum to a multiple of 3
  return [(n, e) for n, e in links if (n + e) % 3 == 0]


def main():
  ref = ref_shim.load_reference()
  ev = ref_shim.load_reference_evaluation()
  hg, pairs = build(ref.Hypergraph, ref.hypergraph_util.AddNodeToEdge)
  random.seed(SEED)
  reduced, removed = ev.RemoveRandomConnections(hg, REMOVAL)
  missing = ev.SampleMissingConnections(hg, len(removed))
  state = hashlib.sha256(repr(random.getstate()).encode()).hexdigest()
  ev.EXPERIMENT_OPTIONS["GOLDEN_MOD3"] = predictor
  data = ev.LinkPredictionData(hypergraph=reduced, embedding=ref.HypergraphEmbedding(),
                               good_links=removed, bad_links=missing, removal_prob=REMOVAL)
  metrics = ev.RunLinkPredictionExperiment(data, "GOLDEN_MOD3")
  result = ev.LinkPredictionDataToResultProto(data)
  reduced_pairs = np.asarray([(n, e) for n, node in reduced.node.items() for e in node.edges], np.int64)
  np.savez(os.path.join(GOLDEN, "eval_hooks.npz"), seed=SEED, removal=REMOVAL, pairs=pairs,
           removed=np.asarray(removed, np.int64), missing=np.asarray(missing, np.int64),
           reduced_pairs=reduced_pairs, random_state_sha=state,
           metrics=np.frombuffer(metrics.SerializeToString(deterministic=True), np.uint8),
           result=np.frombuffer(result.SerializeToString(deterministic=True), np.uint8))
  print("removed %d of %d, %d negatives, metrics %d bytes" % (len(removed), len(pairs), len(missing),
                                                               metrics.ByteSize()))


if __name__ == "__main__":
  main()
