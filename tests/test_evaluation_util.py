"""Link-prediction evaluation hooks (hypergraphembedding_b200/evaluation_util.py) against
(i) the golden vectors written from the unmodified reference by oracle/make_golden_eval.py
(seeded ``random``: same hidden connections, same negatives, same RNG state afterwards, the same
EvaluationMetrics / ExperimentalResult bytes) and (ii) the reference's own known-answer tests
(tests/test_evaluation_util.py:51-101, :106-183, :187-201, :440-473), restated."""
import hashlib
import random

import numpy as np
import pytest

from conftest import load_golden


def _hypergraph(pairs):
  from hypergraphembedding_b200 import AddNodeToEdge, Hypergraph
  hg = Hypergraph()
  for n, e in np.asarray(pairs).tolist():
    AddNodeToEdge(hg, n, e)
  return hg


def _links(hg):
  return {(n, e) for n, node in hg.node.items() for e in node.edges}


def _mod3(hypergraph, embedding, links):
  return [(n, e) for n, e in links if (n + e) % 3 == 0]


def test_hooks_match_the_reference_golden():
  from hypergraphembedding_b200 import HypergraphEmbedding, evaluation_util as ev
  g = load_golden("eval_hooks")
  hg = _hypergraph(g["pairs"])
  random.seed(int(g["seed"]))
  reduced, removed = ev.RemoveRandomConnections(hg, float(g["removal"]))
  missing = ev.SampleMissingConnections(hg, len(removed))
  assert removed == [tuple(p) for p in g["removed"].tolist()]
  assert missing == [tuple(p) for p in g["missing"].tolist()]
  assert hashlib.sha256(repr(random.getstate()).encode()).hexdigest() == str(g["random_state_sha"])
  assert [(n, e) for n, node in reduced.node.items() for e in node.edges] == \
      [tuple(p) for p in g["reduced_pairs"].tolist()]
  ev.RegisterExperiment("GOLDEN_MOD3", _mod3)
  data = ev.LinkPredictionData(hypergraph=reduced, embedding=HypergraphEmbedding(), good_links=removed,
                               bad_links=missing, removal_prob=float(g["removal"]))
  metrics = ev.RunLinkPredictionExperiment(data, "GOLDEN_MOD3")
  assert metrics.SerializeToString(deterministic=True) == g["metrics"].tobytes()
  result = ev.LinkPredictionDataToResultProto(data)
  assert result.SerializeToString(deterministic=True) == g["result"].tobytes()


def test_remove_random_connections_keeps_every_node_and_edge():
  from hypergraphembedding_b200 import evaluation_util as ev
  hg = _hypergraph([(0, 0), (0, 1), (1, 1)])
  hg.name = "KEEP_ME"
  out, removed = ev.RemoveRandomConnections(hg, 1)
  assert set(out.node) == {0, 1} and set(out.edge) == {0, 1}
  assert out != hg and out.name == "KEEP_ME"
  assert _links(out) | set(removed) == _links(hg)
  same, none = ev.RemoveRandomConnections(hg, 0)
  assert none == [] and same == hg


def test_remove_random_connections_fuzz():
  from hypergraphembedding_b200 import evaluation_util as ev
  rng = np.random.RandomState(3)
  for _ in range(10):
    pairs = [(n, e) for n in range(rng.randint(1, 11)) for e in range(rng.randint(1, 11))
             if rng.rand() < rng.rand()]
    if not pairs:
      continue
    hg = _hypergraph(pairs)
    out, removed = ev.RemoveRandomConnections(hg, rng.rand())
    assert _links(out).isdisjoint(removed) and _links(out) | set(removed) == _links(hg)
    assert _links(out) == {(n, e) for e, edge in out.edge.items() for n in edge.nodes}
    assert set(out.node) == set(hg.node) and set(out.edge) == set(hg.edge)


def test_community_prediction_metrics_known_answers():
  from hypergraphembedding_b200.evaluation_util import CalculateCommunityPredictionMetrics as metrics
  m = metrics([(1, 2), (2, 1), (2, 3), (2, 4)], [(1, 2), (2, 4), (2, 5)],
              [(3, 0), (3, 2), (2, 1), (2, 3)])
  assert m.accuracy == pytest.approx(4 / 7, abs=1e-6)
  assert m.precision == pytest.approx(2 / 4, abs=1e-6)
  assert m.recall == pytest.approx(2 / 3, abs=1e-6)
  assert m.f1 == pytest.approx(2 * (2 / 4) * (2 / 3) / (2 / 4 + 2 / 3), abs=1e-6)
  assert (m.num_true_pos, m.num_false_pos, m.num_false_neg, m.num_true_neg) == (2, 2, 1, 2)
  m = metrics([], [(1, 2)], [(2, 3)])
  assert m.accuracy == 0.5 and not m.HasField("precision") and m.recall == 0 and not m.HasField("f1")
  assert (m.num_true_pos, m.num_false_pos, m.num_false_neg, m.num_true_neg) == (0, 0, 1, 1)
  m = metrics([(1, 2)], [], [(1, 2)])
  assert m.accuracy == 0 and not m.HasField("recall") and m.precision == 0 and not m.HasField("f1")
  assert (m.num_true_pos, m.num_false_pos, m.num_false_neg, m.num_true_neg) == (0, 1, 0, 0)


def test_sample_missing_connections_fuzz():
  from hypergraphembedding_b200 import evaluation_util as ev
  rng = np.random.RandomState(5)
  for _ in range(10):
    hg = _hypergraph([(n, e) for n in range(50) for e in range(50) if rng.rand() < 0.1])
    want = int(rng.randint(0, 11))
    got = ev.SampleMissingConnections(hg, want)
    assert len(got) == want == len(set(got))
    for n, e in got:
      assert n in hg.node and e in hg.edge and e not in hg.node[n].edges


def test_add_prediction_records_known_answer():
  from hypergraphembedding_b200 import EvaluationMetrics, evaluation_util as ev
  m = ev.AddPredictionRecords(EvaluationMetrics(), [(0, 0), (0, 1)], [(1, 0), (1, 1)], [(0, 0), (1, 1)])
  assert [(r.node_idx, r.edge_idx, r.label, r.prediction) for r in m.records] == \
      [(0, 0, True, True), (0, 1, True, False), (1, 0, False, False), (1, 1, False, True)]


def test_personalized_classifiers_and_prep():
  from hypergraphembedding_b200 import HypergraphEmbedding, evaluation_util as ev
  rng = np.random.RandomState(11)
  hg = _hypergraph([(n, e) for n in range(24) for e in range(4) if (n // 6 == e or rng.rand() < 0.05)])

  def embed(h):
    emb = HypergraphEmbedding()
    emb.dim = 2
    for n in h.node:
      emb.node[n].values.extend([float(n // 6), rng.rand() * 0.1])
    for e in h.edge:
      emb.edge[e].values.extend([float(e), 0.05])
    return emb

  random.seed(2)
  data = ev.PrepLinkPredictionExperiment(hg, 0.2, embed)
  assert len(data.good_links) == len(data.bad_links) > 0
  for name in ("LP_EDGE_CLASSIFIERS", "LP_NODE_CLASSIFIERS"):
    m = ev.RunLinkPredictionExperiment(data, name)
    assert m.experiment_name == name and len(m.records) == 2 * len(data.good_links)
    assert m.num_true_pos + m.num_false_neg == len(data.good_links)
  with pytest.raises(NotImplementedError):
    ev.RunLinkPredictionExperiment(data, "LP_NODE_EDGE_CLASSIFIER")

  class Half:
    def predict(self, x):
      return (x[:, 0] == x[:, 2]).astype(np.float32)

  kept = ev.NodeEdgeEmbeddingPrediction(data.hypergraph, data.embedding, data.good_links + [(999, 0)],
                                        classifier=Half())
  assert set(kept) <= set(data.good_links)
