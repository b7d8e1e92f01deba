"""Link-prediction evaluation hooks -- the piece that closes the loop from an embedding to the
``ExperimentalResult`` files the reference's ``utilities/`` consume.

Mirrors the hook surface of the reference's ``evaluation_util.py``:

  ``LinkPredictionData``                    :33-35   (same field tuple, by contract)
  ``AddPredictionRecords``                  :38-53
  ``RunLinkPredictionExperiment``           :56-71
  ``LinkPredictionDataToResultProto``       :74-81
  ``RemoveRandomConnections``               :84-124  (same consumption of Python's ``random``)
  ``SampleMissingConnections``              :127-157 (same consumption of Python's ``random``)
  ``CalculateCommunityPredictionMetrics``   :160-204
  ``EXPERIMENT_OPTIONS``                    :586-590

The predictors themselves (one SVC per edge / node, a Keras classifier) are downstream of the
embeddings and out of scope; ``EXPERIMENT_OPTIONS`` is a registry a
caller fills with ``RegisterExperiment`` -- ``RunLinkPredictionExperiment`` looks the predictor up
in it exactly as the reference does.  The per-edge / per-node SVC predictors are provided on top
of scikit-learn (imported on first use) because they need nothing else; the Keras one raises.
"""
import logging
import random
from collections import namedtuple

import numpy as np

from .hypergraph_pb2 import EvaluationMetrics, ExperimentalResult, Hypergraph

log = logging.getLogger()

LinkPredictionData = namedtuple("LinkPredictionData",
                                ("hypergraph", "embedding", "good_links", "bad_links", "removal_prob"))


def _pairs(links):
  return [(int(n), int(e)) for n, e in links]


def AddPredictionRecords(eval_metric, good_links, bad_links, predictions):
  """Appends one record per evaluated link: the good links (label true) in their order, then the
  bad ones; ``prediction`` says whether the predictor kept the link."""
  log.info("Adding link data...")
  kept = set(_pairs(predictions))
  for label, links in ((True, good_links), (False, bad_links)):
    for node, edge in links:
      eval_metric.records.add(node_idx=node, edge_idx=edge, label=label,
                              prediction=(node, edge) in kept)
  return eval_metric


def CalculateCommunityPredictionMetrics(predicted_connections, good_links, bad_links):
  """Precision / recall / F1 / accuracy and the confusion counts of a set of predicted links
  against disjoint positive and negative sets.  Precision, recall and F1 stay unset (proto2
  presence) where the reference leaves them unset: no predictions, no positives, both zero."""
  predicted, pos, neg = set(predicted_connections), set(good_links), set(bad_links)
  assert pos.isdisjoint(neg)
  assert predicted <= (pos | neg)
  assert pos or neg
  hits = len(predicted & pos)
  m = EvaluationMetrics()
  if predicted:
    m.precision = hits / len(predicted)
  if pos:
    m.recall = hits / len(pos)
  if m.precision + m.recall:
    m.f1 = 2 * m.precision * m.recall / (m.precision + m.recall)
  m.num_true_pos = hits
  m.num_false_pos = len(predicted) - hits
  m.num_false_neg = len(pos) - hits
  m.num_true_neg = len(neg - predicted)
  m.accuracy = (m.num_true_pos + m.num_true_neg) / (len(pos) + len(neg))
  return m


def RemoveRandomConnections(original_hypergraph, probability):
  """Copy of the hypergraph with each node-edge connection dropped with `probability`, never a
  node's or an edge's last one; returns (copy, removed (node, edge) pairs in removal order).
  Draws from Python's global ``random`` exactly as the reference: one shuffle of the
  connection list in proto iteration order, then one ``random()`` per connection whose node and
  edge both still have another connection."""
  assert 0 <= probability <= 1
  out = Hypergraph()
  out.CopyFrom(original_hypergraph)
  links = [(n, e) for n, node in original_hypergraph.node.items() for e in node.edges]
  random.shuffle(links)
  node_left = {n: len(node.edges) for n, node in out.node.items()}
  edge_left = {e: len(edge.nodes) for e, edge in out.edge.items()}
  removed = []
  for n, e in links:
    if node_left[n] == 1 or edge_left[e] == 1:
      continue
    if random.random() < probability and probability > 0:
      out.node[n].edges.remove(e)
      out.edge[e].nodes.remove(n)
      node_left[n] -= 1
      edge_left[e] -= 1
      removed.append((n, e))
  return out, removed


def SampleMissingConnections(hypergraph, num_samples):
  """`num_samples` distinct (node, edge) pairs with the node not in the edge -- the negatives of
  the link-prediction task.  Rejection sampling with the reference's draws (``random.choice`` of
  a node, then of an edge, at most 10 x num_samples attempts) and its result order (a set's)."""
  nodes, edges = list(hypergraph.node), list(hypergraph.edge)
  assert num_samples < len(nodes) * len(edges)
  assert nodes and edges
  member = {n: set(node.edges) for n, node in hypergraph.node.items()}
  found = set()
  tries = 10 * num_samples
  while len(found) < num_samples and tries:
    tries -= 1
    n = random.choice(nodes)
    e = random.choice(edges)
    if e not in member[n]:
      found.add((n, e))
  if len(found) < num_samples:
    log.critical("SampleMissingConnections failed to find %i samples", num_samples)
  return list(found)


def RunLinkPredictionExperiment(link_prediction_data, experiment_name):
  """Runs the predictor registered under `experiment_name` on bad + good links and returns the
  ``EvaluationMetrics`` proto (metrics, experiment name, one record per link)."""
  assert experiment_name in EXPERIMENT_OPTIONS
  hypergraph, embedding, good_links, bad_links, _ = link_prediction_data
  log.info("Predicting links on subset graph")
  predicted = EXPERIMENT_OPTIONS[experiment_name](hypergraph, embedding, bad_links + good_links)
  log.info("Evaluating link prediction performance")
  metrics = CalculateCommunityPredictionMetrics(predicted, good_links, bad_links)
  metrics.experiment_name = experiment_name
  log.info("Result:\n%s", metrics)
  AddPredictionRecords(metrics, good_links, bad_links, predicted)
  return metrics


def LinkPredictionDataToResultProto(lp_data):
  """``ExperimentalResult`` holding the (reduced) hypergraph, its embedding and the removal
  probability; the caller appends the metrics of each experiment it runs."""
  log.info("Storing data into Experimental Result proto")
  res = ExperimentalResult()
  res.removal_probability = lp_data.removal_prob
  res.hypergraph.CopyFrom(lp_data.hypergraph)
  res.embedding.CopyFrom(lp_data.embedding)
  return res


def PrepLinkPredictionExperiment(hypergraph, removal_prob, embedding_function):
  """The caller side of the hook (runner.py:366-375 flow, experiment_util): hide connections
  with `removal_prob`, sample as many negatives, embed the reduced hypergraph with
  `embedding_function(hypergraph) -> HypergraphEmbedding`."""
  reduced, good_links = RemoveRandomConnections(hypergraph, removal_prob)
  bad_links = SampleMissingConnections(hypergraph, len(good_links))
  return LinkPredictionData(hypergraph=reduced, embedding=embedding_function(reduced),
                            good_links=good_links, bad_links=bad_links, removal_prob=removal_prob)


# ---- predictors -------------------------------------------------------------------------------


def _personalized_classifier_prediction(hypergraph, embedding, links, per_edge):
  """One RBF SVC (C = 1, gamma = 0.1) per edge over node vectors (or per node over edge vectors),
  trained on the members against twice as many sampled non-members
  (evaluation_util.py:317-350, :389-447); the degenerate classifiers accept everything (no
  non-member) or nothing (no member)."""
  from sklearn.svm import SVC
  from sklearn.utils import shuffle
  assert embedding.dim > 0
  vectors = embedding.node if per_edge else embedding.edge
  owners = hypergraph.edge if per_edge else hypergraph.node
  asked = [(n, e) if per_edge else (e, n) for n, e in links if n in embedding.node and e in embedding.edge]
  models = {}
  for owner in set(o for _, o in asked):
    members = set(owners[owner].nodes if per_edge else owners[owner].edges) if owner in owners else set()
    if not members:
      models[owner] = 0
      continue
    others = vectors.keys() - members
    if not others:
      models[owner] = 1
      continue
    others = random.sample(sorted(others), min(len(others), 2 * len(members)))
    x = [list(vectors[i].values) for i in members] + [list(vectors[i].values) for i in others]
    y = [1] * len(members) + [0] * len(others)
    x, y = shuffle(x, y)
    models[owner] = SVC(C=1, gamma=0.1).fit(x, y)
  kept = []
  for item, owner in asked:
    model = models[owner]
    verdict = model if isinstance(model, int) else model.predict([list(vectors[item].values)])[0]
    if verdict > 0:
      kept.append((item, owner) if per_edge else (owner, item))
  return kept


def PersonalizedEdgeClassifierPrediction(hypergraph, embedding, links, run_in_parallel=False):
  return _personalized_classifier_prediction(hypergraph, embedding, links, per_edge=True)


def PersonalizedNodeClassifierPrediction(hypergraph, embedding, links, run_in_parallel=False):
  return _personalized_classifier_prediction(hypergraph, embedding, links, per_edge=False)


def NodeEdgeEmbeddingPrediction(hypergraph, embedding, potential_links, classifier=None,
                                disable_pbar=False):
  """Keeps the links a binary classifier over [node vector, edge vector] scores above 0.5
  (evaluation_util.py:508-548).  The reference trains a Keras model when `classifier` is None;
  Keras is not part of this build, so a classifier with ``predict`` must be passed."""
  if classifier is None:
    raise NotImplementedError("NodeEdgeEmbeddingPrediction needs a classifier: the Keras model of "
                              "evaluation_util.py:466-505 is out of scope of this build")
  links = [(n, e) for n, e in potential_links if n in hypergraph.node and e in hypergraph.edge]
  if not links:
    return []
  for n, e in links:
    assert n in embedding.node and e in embedding.edge
  x = np.array([np.concatenate((embedding.node[n].values, embedding.edge[e].values)) for n, e in links])
  return [link for link, p in zip(links, classifier.predict(x)) if p > 0.5]


# Each entry is a function taking (hypergraph, embedding, links) -> predicted links
EXPERIMENT_OPTIONS = {
    "LP_EDGE_CLASSIFIERS": PersonalizedEdgeClassifierPrediction,
    "LP_NODE_CLASSIFIERS": PersonalizedNodeClassifierPrediction,
    "LP_NODE_EDGE_CLASSIFIER": NodeEdgeEmbeddingPrediction,
}


def RegisterExperiment(name, predictor):
  """Adds a predictor ``(hypergraph, embedding, links) -> links`` to ``EXPERIMENT_OPTIONS``."""
  EXPERIMENT_OPTIONS[name] = predictor
  return predictor
