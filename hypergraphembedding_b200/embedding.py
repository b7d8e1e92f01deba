"""Drop-in for the hypergraph2vec part of the reference's ``hypergraph_embedding/embedding.py``:
``_hypergraph2vec_skeleton``:269, ``EmbedHg2vBoolean``:308 (FOBE), ``EmbedHg2vAdjJaccard``:330,
``EmbedHg2vNeighborhoodWeightedJaccard``:358, ``EmbedHg2vAlgDist``:387 (HOBE) and the matching
entries of ``EMBEDDING_OPTIONS``:417 -- what ``runner.py --embedding-method HG2V_*`` calls.

Same arguments, defaults, consumption of the global numpy RNG (the sampler's draws, then one
seed draw per embedding table as Keras' initializers make, then one shuffle per epoch) and result (a ``HypergraphEmbedding`` keyed by the original
ids with the reference's ``method_name``).  Sampling and training run in libhge_b200.so.
"""
import logging

from .algebraic_distance import EmbedAlgebraicDistance
from .hg2v_model import (BooleanModel, EarlyStopping, KerasModelToEmbedding, UnweightedFloatModel)
from .hg2v_sample import (AlgebraicDistanceSamples, BooleanSamples, PlotDistributions,
                          SamplesToModelInput, WeightedJaccardSamples)
from .hg2v_weighting import UniformWeight, WeightByNeighborhood
from .hypergraph_util import CompressRange

log = logging.getLogger()


def _hypergraph2vec_skeleton(hypergraph, dimension, num_neighbors, sampler_fn, model_fn,
                             fit_batch_size, fit_epochs, debug_summary_path, disable_pbar):
  """embedding.py:269-305."""
  log.info("Compressing index space")
  # we want to do this in order to reduce the embedding problem size
  compressed_hg, inv_node_map, inv_edge_map = CompressRange(hypergraph)

  log.info("Sampling")
  samples = sampler_fn(compressed_hg)

  if debug_summary_path is not None:
    PlotDistributions(debug_summary_path, samples)

  log.info("Converting samples to model input")
  input_features, output_probs = SamplesToModelInput(samples, num_neighbors=num_neighbors,
                                                     weighted=False)

  log.info("Getting model")
  model = model_fn(compressed_hg)
  stopper = EarlyStopping(monitor="loss", min_delta=1e-3)
  try:
    model.fit(input_features, output_probs, batch_size=fit_batch_size, epochs=fit_epochs,
              callbacks=[stopper], verbose=0 if disable_pbar else 1)
    log.info("Recording embeddings.")
    return KerasModelToEmbedding(compressed_hg, model, inv_node_map, inv_edge_map)
  finally:
    model.close()


def EmbedHg2vBoolean(hypergraph, dimension, num_neighbors=5, num_samples=200, batch_size=256,
                     epochs=10, neg_samples=0, debug_summary_path=None, disable_pbar=False):
  """embedding.py:308-327 (FOBE)."""
  sampler_fn = lambda hg: BooleanSamples(hg, num_neighbors=num_neighbors, num_samples=num_samples,
                                         neg_samples=neg_samples, disable_pbar=disable_pbar)
  model_fn = lambda hg: BooleanModel(hg, dimension=dimension, num_neighbors=num_neighbors)
  embedding = _hypergraph2vec_skeleton(hypergraph, dimension, num_neighbors, sampler_fn, model_fn,
                                       batch_size, epochs, debug_summary_path, disable_pbar)
  embedding.method_name = "HG2V_BOOLEAN"
  return embedding


def EmbedHg2vAdjJaccard(hypergraph, dimension, num_neighbors=5, num_samples=200, batch_size=256,
                        epochs=10, debug_summary_path=None, disable_pbar=False):
  """embedding.py:330-355."""

  def sampler_fn(hypergraph):
    node2weight, edge2weight = UniformWeight(hypergraph)
    return WeightedJaccardSamples(hypergraph, node2weight, edge2weight, num_neighbors=num_neighbors,
                                  num_samples=num_samples, disable_pbar=disable_pbar)

  model_fn = lambda hg: UnweightedFloatModel(hg, dimension=dimension, num_neighbors=num_neighbors)
  embedding = _hypergraph2vec_skeleton(hypergraph, dimension, num_neighbors, sampler_fn, model_fn,
                                       batch_size, epochs, debug_summary_path, disable_pbar)
  embedding.method_name = "HG2V_ADJ_JAC"
  return embedding


def EmbedHg2vNeighborhoodWeightedJaccard(hypergraph, dimension, alpha=0, num_neighbors=5,
                                         num_samples=200, batch_size=256, epochs=10,
                                         debug_summary_path=None, disable_pbar=False):
  """embedding.py:358-384."""

  def sampler_fn(hypergraph):
    node2feature, edge2feature = WeightByNeighborhood(hypergraph, alpha)
    return WeightedJaccardSamples(hypergraph, node2feature, edge2feature,
                                  num_neighbors=num_neighbors, num_samples=num_samples,
                                  disable_pbar=disable_pbar)

  model_fn = lambda hg: UnweightedFloatModel(hg, dimension=dimension, num_neighbors=num_neighbors)
  embedding = _hypergraph2vec_skeleton(hypergraph, dimension, num_neighbors, sampler_fn, model_fn,
                                       batch_size, epochs, debug_summary_path, disable_pbar)
  embedding.method_name = "HG2V_NEIGH_JAC"
  return embedding


def EmbedHg2vAlgDist(hypergraph, dimension, alpha=0, num_neighbors=5, num_samples=200,
                     batch_size=256, epochs=10, debug_summary_path=None, disable_pbar=False):
  """embedding.py:387-414 (HOBE): algebraic-distance embedding (dimension 10, 20 sweeps), samples
  weighted by it, UnweightedFloatModel trained on them."""
  del alpha   # accepted and unused, as in the reference

  def sampler_fn(hypergraph):
    log.info("Embedding weighted by algebraic distance.")
    alg_emb = EmbedAlgebraicDistance(hypergraph, dimension=10, iterations=20,
                                     disable_pbar=disable_pbar)
    return AlgebraicDistanceSamples(hypergraph, alg_emb, num_neighbors=num_neighbors,
                                    num_samples=num_samples, disable_pbar=disable_pbar)

  model_fn = lambda hg: UnweightedFloatModel(hg, dimension=dimension, num_neighbors=num_neighbors)
  embedding = _hypergraph2vec_skeleton(hypergraph, dimension, num_neighbors, sampler_fn, model_fn,
                                       batch_size, epochs, debug_summary_path, disable_pbar)
  embedding.method_name = "HG2V_ALG_DIST"
  return embedding


# the part of embedding.py:417-436 that is on this path
EMBEDDING_OPTIONS = {
    "ALG_DIST": EmbedAlgebraicDistance,
    "HG2V_BOOLEAN": EmbedHg2vBoolean,
    "HG2V_ADJ_JAC": EmbedHg2vAdjJaccard,
    "HG2V_NEIGH_JAC": EmbedHg2vNeighborhoodWeightedJaccard,
    "HG2V_ALG_DIST": EmbedHg2vAlgDist,
}
