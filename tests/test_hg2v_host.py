"""Host-side logic around the hypergraph2vec trainer that needs no GPU: the float64 restatement
against a finite-difference gradient (so the oracle the kernels are checked against is itself
right), EarlyStopping semantics, KerasModelToEmbedding, and the wiring of the EmbedHg2v*
skeleton (embedding.py:269-305) with stand-in sampler / model objects."""
import numpy as np
import pytest

from oracle import hg2v_model_ref as ref


def _problem(seed, nodes=12, edges=7, k=2, m=40):
  rng = np.random.default_rng(seed)
  feats = [rng.integers(0, nodes + 1, m), rng.integers(0, edges + 1, m), rng.integers(0, nodes + 1, m),
           rng.integers(0, edges + 1, m)]
  feats += [rng.integers(0, nodes + 1, m) for _ in range(k)]
  feats += [rng.integers(0, edges + 1, m) for _ in range(k)]
  targets = [rng.random(m), rng.random(m), rng.random(m)]
  N = rng.uniform(-0.6, 0.6, (nodes + 1, 5))
  E = rng.uniform(-0.6, 0.6, (edges + 1, 5))
  return N, E, feats, targets


def _loss(N, E, feats, targets, k, activation, loss):
  ln, le, rn, re = feats[:4]
  p_nn = ref.act((N[ln] * N[rn]).sum(1), activation)
  p_ee = ref.act((E[le] * E[re]).sum(1), activation)
  a = ref.act(np.einsum("mkd,md->mk", N[np.stack(feats[4:4 + k], 1)], N[ln]), activation).mean(1)
  b = ref.act(np.einsum("mkd,md->mk", E[np.stack(feats[4 + k:], 1)], E[re]), activation).mean(1)
  return sum(ref.loss_and_grad(p, y, loss)[0].mean() for p, y in zip((p_nn, p_ee, a * b), targets))


@pytest.mark.parametrize("activation,loss", [("sigmoid", "kld"), ("sigmoid", "mse"), ("relu", "mse")])
def test_restatement_gradient_against_finite_differences(activation, loss):
  """One Adagrad step from zero accumulators moves every touched parameter by
  lr * g / (|g| + eps); recover g's sign and support from it and compare the loss decrease with
  a numerical directional derivative."""
  k = 2
  N, E, feats, targets = _problem(1, k=k)
  N1, E1 = N.copy(), E.copy()
  accN, accE = np.zeros_like(N), np.zeros_like(E)
  ref.batch_step(N1, E1, accN, accE, feats, targets, k, activation, loss)
  gN, gE = np.sqrt(accN) * np.sign(N - N1), np.sqrt(accE) * np.sign(E - E1)   # acc = g^2 after one step
  h = 1e-6
  rng = np.random.default_rng(2)
  dN, dE = rng.standard_normal(N.shape), rng.standard_normal(E.shape)
  numeric = (_loss(N + h * dN, E + h * dE, feats, targets, k, activation, loss) -
             _loss(N - h * dN, E - h * dE, feats, targets, k, activation, loss)) / (2 * h)
  analytic = (gN * dN).sum() + (gE * dE).sum()
  assert abs(numeric - analytic) <= 1e-5 * max(1.0, abs(analytic)), (numeric, analytic)


def test_early_stopping_is_keras_patience_zero():
  from hypergraphembedding_b200.hg2v_model import EarlyStopping
  stop = EarlyStopping(monitor="loss", min_delta=1e-3)
  stop.on_train_begin()
  assert not stop.on_epoch_end(0, 1.0)            # first epoch always improves on +inf
  assert not stop.on_epoch_end(1, 0.9)
  assert stop.on_epoch_end(2, 0.8995)             # improved by less than min_delta -> stop
  assert stop.stopped_epoch == 2
  patient = EarlyStopping(min_delta=0, patience=2)
  patient.on_train_begin()
  assert [patient.on_epoch_end(i, v) for i, v in enumerate([1.0, 1.0, 1.0])] == [False, False, True]


def test_model_to_embedding_skips_the_padding_row():
  from hypergraphembedding_b200 import Hypergraph
  from hypergraphembedding_b200.hg2v_model import KerasModelToEmbedding

  class Layer(object):
    def __init__(self, w):
      self.w = w

    def get_weights(self):
      return [self.w]

  class Model(object):
    layers = {"node_embedding": Layer(np.arange(12, dtype=np.float32).reshape(4, 3)),
              "edge_embedding": Layer(-np.arange(9, dtype=np.float32).reshape(3, 3))}

    def get_layer(self, name):
      return self.layers[name]

  hg = Hypergraph()
  for n, e in [(0, 0), (1, 0), (2, 1)]:
    hg.node[n].edges.append(e)
    hg.edge[e].nodes.append(n)
  emb = KerasModelToEmbedding(hg, Model(), {0: 70, 1: 5, 2: 31}, {0: 9, 1: 2})
  assert emb.dim == 3 and sorted(emb.node) == [5, 31, 70] and sorted(emb.edge) == [2, 9]
  assert list(emb.node[70].values) == [3, 4, 5]          # compressed node 0 -> table row 1
  assert list(emb.node[31].values) == [9, 10, 11]
  assert list(emb.edge[2].values) == [-6, -7, -8]


def test_skeleton_wiring(monkeypatch):
  """embedding.py:269-305: compress, sample on the compressed hypergraph, pack, fit with the
  reference's settings, map the tables back to the original ids."""
  from hypergraphembedding_b200 import Hypergraph, SimilarityRecord, embedding

  calls = {}

  class Model(object):
    def __init__(self, hg):
      calls["model_nodes"] = sorted(hg.node)

    def fit(self, x, y, batch_size, epochs, callbacks, verbose):
      calls["fit"] = (len(x), len(y), batch_size, epochs, callbacks[0].min_delta, verbose)

    def get_layer(self, name):
      table = np.ones((4, 2), np.float32) * (1 if name == "node_embedding" else 2)

      class L(object):
        def get_weights(self):
          return [table]
      return L()

    def close(self):
      calls["closed"] = True

  hg = Hypergraph()
  for n, e in [(10, 5), (20, 5), (30, 8)]:
    hg.node[n].edges.append(e)
    hg.edge[e].nodes.append(n)

  def sampler(compressed):
    calls["sampler_nodes"] = sorted(compressed.node)
    return [SimilarityRecord(left_node_idx=0, right_node_idx=1, node_node_prob=1)]

  emb = embedding._hypergraph2vec_skeleton(hg, 2, 3, sampler, Model, 256, 10, None, True)
  assert calls["sampler_nodes"] == [0, 1, 2] and calls["model_nodes"] == [0, 1, 2]
  assert calls["fit"] == (4 + 2 * 3, 3, 256, 10, 1e-3, 0) and calls["closed"]
  assert sorted(emb.node) == [10, 20, 30] and sorted(emb.edge) == [5, 8]
  assert list(emb.node[10].values) == [1, 1] and list(emb.edge[8].values) == [2, 2]
  assert sorted(embedding.EMBEDDING_OPTIONS) == ["ALG_DIST", "HG2V_ADJ_JAC", "HG2V_ALG_DIST",
                                                 "HG2V_BOOLEAN", "HG2V_NEIGH_JAC"]


@pytest.mark.parametrize("activation,loss", [("sigmoid", "kld"), ("relu", "mse")])
def test_restatement_equals_torch_autograd_with_torch_adagrad(activation, loss):
  """An independent derivation of the same training step: the model written with torch ops,
  gradients from autograd, torch.optim.Adagrad(lr=0.01, eps=1e-7, initial accumulator 0) -- the
  update rule Keras' Adagrad has -- over several batches with duplicated rows."""
  import torch
  k = 2
  N0, E0, feats, targets = _problem(7, nodes=15, edges=9, k=k, m=96)
  N, E = N0.copy(), E0.copy()
  accN, accE = np.zeros_like(N), np.zeros_like(E)
  tN = torch.tensor(N0, dtype=torch.float64, requires_grad=True)
  tE = torch.tensor(E0, dtype=torch.float64, requires_grad=True)
  opt = torch.optim.Adagrad([tN, tE], lr=0.01, eps=1e-7, initial_accumulator_value=0.0)
  act = torch.sigmoid if activation == "sigmoid" else torch.relu

  def loss_fn(p, y):
    if loss == "kld":
      yc, pc = y.clamp(1e-7, 1.0), p.clamp(1e-7, 1.0)
      return (yc * torch.log(yc / pc)).mean()
    return ((p - y)**2).mean()

  for lo in range(0, 96, 32):
    sel = slice(lo, lo + 32)
    f = [np.asarray(c[sel]) for c in feats]
    t = [np.asarray(c[sel]) for c in targets]
    want_loss = ref.batch_step(N, E, accN, accE, f, t, k, activation, loss)
    ft = [torch.as_tensor(c, dtype=torch.long) for c in f]
    tt = [torch.as_tensor(c, dtype=torch.float64) for c in t]
    Ln, Le, Rn, Re = tN[ft[0]], tE[ft[1]], tN[ft[2]], tE[ft[3]]
    p_nn, p_ee = act((Ln * Rn).sum(1)), act((Le * Re).sum(1))
    a = torch.stack([act((tN[ft[4 + i]] * Ln).sum(1)) for i in range(k)], 1).mean(1)
    b = torch.stack([act((tE[ft[4 + k + i]] * Re).sum(1)) for i in range(k)], 1).mean(1)
    total = loss_fn(p_nn, tt[0]) + loss_fn(p_ee, tt[1]) + loss_fn(a * b, tt[2])
    opt.zero_grad()
    total.backward()
    opt.step()
    assert abs(float(total.detach()) - want_loss) < 1e-12
    assert np.abs(tN.detach().numpy() - N).max() < 1e-12
    assert np.abs(tE.detach().numpy() - E).max() < 1e-12
  assert np.abs(N - N0).max() > 1e-3
