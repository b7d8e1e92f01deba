"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the hypergraph2vec training step.

PARITY UNPINNED against Keras itself: the reference builds its models with Keras 2.x on
TensorFlow 1.x (hg2v_model.py:51-203, fit loop embedding.py:269-305); neither is installed here
and the reference's tests hold no numeric vectors for the fit (tests/test_embedding.py only
checks shapes), so this file restates the published semantics of the pinned API calls the
reference makes, and the CUDA trainer is checked against THIS restatement:

  model      two embedding tables with a trainable padding row 0 (Embedding(input_dim=max+2),
             hg2v_model.py:75-84); node_node = act(<N[ln], N[rn]>), edge_edge = act(<E[le], E[re]>),
             node_edge = mean_i act(<N[nn_i], N[ln]>) * mean_i act(<E[ne_i], E[re]>)   (:87-115);
             act = sigmoid for BooleanModel, relu for UnweightedFloatModel (:159-177)
  loss       sum over the three outputs of the batch mean of
             kullback_leibler_divergence: sum(clip(y) * log(clip(y) / clip(p))), clip to [1e-7, 1]
             (keras.losses, K.epsilon() = 1e-7)                       -- BooleanModel (:126)
             mean_squared_error: (p - y)^2                             -- UnweightedFloatModel (:202)
  optimizer  keras.optimizers.Adagrad defaults: lr = 0.01, epsilon = 1e-7, accumulators start at 0,
             a += g^2 ; p -= lr * g / (sqrt(a) + epsilon), on the dense gradient (duplicates of a
             row inside a batch are summed first)
  fit        per epoch np.random.shuffle(arange(M)) from the global numpy RNG, consecutive
             batches of `batch_size` (the last one may be short), epoch loss = sample-weighted
             mean of the batch losses; EarlyStopping(monitor="loss", min_delta=1e-3, patience=0)
             (embedding.py:289-299)

Everything is float64 here except where noted, so it is also the accuracy yardstick for the fp32
kernels.  The restatement is itself checked two independent ways in tests/test_hg2v_host.py:
finite differences of the loss, and the same model written with torch ops + autograd +
torch.optim.Adagrad(lr=0.01, eps=1e-7), which it matches to 1e-12 over several batches.
"""
import numpy as np

EPS = 1e-7
LR = 0.01


def act(z, kind):
  if kind == "sigmoid":
    return 1.0 / (1.0 + np.exp(-z))
  return np.maximum(z, 0.0)


def act_grad(z, p, kind):
  if kind == "sigmoid":
    return p * (1.0 - p)
  return (z > 0).astype(z.dtype)


def loss_and_grad(p, y, kind):
  """Per-sample loss and d loss / d p."""
  if kind == "kld":
    yc = np.clip(y, EPS, 1.0)
    pc = np.clip(p, EPS, 1.0)
    inside = (p >= EPS) & (p <= 1.0)
    return yc * np.log(yc / pc), np.where(inside, -yc / pc, 0.0)
  return (p - y)**2, 2.0 * (p - y)


def batch_step(N, E, accN, accE, feats, targets, k, activation, loss):
  """One Adagrad step on one batch, in place.  feats = [ln, le, rn, re, nn_0.., ne_0..] (int
  arrays of the batch), targets = [nn, ee, ne].  Returns the batch loss (sum over the outputs of
  the batch means)."""
  ln, le, rn, re = feats[0], feats[1], feats[2], feats[3]
  nbr_n = np.stack(feats[4:4 + k], axis=1) if k else np.zeros((len(ln), 0), np.int64)
  nbr_e = np.stack(feats[4 + k:4 + 2 * k], axis=1) if k else np.zeros((len(ln), 0), np.int64)
  m = len(ln)
  Ln, Rn, Le, Re = N[ln], N[rn], E[le], E[re]
  z_nn = (Ln * Rn).sum(1)
  z_ee = (Le * Re).sum(1)
  p_nn, p_ee = act(z_nn, activation), act(z_ee, activation)
  X = N[nbr_n]                                   # [m, k, d]
  Y = E[nbr_e]
  za = np.einsum("mkd,md->mk", X, Ln)
  zb = np.einsum("mkd,md->mk", Y, Re)
  a, b = act(za, activation), act(zb, activation)
  A = a.mean(1) if k else np.zeros(m)
  B = b.mean(1) if k else np.zeros(m)
  p_ne = A * B
  l_nn, g_nn = loss_and_grad(p_nn, targets[0], loss)
  l_ee, g_ee = loss_and_grad(p_ee, targets[1], loss)
  l_ne, g_ne = loss_and_grad(p_ne, targets[2], loss)
  batch_loss = l_nn.mean() + l_ee.mean() + l_ne.mean()
  d_nn = g_nn / m * act_grad(z_nn, p_nn, activation)
  d_ee = g_ee / m * act_grad(z_ee, p_ee, activation)
  gN = np.zeros_like(N)
  gE = np.zeros_like(E)
  np.add.at(gN, ln, d_nn[:, None] * Rn)
  np.add.at(gN, rn, d_nn[:, None] * Ln)
  np.add.at(gE, le, d_ee[:, None] * Re)
  np.add.at(gE, re, d_ee[:, None] * Le)
  if k:
    da = (g_ne / m * B / k)[:, None] * act_grad(za, a, activation)     # [m, k]
    db = (g_ne / m * A / k)[:, None] * act_grad(zb, b, activation)
    np.add.at(gN, ln, np.einsum("mk,mkd->md", da, X))
    np.add.at(gE, re, np.einsum("mk,mkd->md", db, Y))
    np.add.at(gN, nbr_n.ravel(), (da[:, :, None] * Ln[:, None, :]).reshape(-1, N.shape[1]))
    np.add.at(gE, nbr_e.ravel(), (db[:, :, None] * Re[:, None, :]).reshape(-1, E.shape[1]))
  for P, acc, g in ((N, accN, gN), (E, accE, gE)):
    acc += g * g
    P -= LR * g / (np.sqrt(acc) + EPS)
  return batch_loss


def fit(N, E, feats, targets, k, activation, loss, batch_size, epochs, order=None,
        min_delta=1e-3):
  """The fit loop of embedding.py:289-299 on float64 copies of the tables.  `order` (optional):
  one index permutation per epoch instead of np.random.shuffle.  Returns (N, E, epoch losses)."""
  N, E = np.array(N, dtype=np.float64), np.array(E, dtype=np.float64)
  accN, accE = np.zeros_like(N), np.zeros_like(E)
  feats = [np.asarray(f, dtype=np.int64) for f in feats]
  targets = [np.asarray(t, dtype=np.float64) for t in targets]
  m = len(feats[0])
  losses = []
  best = np.inf
  for epoch in range(epochs):
    if order is not None:
      index = np.asarray(order[epoch])
    else:
      index = np.arange(m)
      np.random.shuffle(index)
    total = 0.0
    for lo in range(0, m, batch_size):
      sel = index[lo:lo + batch_size]
      total += len(sel) * batch_step(N, E, accN, accE, [f[sel] for f in feats],
                                     [t[sel] for t in targets], k, activation, loss)
    losses.append(total / m)
    # keras.callbacks.EarlyStopping, mode "min", patience 0: an epoch counts as an improvement
    # only if it beats the best loss by more than min_delta; the first one that does not stops
    if losses[-1] + min_delta < best:
      best = losses[-1]
    else:
      break
  return N, E, losses
