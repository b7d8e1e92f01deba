"""Synthetic hypergraphs of the shapes BASELINE.json names.

All generators are seeded ``np.random.Generator(PCG64(seed))`` streams and return a scipy
N x E bool CSR incidence matrix in canonical form (sorted, unique column ids) in which every
row and every column is non-empty -- the reference divides 0/0 on an isolated node or edge
(algebraic_distance.py:49).  Ids are shuffled so that no locality is gifted to the gathers.
"""
import numpy as np
import scipy.sparse as sps


def _power_law_probs(n, exponent):
  p = (np.arange(n, dtype=np.float64) + 1.0)**(-exponent)
  return p / p.sum()


def _draw(rng, cdf, count):
  return np.searchsorted(cdf, rng.random(count), side="right").astype(np.int64)


def power_law_hypergraph(num_nodes, num_edges, num_incidences, node_exponent=0.6,
                         edge_exponent=0.8, max_edge_size=100000, seed=1234):
  """Config 2 family: draw (node, edge) pairs with node ~ (rank+1)^-a, edge ~ (rank+1)^-b,
  dedupe, cap the edge size, give every isolated node / edge one uniformly random incidence,
  shuffle ids."""
  rng = np.random.Generator(np.random.PCG64(seed))
  node_cdf = np.cumsum(_power_law_probs(num_nodes, node_exponent))
  edge_cdf = np.cumsum(_power_law_probs(num_edges, edge_exponent))
  node_cdf[-1] = 1.0
  edge_cdf[-1] = 1.0
  keys = np.empty(0, dtype=np.int64)
  want = num_incidences
  while len(keys) < num_incidences:
    draw = int((want - len(keys)) * 1.15) + 1024
    n = np.minimum(_draw(rng, node_cdf, draw), num_nodes - 1)
    e = np.minimum(_draw(rng, edge_cdf, draw), num_edges - 1)
    keys = np.unique(np.concatenate([keys, n * num_edges + e]))
    # cap the largest edges
    e_all = keys % num_edges
    sizes = np.bincount(e_all, minlength=num_edges)
    big = np.nonzero(sizes > max_edge_size)[0]
    if len(big):
      drop = []
      order = np.argsort(e_all, kind="stable")
      starts = np.concatenate([[0], np.cumsum(sizes)])
      for b in big:
        members = order[starts[b]:starts[b + 1]]
        drop.append(rng.choice(members, size=len(members) - max_edge_size, replace=False))
      keep = np.ones(len(keys), dtype=bool)
      keep[np.concatenate(drop)] = False
      keys = keys[keep]
  if len(keys) > num_incidences:
    keys = np.sort(rng.choice(keys, size=num_incidences, replace=False))
  n = keys // num_edges
  e = keys % num_edges
  iso_n = np.nonzero(np.bincount(n, minlength=num_nodes) == 0)[0]
  iso_e = np.nonzero(np.bincount(e, minlength=num_edges) == 0)[0]
  n = np.concatenate([n, iso_n, rng.integers(0, num_nodes, len(iso_e))])
  e = np.concatenate([e, rng.integers(0, num_edges, len(iso_n)), iso_e])
  node_perm = rng.permutation(num_nodes)
  edge_perm = rng.permutation(num_edges)
  return _to_csr(node_perm[n], edge_perm[e], num_nodes, num_edges)


def bipartite_author_paper(num_authors=1700000, num_papers=2000000, mean_extra_authors=2.0,
                           max_paper_size=50, author_exponent=0.7, seed=4321):
  """Config 3 family (AMiner-shaped): every paper (edge) has 1 + Poisson(mean) authors
  (capped), authors drawn ~ (rank+1)^-a; isolated authors get one random paper."""
  rng = np.random.Generator(np.random.PCG64(seed))
  sizes = np.minimum(1 + rng.poisson(mean_extra_authors, num_papers), max_paper_size)
  e = np.repeat(np.arange(num_papers, dtype=np.int64), sizes)
  cdf = np.cumsum(_power_law_probs(num_authors, author_exponent))
  cdf[-1] = 1.0
  n = np.minimum(_draw(rng, cdf, len(e)), num_authors - 1)
  iso_n = np.nonzero(np.bincount(n, minlength=num_authors) == 0)[0]
  n = np.concatenate([n, iso_n])
  e = np.concatenate([e, rng.integers(0, num_papers, len(iso_n))])
  node_perm = rng.permutation(num_authors)
  edge_perm = rng.permutation(num_papers)
  return _to_csr(node_perm[n], edge_perm[e], num_authors, num_papers)


def community_hypergraph(num_nodes, num_edges, num_incidences, size_exponent=1.5, min_size=3,
                         max_size=1000000, seed=99):
  """Config 5 family (SNAP-community-shaped): edge sizes ~ power law on [min, max] scaled to
  the incidence budget, members uniform; isolated nodes get one random community."""
  rng = np.random.Generator(np.random.PCG64(seed))
  u = rng.random(num_edges)
  a = 1.0 - size_exponent
  raw = ((max_size**a - min_size**a) * u + min_size**a)**(1.0 / a)
  sizes = np.maximum(min_size, np.minimum(max_size, raw * (num_incidences / raw.sum())))
  sizes = np.minimum(sizes.astype(np.int64), num_nodes)
  e = np.repeat(np.arange(num_edges, dtype=np.int64), sizes)
  n = rng.integers(0, num_nodes, len(e))
  iso_n = np.nonzero(np.bincount(n, minlength=num_nodes) == 0)[0]
  n = np.concatenate([n, iso_n])
  e = np.concatenate([e, rng.integers(0, num_edges, len(iso_n))])
  return _to_csr(n, e, num_nodes, num_edges)


def zipf_hypergraph(num_nodes=10000000, num_edges=5000000, zipf_exponent=2.2, max_degree=1000,
                    edge_exponent=0.5, max_edge_size=1000, seed=2024):
  """Config 4 family (FOBE sampling at scale): node degree ~ Zipf(a) >= 1 capped at
  `max_degree`, edges drawn ~ (rank+1)^-b with the edge size capped at `max_edge_size` (members
  beyond the cap are re-drawn uniformly), which keeps sum(size^2) -- the size of the candidate
  rows of A * A.T -- bounded; every edge non-empty; ids shuffled."""
  rng = np.random.Generator(np.random.PCG64(seed))
  deg = np.minimum(rng.zipf(zipf_exponent, num_nodes), max_degree).astype(np.int64)
  n = np.repeat(np.arange(num_nodes, dtype=np.int64), deg)
  cdf = np.cumsum(_power_law_probs(num_edges, edge_exponent))
  cdf[-1] = 1.0
  e = np.minimum(_draw(rng, cdf, len(n)), num_edges - 1)
  for _ in range(8):
    sizes = np.bincount(e, minlength=num_edges)
    if sizes.max() <= max_edge_size:
      break
    # rank of every incidence inside its edge; those beyond the cap move to a uniform edge
    order = np.argsort(e, kind="stable")
    starts = np.concatenate([[0], np.cumsum(sizes)])[:-1]
    rank = np.empty(len(e), dtype=np.int64)
    rank[order] = np.arange(len(e), dtype=np.int64) - np.repeat(starts, sizes)
    over = np.nonzero(rank >= max_edge_size)[0]
    e[over] = rng.integers(0, num_edges, len(over))
  iso_e = np.nonzero(np.bincount(e, minlength=num_edges) == 0)[0]
  n = np.concatenate([n, rng.integers(0, num_nodes, len(iso_e))])
  e = np.concatenate([e, iso_e])
  node_perm = rng.permutation(num_nodes)
  edge_perm = rng.permutation(num_edges)
  return _to_csr(node_perm[n], edge_perm[e], num_nodes, num_edges)


def induced_on_first_nodes(A, num_nodes):
  """Sub-hypergraph induced by nodes [0, num_nodes): their rows, and only the edges they touch,
  relabelled densely in ascending id order (what CompressRange would produce).  Returns the
  canonical CSR."""
  sub = sps.csr_matrix(A)[:num_nodes]
  used = np.unique(sub.indices)
  relabel = np.full(A.shape[1], -1, dtype=np.int64)
  relabel[used] = np.arange(len(used))
  m = sps.csr_matrix((np.ones(sub.nnz, dtype=bool), relabel[sub.indices].astype(np.int32),
                      sub.indptr), shape=(num_nodes, len(used)))
  m.sort_indices()
  return m


def _to_csr(n, e, num_nodes, num_edges):
  m = sps.csr_matrix((np.ones(len(n), dtype=bool), (n, e)), shape=(num_nodes, num_edges),
                     dtype=bool)
  m.sum_duplicates()
  m.sort_indices()
  return m


def legacy_initial_vectors(num_nodes, num_edges, dimension, seed):
  """The reference's initial vectors (algebraic_distance.py:140-141) for a given legacy seed,
  cast to fp32: the oracle and the kernels start from identical values."""
  np.random.seed(seed)
  xn = np.random.random((num_nodes, dimension)).astype(np.float32)
  xe = np.random.random((num_edges, dimension)).astype(np.float32)
  return xn, xe
