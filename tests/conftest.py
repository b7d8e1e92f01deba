import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
  config.addinivalue_line("markers", "reference: needs the reference tree at /root/reference")


def load_golden(name):
  path = os.path.join(GOLDEN_DIR, name + ".npz")
  with np.load(path, allow_pickle=False) as z:
    return {k: z[k] for k in z.files}


def hypergraph_from_pairs(pairs):
  """Builds a Hypergraph proto from (node, edge) pairs the way the reference's loaders do
  (AddNodeToEdge per incidence)."""
  from hypergraphembedding_b200 import Hypergraph
  node_edges, edge_nodes = {}, {}
  for n, e in np.asarray(pairs).tolist():
    node_edges.setdefault(n, []).append(e)
    edge_nodes.setdefault(e, []).append(n)
  hg = Hypergraph()
  for n, edges in node_edges.items():
    hg.node[n].edges.extend(edges)
  for e, nodes in edge_nodes.items():
    hg.edge[e].nodes.extend(nodes)
  return hg


def csr_from_pairs(pairs, shape=None):
  import scipy.sparse as sps
  pairs = np.asarray(pairs)
  m = sps.csr_matrix((np.ones(len(pairs), dtype=bool), (pairs[:, 0], pairs[:, 1])), shape=shape,
                     dtype=bool)
  m.sum_duplicates()
  m.sort_indices()
  return m


def embedding_arrays(embedding, node_ids, edge_ids):
  xn = np.stack([np.asarray(embedding.node[int(i)].values, dtype=np.float32) for i in node_ids])
  xe = np.stack([np.asarray(embedding.edge[int(i)].values, dtype=np.float32) for i in edge_ids])
  return xn, xe


def incidence_distances(A, xn, xe):
  """fp64 L2 distance of every stored incidence (the quantity the parity bar is stated on)."""
  A = A.tocoo()
  d = np.asarray(xn, dtype=np.float64)[A.row] - np.asarray(xe, dtype=np.float64)[A.col]
  return np.sqrt((d * d).sum(axis=1))


@pytest.fixture(scope="session")
def gpu_ctx():
  from hypergraphembedding_b200 import _native
  return _native.default_context()
