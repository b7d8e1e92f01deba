"""The numpy-fp32 restatement of SparseWeightedJaccard / CentroidFromRows that the GPU test
(tests/test_jaccard_samples_gpu.py) holds the kernels to, pinned against the UNMODIFIED reference
where its tree is present (hg2v_sample.py:250-299); skipped elsewhere (the GPU box)."""
import numpy as np
import pytest
import scipy.sparse as sps

from oracle import ref_shim
from test_jaccard_samples_gpu import _reference_loop

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not present")]


def _row(M, r):
  return M.indices[M.indptr[r]:M.indptr[r + 1]], M.data[M.indptr[r]:M.indptr[r + 1]]


def test_restatement_equals_the_reference_bit_for_bit():
  ref = ref_shim.load_reference()
  F = sps.random(24, 700, density=0.4, random_state=3, format="csr", dtype=np.float32)
  F.data[::17] *= -1          # the formula has no sign restriction
  F.data[::29] = 0            # explicit zeros are not part of nonzero()
  F.sort_indices()
  for a in range(8):
    for b in range(8, 16):
      want = ref.hg2v_sample.SparseWeightedJaccard(F[a], F[b])
      got = _reference_loop(*_row(F, a), *_row(F, b))
      assert np.float32(want) == got and isinstance(want, (np.float32, int)), (a, b)
  # centroid rows: members added in ascending order in fp32, one division
  G = sps.random(6, 24, density=0.4, random_state=4, format="csr", dtype=np.float32)
  G.data[:] = 1
  G.sort_indices()
  for g in range(6):
    members = G.indices[G.indptr[g]:G.indptr[g + 1]]
    if len(members) == 0:
      continue
    vals, (rows, cols) = ref.hg2v_sample.CentroidFromRows(g, G, F)
    mine = F[members].sum(axis=0) / len(members)
    assert mine.dtype == np.float32
    mc = mine.nonzero()[1]
    assert list(cols) == mc.tolist()
    assert np.array_equal(np.asarray(vals, np.float32), np.asarray(mine)[0, mc])
    centroid = sps.csr_matrix((vals, (np.zeros(len(cols), int), cols)), shape=(1, F.shape[1]))
    for a in range(4):
      want = ref.hg2v_sample.SparseWeightedJaccard(F[a], centroid)
      got = _reference_loop(*_row(F, a), np.asarray(cols), np.asarray(vals, np.float32))
      assert np.float32(want) == got
