"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/algdist_ref.c (built by
``make -C oracle`` / ``__graft_entry__.build()`` into oracle/_build/liboracle.so)."""
import ctypes
import os
import subprocess

import numpy as np
import scipy.sparse as sps

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "liboracle.so")
_lib = None


def load(build_if_missing=True):
  global _lib
  if _lib is None:
    if not os.path.exists(LIB_PATH) and build_if_missing:
      subprocess.check_call(["make", "-s", "-C", HERE])
    _lib = ctypes.CDLL(LIB_PATH)
    _lib.oracle_algdist.restype = ctypes.c_int
    _lib.oracle_algdist.argtypes = [ctypes.c_int64, ctypes.c_int64] + [ctypes.c_void_p] * 4 + [
        ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    _lib.oracle_max_threads.restype = ctypes.c_int
  return _lib


def max_threads():
  return int(load().oracle_max_threads())


def algdist(A, B, xn, xe, iterations, threads=0):
  """f64 relaxation of (xn, xe) copies; returns the new arrays.  threads=0: all cores."""
  lib = load()
  A = sps.csr_matrix(A)
  B = sps.csr_matrix(B)
  a_ptr = np.ascontiguousarray(A.indptr, dtype=np.int64)
  a_idx = np.ascontiguousarray(A.indices, dtype=np.int32)
  b_ptr = np.ascontiguousarray(B.indptr, dtype=np.int64)
  b_idx = np.ascontiguousarray(B.indices, dtype=np.int32)
  xn = np.array(xn, dtype=np.float64, order="C", copy=True)
  xe = np.array(xe, dtype=np.float64, order="C", copy=True)
  p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
  rc = lib.oracle_algdist(A.shape[0], A.shape[1], p(a_ptr), p(a_idx), p(b_ptr), p(b_idx),
                          xn.shape[1], int(iterations), p(xn), p(xe), int(threads))
  if rc != 0:
    raise MemoryError("oracle_algdist failed")
  return xn, xe
