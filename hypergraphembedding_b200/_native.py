"""ctypes binding of libhge_b200.so (include/hge_b200.h).

There is deliberately no fallback here: if the shared library is missing or no CUDA
device is usable, the first call raises.  Build the library with
``python -m hypergraphembedding_b200.build`` (or ``__graft_entry__.build()``).
"""
import ctypes
import os
import threading

import numpy as np

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libhge_b200.so")

HGE_OK = 0
HGE_ERR_INVALID = -1
HGE_ERR_CUDA = -2
HGE_ERR_EMPTY_ROW = -3
HGE_ERR_NOMEM = -4
HGE_ERR_UNSUPPORTED = -5
MEM_HOST = 0
MEM_DEVICE = 1

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_u32p = ctypes.POINTER(ctypes.c_uint32)
c_f32p = ctypes.POINTER(ctypes.c_float)
c_f64p = ctypes.POINTER(ctypes.c_double)
c_vp = ctypes.c_void_p


class NativeLibraryMissing(ImportError):
  pass


class NativeError(RuntimeError):
  pass


_lib = None
_lock = threading.Lock()

# name -> (restype, argtypes); checked against include/hge_b200.h by tests/test_abi.py
SIGNATURES = {
    "hge_version": (ctypes.c_int, []),
    "hge_last_error": (ctypes.c_char_p, []),
    "hge_ctx_create": (ctypes.c_int, [ctypes.c_int, c_vp, ctypes.POINTER(c_vp)]),
    "hge_ctx_destroy": (ctypes.c_int, [c_vp]),
    "hge_ctx_set_stream": (ctypes.c_int, [c_vp, c_vp]),
    "hge_ctx_sync": (ctypes.c_int, [c_vp]),
    "hge_ctx_set_tuning": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "hge_ctx_launch_count": (ctypes.c_int64, [c_vp]),
    "hge_incidence_create": (ctypes.c_int, [c_vp, ctypes.c_int32, ctypes.c_int32, c_vp, c_vp, c_vp,
                                            c_vp, ctypes.c_int, ctypes.POINTER(c_vp)]),
    "hge_incidence_destroy": (ctypes.c_int, [c_vp]),
    "hge_incidence_nnz": (ctypes.c_int64, [c_vp]),
    "hge_algdist_run": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, c_vp]),
    "hge_algdist_create": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, ctypes.c_int,
                                          ctypes.POINTER(c_vp)]),
    "hge_algdist_destroy": (ctypes.c_int, [c_vp]),
    "hge_algdist_load": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_int]),
    "hge_algdist_node_half": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "hge_algdist_edge_half": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "hge_algdist_edge_partial": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp]),
    "hge_algdist_edge_finalize": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, c_vp]),
    "hge_algdist_minmax_ptr": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.POINTER(c_vp)]),
    "hge_algdist_ld": (ctypes.c_int, [c_vp]),
    "hge_algdist_store": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, c_vp, ctypes.c_int]),
}


def load_library():
  """Loads libhge_b200.so once.  Raises NativeLibraryMissing when it has not been built."""
  global _lib
  with _lock:
    if _lib is not None:
      return _lib
    if not os.path.exists(LIB_PATH):
      raise NativeLibraryMissing(
          "%s is missing: build it with `python -m hypergraphembedding_b200.build`. "
          "hypergraphembedding_b200 has no CPU or PyTorch fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
      fn = getattr(lib, name)  # AttributeError if the .so is stale
      fn.restype = restype
      fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error():
  msg = load_library().hge_last_error()
  return msg.decode("utf-8", "replace") if msg else ""


def check(rc, what=""):
  """Maps a status code to the exception the reference would raise at that point."""
  if rc == HGE_OK:
    return
  msg = "%s%s" % (what + ": " if what else "", last_error())
  if rc == HGE_ERR_INVALID:
    raise AssertionError(msg)           # the reference validates with assert
  if rc == HGE_ERR_EMPTY_ROW:
    raise ZeroDivisionError(msg)        # algebraic_distance.py:49 divides 0/0
  if rc == HGE_ERR_NOMEM:
    raise MemoryError(msg)
  raise NativeError("%s (status %d)" % (msg, rc))


def ptr(x):
  """Address of a numpy array / torch tensor / None as a void pointer."""
  if x is None:
    return None
  if isinstance(x, np.ndarray):
    assert x.flags["C_CONTIGUOUS"], "array must be C-contiguous"
    return x.ctypes.data_as(c_vp)
  if hasattr(x, "data_ptr"):
    assert x.is_contiguous(), "tensor must be contiguous"
    return c_vp(x.data_ptr())
  if isinstance(x, int):
    return c_vp(x)
  raise TypeError("cannot take the address of %r" % type(x))


def is_device(x):
  return hasattr(x, "is_cuda") and x.is_cuda


class Context(object):
  """One hge_ctx per (device, stream)."""

  def __init__(self, device=0, stream=None):
    self.lib = load_library()
    self.device = int(device)
    handle = c_vp()
    stream_ptr = c_vp(int(stream)) if stream else None
    check(self.lib.hge_ctx_create(self.device, stream_ptr, ctypes.byref(handle)), "hge_ctx_create")
    self.handle = handle

  def set_stream(self, stream):
    check(self.lib.hge_ctx_set_stream(self.handle, c_vp(int(stream)) if stream else None))

  def set_tuning(self, light_max_deg=0, chunk=0, blocks_per_sm=0):
    check(self.lib.hge_ctx_set_tuning(self.handle, light_max_deg, chunk, blocks_per_sm))

  def sync(self):
    check(self.lib.hge_ctx_sync(self.handle), "hge_ctx_sync")

  @property
  def launch_count(self):
    return int(self.lib.hge_ctx_launch_count(self.handle))

  def close(self):
    if getattr(self, "handle", None):
      self.lib.hge_ctx_destroy(self.handle)
      self.handle = None

  def __del__(self):
    try:
      self.close()
    except Exception:
      pass


_default_ctx = {}


def default_context(device=None):
  """Process-wide context on the current torch device's current stream (or device 0 with a
  private stream when torch has not initialised CUDA)."""
  stream = None
  if device is None:
    device = 0
    try:
      import torch
      if torch.cuda.is_available():
        device = torch.cuda.current_device()
    except ImportError:
      pass
  key = int(device)
  ctx = _default_ctx.get(key)
  if ctx is None:
    ctx = Context(key, stream)
    _default_ctx[key] = ctx
  return ctx


def _as_i64(a):
  return np.ascontiguousarray(a, dtype=np.int64)


def _as_i32(a):
  return np.ascontiguousarray(a, dtype=np.int32)


class Incidence(object):
  """Device-resident incidence (hge_incidence): int32 CSR of node->edge and edge->node."""

  def __init__(self, ctx, num_nodes, num_edges, n2e_ptr, n2e_idx, e2n_ptr, e2n_idx):
    self.ctx = ctx
    self.num_nodes = int(num_nodes)
    self.num_edges = int(num_edges)
    device = is_device(n2e_idx)
    if device:
      keep = (n2e_ptr, n2e_idx, e2n_ptr, e2n_idx)   # borrowed by the library
      import torch
      assert n2e_ptr.dtype == torch.int64 and e2n_ptr.dtype == torch.int64
      assert n2e_idx.dtype == torch.int32 and e2n_idx.dtype == torch.int32
    else:
      keep = (_as_i64(n2e_ptr), _as_i32(n2e_idx), _as_i64(e2n_ptr), _as_i32(e2n_idx))
    self._keep = keep
    handle = c_vp()
    check(ctx.lib.hge_incidence_create(ctx.handle, self.num_nodes, self.num_edges, ptr(keep[0]),
                                       ptr(keep[1]), ptr(keep[2]), ptr(keep[3]),
                                       MEM_DEVICE if device else MEM_HOST, ctypes.byref(handle)),
          "hge_incidence_create")
    self.handle = handle
    if not device:
      self._keep = None  # the library copied the arrays

  @property
  def nnz(self):
    return int(self.ctx.lib.hge_incidence_nnz(self.handle))

  def close(self):
    if getattr(self, "handle", None):
      self.ctx.lib.hge_incidence_destroy(self.handle)
      self.handle = None

  def __del__(self):
    try:
      self.close()
    except Exception:
      pass


def algdist_run(ctx, inc, xn, xe, iterations, lohi=None):
  """hge_algdist_run on numpy (host) or torch CUDA (device) fp32 [rows, R] arrays, in place."""
  device = is_device(xn)
  assert device == is_device(xe)
  R = int(xn.shape[1])
  assert tuple(xn.shape) == (inc.num_nodes, R) and tuple(xe.shape) == (inc.num_edges, R)
  if not device:
    assert xn.dtype == np.float32 and xe.dtype == np.float32
  if lohi is not None:
    assert lohi.dtype == np.float32 and lohi.shape == (iterations, 2, R)
  check(ctx.lib.hge_algdist_run(ctx.handle, inc.handle, ptr(xn), ptr(xe), R, int(iterations),
                                MEM_DEVICE if device else MEM_HOST, ptr(lohi)), "hge_algdist_run")
  return xn, xe


class AlgDistState(object):
  """Stepwise relaxation (hge_algdist_*), device pointers only."""

  def __init__(self, ctx, inc, R, max_iterations):
    self.ctx, self.inc, self.R = ctx, inc, int(R)
    handle = c_vp()
    check(ctx.lib.hge_algdist_create(ctx.handle, inc.handle, self.R, int(max_iterations),
                                     ctypes.byref(handle)), "hge_algdist_create")
    self.handle = handle
    self.ld = int(ctx.lib.hge_algdist_ld(handle))

  def load(self, xn, xe):
    mem = MEM_DEVICE if is_device(xn) else MEM_HOST
    check(self.ctx.lib.hge_algdist_load(self.handle, ptr(xn), ptr(xe), mem), "hge_algdist_load")

  def node_half(self, sweep):
    check(self.ctx.lib.hge_algdist_node_half(self.handle, sweep), "hge_algdist_node_half")

  def edge_half(self, sweep):
    check(self.ctx.lib.hge_algdist_edge_half(self.handle, sweep), "hge_algdist_edge_half")

  def edge_partial(self, sweep, partial):
    check(self.ctx.lib.hge_algdist_edge_partial(self.handle, sweep, ptr(partial)))

  def edge_finalize(self, sweep, partial, inv_s_edge=None):
    check(self.ctx.lib.hge_algdist_edge_finalize(self.handle, sweep, ptr(partial), ptr(inv_s_edge)))

  def minmax_ptr(self, sweep):
    out = c_vp()
    check(self.ctx.lib.hge_algdist_minmax_ptr(self.handle, sweep, ctypes.byref(out)))
    return out.value

  def store(self, sweeps_done, xn, xe):
    mem = MEM_DEVICE if is_device(xn) else MEM_HOST
    check(self.ctx.lib.hge_algdist_store(self.handle, sweeps_done, ptr(xn), ptr(xe), mem),
          "hge_algdist_store")

  def close(self):
    if getattr(self, "handle", None):
      self.ctx.lib.hge_algdist_destroy(self.handle)
      self.handle = None

  def __del__(self):
    try:
      self.close()
    except Exception:
      pass
