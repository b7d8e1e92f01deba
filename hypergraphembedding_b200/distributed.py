"""Multi-GPU relaxation: the NODES (and their incidences) are 1-D row-partitioned over the ranks,
one process per GPU, balanced by incidence count; the E x R edge block is replicated
(SURVEY.md section 8e).

Per sweep and rank:
  node half   purely local: gathers the replicated edge rows.
  edge half   each rank sums  w_n * xn'[n]  over its LOCAL members of every edge
              (hge_algdist_edge_partial), the partial sums are all-reduced over NVLink
              (NCCL), and every rank blends / rescales all edge rows redundantly
              (hge_algdist_edge_finalize).  The edge rows are processed in slices so the
              all-reduce of slice k runs while slice k+1 is being gathered.
  rescale     the per-column (min, max) of the sweep -- local nodes + all edges -- is
              all-reduced (MIN / MAX on order-preserving int32 encodings, 2 x R words).
One-time set-up collectives: edge degrees (SUM of local counts) and the edges' inverse weight
sums.  Nothing else crosses GPUs: the 8.3 GB of node rows of config 5 never move.

``ShardedRelaxation`` takes the per-rank kernels as an ``ops`` object; the product default is
``NativeOps`` (libhge_b200.so).  tests/test_distributed_gloo.py drives the same orchestration
over ``gloo`` on CPU with a numpy stand-in for the kernels (test infrastructure, not a fallback).
"""
import os

import numpy as np
import scipy.sparse as sps

from . import _native


def partition_rows_by_nnz(indptr, parts):
  """Boundaries b[0..parts] of contiguous row blocks with (almost) equal incidence counts."""
  indptr = np.asarray(indptr, dtype=np.int64)
  rows = len(indptr) - 1
  targets = indptr[-1] * np.arange(1, parts, dtype=np.float64) / parts
  cuts = np.searchsorted(indptr, targets, side="left")
  bounds = np.concatenate([[0], np.clip(cuts, 0, rows), [rows]]).astype(np.int64)
  return np.maximum.accumulate(bounds)


def local_shard(node2edges, rank, world):
  """(A_local, first_row, last_row): this rank's block of node rows of the N x E incidence."""
  A = sps.csr_matrix(node2edges)
  bounds = partition_rows_by_nnz(A.indptr, world)
  r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
  return A[r0:r1], r0, r1


class _CudaView(object):
  """Zero-copy torch view of library-owned device memory (__cuda_array_interface__)."""

  def __init__(self, ptr, shape, typestr):
    self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr,
                                     "data": (int(ptr), False), "version": 2}


# Exchange arenas are expensive to set up (cudaMalloc + memset, IPC handle swap, mapping of every
# peer, two host barriers on tear-down), so they are pooled per process and re-used by later
# relaxations of the same shape; all ranks run the same call sequence, so they hit and miss the
# pool together.  release_peer_arenas() is the collective tear-down.
_ARENA_POOL = {}
_STATUS_CHANNEL = {}


def _status_channel(torch):
  """(side stream, int32[1] device tensor) of the current device for the status agreements."""
  dev = torch.cuda.current_device()
  if dev not in _STATUS_CHANNEL:
    _STATUS_CHANNEL[dev] = (torch.cuda.Stream(device=dev), torch.zeros(1, dtype=torch.int32, device="cuda"))
  return _STATUS_CHANNEL[dev]
_TOPOLOGY_OK = {}


def release_peer_arenas(dist=None, group=None):
  """Collective: unmaps and frees every pooled exchange arena that is not attached to a live
  relaxation.  Call on all ranks before the process group is destroyed (otherwise the memory is
  released at process exit)."""
  entries = []
  for key in list(_ARENA_POOL):
    entries += [e for e in _ARENA_POOL[key] if not e["in_use"]]
    _ARENA_POOL[key] = [e for e in _ARENA_POOL[key] if e["in_use"]]
    if not _ARENA_POOL[key]:
      del _ARENA_POOL[key]
  if not entries:
    return
  if dist is None:
    import torch.distributed as dist
  for e in entries:
    e["ctx"].sync()
  dist.barrier(group=group)
  for e in entries:
    e["arena"].close_peers()
  dist.barrier(group=group)
  for e in entries:
    e["arena"].close()


class NativeOps(object):
  """The per-rank kernels of libhge_b200.so behind the interface ShardedRelaxation drives."""

  def __init__(self, A_local, R, iterations, num_slices, ctx=None, B_local=None, shape=None,
               csr_device=None, csr_host=None):
    """A_local: scipy n_local x E incidence block (B_local: its transpose, optional), or, with
    shape=(n_local, E), csr_device = (n2e_ptr, n2e_idx, e2n_ptr, e2n_idx) torch CUDA tensors or
    csr_host = the same four as int64 / int32 numpy arrays (e.g. views of pinned memory, which
    upload at full PCIe speed; scipy's own arrays are pageable)."""
    import torch
    self.torch = torch
    self.ctx = ctx or _native.default_context()
    self.device = torch.device("cuda", self.ctx.device)
    self.R, self.iterations = R, iterations
    if csr_device is not None or csr_host is not None:
      n_loc, num_edges = shape
      self.inc = _native.Incidence(self.ctx, n_loc, num_edges, *(csr_device or csr_host),
                                   sharded=True, num_slices=num_slices)
    else:
      A = sps.csr_matrix(A_local)
      if B_local is None:
        B = A.T.tocsr()
        B.sort_indices()
      else:
        B = B_local
      n_loc, num_edges = A.shape
      self.inc = _native.Incidence(self.ctx, n_loc, num_edges,
                                   np.asarray(A.indptr, np.int64), np.asarray(A.indices, np.int32),
                                   np.asarray(B.indptr, np.int64), np.asarray(B.indices, np.int32),
                                   sharded=True, num_slices=num_slices)
    self.num_slices = num_slices
    self.num_edges = num_edges
    self.num_local_nodes = n_loc
    self.state = None

  def edge_sums(self):
    """Torch views of the shard's local edge degrees (int32 [E]) and weight sums (f64 [E])."""
    deg_ptr, wsum_ptr = self.inc.edge_sums()
    deg = self.torch.as_tensor(_CudaView(deg_ptr, (self.num_edges,), "<i4"), device=self.device)
    wsum = self.torch.as_tensor(_CudaView(wsum_ptr, (self.num_edges,), "<f8"), device=self.device)
    return deg, wsum

  def finish(self):
    self.inc.finish_sharded()
    self.state = _native.AlgDistState(self.ctx, self.inc, self.R, self.iterations)
    self.ld = self.state.ld

  def new_partial_buffer(self):
    return self.torch.empty((self.num_edges, self.ld), dtype=self.torch.float32, device=self.device)

  def slice_range(self, k):
    return self.inc.slice_range(k)

  def load(self, xn, xe):
    self.state.load(xn, xe)

  def node_half(self, t):
    self.state.node_half(t)

  def edge_partial(self, t, k, partial):
    self.state.edge_partial(t, k, partial)

  def edge_finalize(self, t, k, partial):
    self.state.edge_finalize(t, k, partial)

  def minmax(self, t):
    """int32 [2, ld] torch view of sweep t's encoded (min, max) slots."""
    view = _CudaView(self.state.minmax_ptr(t), (2, self.ld), "<i4")
    return self.torch.as_tensor(view, device=self.device)

  def store(self, sweeps_done, xn, xe):
    self.state.store(sweeps_done, xn, xe)

  # ---- peer-memory exchange (csrc/hge_p2p.cu) ------------------------------------------
  def enable_p2p(self, dist, group, pooled=True, agree=None):
    """Attaches this rank's exchange arena to the relaxation state.  A pooled arena of the same
    shape is re-used; otherwise one is created, its 64-byte IPC handle swapped with the other
    ranks and the peers' arenas mapped.  `agree(fn, what)` makes a rank-local failure of the
    arena allocation collective (ShardedRelaxation._all_or_none)."""
    agree = agree or (lambda fn, what: fn())
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    key = (id(self.ctx), id(group), rank, world, self.num_local_nodes, self.num_edges, self.ld)
    # several relaxations of one shape can be alive at once (all ranks create and close them in
    # the same order, so they pick the same pooled arena or miss together)
    free = [e for e in _ARENA_POOL.get(key, []) if not e["in_use"]] if pooled else []
    entry = free[0] if free else None
    if entry is not None:
      self.arena = entry["arena"]
    else:
      self.arena = agree(lambda: _native.PeerArena(self.ctx, rank, world, self.num_local_nodes,
                                                   self.num_edges, self.ld), "allocating the exchange arena")
      mine = self.torch.from_numpy(self.arena.export()).to(self.device)
      every = [self.torch.empty_like(mine) for _ in range(world)]
      dist.all_gather(every, mine, group=group)
      handles = self.torch.stack(every).cpu().numpy()
      agree(lambda: self.arena.open_peers(handles), "mapping the peers' exchange arenas")
      dist.barrier(group=group)
      entry = {"arena": self.arena, "ctx": self.ctx, "in_use": False}
      if pooled:
        _ARENA_POOL.setdefault(key, []).append(entry)
    entry["in_use"] = True
    self._arena_entry = entry
    self._arena_pooled = pooled
    self.state.attach_p2p(self.arena)
    self._dist, self._group = dist, group

  def sweep_p2p(self, t):
    self.state.sweep_p2p(t)

  def check_p2p(self):
    self.arena.check()

  def close(self):
    arena = getattr(self, "arena", None)
    pooled = arena is not None and self._arena_pooled
    if arena is not None and not pooled:
      # nobody may free an arena a peer still has mapped or is still writing to
      self.ctx.sync()
      self._dist.barrier(group=self._group)
      arena.close_peers()
      self._dist.barrier(group=self._group)
    if self.state is not None:
      if pooled:
        self.ctx.sync()        # the state's kernels may still be writing into the arena
      self.state.close()
      self.state = None
    if arena is not None:
      if pooled:
        self._arena_entry["in_use"] = False
      else:
        arena.close()
      self.arena = None
    self.inc.close()


class ShardedRelaxation(object):
  """Row-partitioned algebraic-distance relaxation over a torch.distributed process group."""

  def __init__(self, A_local, R, iterations, group=None, num_slices=1, ops_factory=None,
               comm="auto", **ops_kwargs):
    import torch
    import torch.distributed as dist
    self.torch, self.dist, self.group = torch, dist, group
    self.R, self.iterations = int(R), int(iterations)
    self.check_barriers = True
    self.num_local_nodes, self.num_edges = ops_kwargs.get("shape") or A_local.shape
    num_slices = max(1, min(int(num_slices), self.num_edges))
    factory = ops_factory or NativeOps
    ctx = ops_kwargs.get("ctx")
    if ctx is not None and hasattr(ctx, "bind_torch_stream"):
      ctx.bind_torch_stream()     # the collectives below are ordered with the kernels by stream
    self.ops = self._all_or_none(lambda: factory(A_local, self.R, self.iterations, num_slices, **ops_kwargs),
                                 "uploading the shard")
    # the one set-up exchange: global edge degrees and the edges' weight sums
    deg, wsum = self.ops.edge_sums()
    w0 = dist.all_reduce(deg, op=dist.ReduceOp.SUM, group=group, async_op=True)
    w1 = dist.all_reduce(wsum, op=dist.ReduceOp.SUM, group=group, async_op=True)
    w0.wait()
    w1.wait()
    if bool((deg == 0).any()):       # all-reduced: every rank sees it
      raise ZeroDivisionError("an edge has no incidence on any rank (algebraic_distance.py:49)")
    # an isolated NODE is only seen by the rank that holds it: agree before the next collective
    self._all_or_none(self.ops.finish, "building the shard's schedule")
    self.num_slices = num_slices
    # exchange strategy: "p2p" = fused into the kernels over peer memory (one node, one GPU
    # per rank), "nccl" = host-interleaved NCCL / gloo collectives
    assert comm in ("auto", "p2p", "nccl")
    can_p2p = hasattr(self.ops, "enable_p2p") and dist.get_backend(group) == "nccl" and \
        self._one_gpu_per_rank_on_one_node()   # cached per group
    if comm == "p2p" and not can_p2p:
      raise RuntimeError("peer-memory exchange needs NCCL ranks on distinct GPUs of one node")
    self.use_p2p = can_p2p and comm in ("auto", "p2p")
    if self.use_p2p:
      self.ops.enable_p2p(dist, group, agree=self._all_or_none)
      self.partial = None
    else:
      self.partial = self.ops.new_partial_buffer()

  def _all_or_none(self, fn, what):
    """Runs a rank-local step that may fail (an isolated node, out of memory, a CUDA error) and
    makes the outcome collective: the status word is all-reduced (MAX), and when any rank failed
    EVERY rank raises before the next collective instead of waiting in it for the NCCL timeout.
    A ZeroDivisionError (the reference's 0/0 on an isolated node, algebraic_distance.py:49) is
    raised as such on all ranks."""
    torch, dist = self.torch, self.dist
    result, error, code = None, None, 0
    try:
      result = fn()
    except ZeroDivisionError as exc:
      error, code = exc, 1
    except Exception as exc:       # noqa: BLE001 -- re-raised below, on every rank
      error, code = exc, 2
    if dist.get_backend(self.group) == "nccl":
      # The outcome of fn is known on the host; the agreement must not wait for the set-up work fn
      # queued on the main stream (a .item() there drains the whole pipeline three times per
      # construction and was the source of 10-30 ms hiccups of the host-buffer step).  It runs on
      # a side stream with a persistent status word: NCCL orders the collective after that stream
      # only, and .item() synchronises that stream only.
      side, status = _status_channel(torch)
      with torch.cuda.stream(side):
        status.fill_(code)
        dist.all_reduce(status, op=dist.ReduceOp.MAX, group=self.group)
        worst = int(status.item())
    else:
      status = torch.tensor([code], dtype=torch.int32)
      dist.all_reduce(status, op=dist.ReduceOp.MAX, group=self.group)
      worst = int(status.item())
    if error is not None:
      raise error
    if worst == 1:
      raise ZeroDivisionError("a node on another rank has no incidence (algebraic_distance.py:49)")
    if worst:
      raise RuntimeError("another rank failed while %s; see its log" % what)
    return result

  def _one_gpu_per_rank_on_one_node(self):
    import socket
    torch, dist = self.torch, self.dist
    key = (id(self.group), torch.cuda.current_device())
    if key in _TOPOLOGY_OK:
      return _TOPOLOGY_OK[key]
    props = torch.cuda.get_device_properties(torch.cuda.current_device())
    mine = (socket.gethostname(), str(getattr(props, "uuid", torch.cuda.current_device())))
    every = [None] * dist.get_world_size(self.group)
    dist.all_gather_object(every, mine, group=self.group)
    _TOPOLOGY_OK[key] = len({h for h, _ in every}) == 1 and len({u for _, u in every}) == len(every)
    return _TOPOLOGY_OK[key]

  def sweep(self, t):
    if self.use_p2p:
      self.ops.sweep_p2p(t)
      return
    self.ops.node_half(t)
    self.sweep_after_node_half(t)

  def sweep_after_node_half(self, t):
    dist, ops = self.dist, self.ops
    pending = []
    for k in range(self.num_slices):
      ops.edge_partial(t, k, self.partial)
      r0, r1 = ops.slice_range(k)
      # NCCL orders the collective after the partial-sum kernel already queued on this stream
      # and runs it on its own stream, so it overlaps the next slice's gather
      pending.append(dist.all_reduce(self.partial[r0:r1], op=dist.ReduceOp.SUM, group=self.group,
                                     async_op=True))
    for k in range(self.num_slices):
      pending[k].wait()
      ops.edge_finalize(t, k, self.partial)
    mm = ops.minmax(t)
    w0 = dist.all_reduce(mm[0], op=dist.ReduceOp.MIN, group=self.group, async_op=True)
    w1 = dist.all_reduce(mm[1], op=dist.ReduceOp.MAX, group=self.group, async_op=True)
    w0.wait()
    w1.wait()

  def run(self, xn_local, xe):
    """In place: xn_local [n_local, R] (this rank's node rows) and xe [E, R] (replicated, must
    be identical on every rank)."""
    if self.iterations == 0:
      return xn_local, xe
    ctx = getattr(self.ops, "ctx", None)
    if ctx is not None and hasattr(ctx, "bind_torch_stream"):
      ctx.bind_torch_stream()
    self.ops.load(xn_local, xe)
    for t in range(self.iterations):
      self.sweep(t)
      if t == 0 and self.use_p2p and self.check_barriers:
        # a rank that never arrives shows at the first barrier: stop here, not after
        # `iterations` sweeps over rows that were never exchanged
        self.ops.check_p2p()
    self.ops.store(self.iterations, xn_local, xe)
    if self.use_p2p and self.check_barriers:
      self.ops.check_p2p()
    return xn_local, xe

  def close(self):
    self.ops.close()


def sharded_pair_weights(xa, xb, ia_local, ib_local, alpha, group=None, ctx=None, ops=None):
  """Distance weights alpha + (1 - alpha) * (1 - zero_one(||xa[i] - xb[j]||)) of pairs that are
  sharded over the ranks (embarrassingly parallel over the pairs once the
  vectors are replicated).  Every rank passes its own slice of the pair list and the full,
  identical vector blocks; the only exchange is the all-reduce of (min, max) of the distances,
  so the weights equal those of one call over the concatenated list
  (hg2v_weighting.py:40-45, 94-95).  Returns this rank's weights."""
  import torch
  import torch.distributed as dist
  ops = ops or _NativePairOps(ctx or _native.default_context())
  d = ops.pair_l2(xa, xb, ia_local, ib_local)
  lo, hi = ops.minmax(d)
  mm = torch.tensor([lo, -hi], dtype=torch.float64)
  if dist.get_backend(group) == "nccl":
    mm = mm.cuda()
  dist.all_reduce(mm, op=dist.ReduceOp.MIN, group=group)
  lo, hi = float(mm[0]), -float(mm[1])
  if not lo <= hi:       # no pair anywhere
    return d
  return ops.apply(d, alpha, lo, hi)


class _NativePairOps(object):
  def __init__(self, ctx):
    self.ctx = ctx

  def pair_l2(self, xa, xb, ia, ib):
    return _native.pair_l2(self.ctx, xa, xb, ia, ib)

  def minmax(self, d):
    return _native.scale_minmax(self.ctx, d)

  def apply(self, d, alpha, lo, hi):
    return _native.scale_apply(self.ctx, d, alpha, lo, hi)
