"""Batch sharding over the GPUs of one node (BASELINE.json configs[4]).

The path shards over independent ciphertexts: every rank owns an engine replica (tables + keys) and computes a block
of the batch; nothing inside an HMult+Relin crosses GPUs.  What does cross GPUs is the batch itself when it does not
already live where it is computed.  This module is that data plane:

  shard_range     contiguous blocks of ceil(B / G) units per rank
  ExchangePlan    who stores which unit ("home"), who computes it, cut into pipeline ticks of `chunk` units
                    rooted(...)  the whole batch lives in the root's HBM (scatter inputs / gather results)
                    spread(...)  the batch lives evenly on all ranks, but in the producer's partition: every rank
                                 computes a 1/G sub-slice of every home's slice (all-to-all repartition and back)
                    local(...)   every rank stores exactly what it computes (no exchange)
  Exchange        the double-buffered pipeline: at tick t a rank receives the inputs of its chunk t and returns the
                  results of chunk t-2 (one grouped send/recv over torch.distributed: NCCL over NVLink on the GPUs,
                  gloo in the CPU tests) while it computes chunk t-1.  Units that are already local are computed in
                  place, never copied.

torch.distributed carries the point-to-point groups, the barrier and the max-over-ranks of the device time; the
arithmetic is the caller's `compute` callback (bench.py: pfhe_multiply_and_relin_batch on pointer arrays).
"""
from dataclasses import dataclass, field


def shard_range(total, rank, world):
    """[begin, end) of the ciphertext indices rank `rank` owns."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("rank/world is invalid")
    per = -(-total // world)
    begin = min(total, rank * per)
    return begin, min(total, begin + per)


def max_over_ranks(value, dist=None, device=None):
    """max of a python float over all ranks (device time of the slowest rank)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class _Mapped:
    """device memory of another rank mapped into this process (pfhe_ipc_open); exposes __cuda_array_interface__"""

    def __init__(self, ptr, offset, shape):
        self.ptr, self.offset = ptr, offset
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<i8", "data": (ptr, False), "version": 2}

    def close(self):
        from ._lib import lib
        if self.ptr:
            lib.pfhe_ipc_close(self.ptr, self.offset)
            self.ptr = 0


def peer_view(tensor, src, rank, dist, device):
    """A tensor on every rank that aliases rank `src`'s contiguous int64 `tensor` (CUDA IPC mapping of the allocation,
    opened with the importing rank's own device current so that its kernels address the owner's HBM over NVLink):
    kernels of the other ranks read their operands from, and write their results to, the owner's memory directly --
    no staging copy, no communication kernel.  `tensor` is only looked at on `src`.  Returns (view, mapping); keep the
    mapping alive while the view is in use and close() it afterwards.  The caller orders accesses across ranks
    (barrier + synchronize around a pass)."""
    import ctypes
    import torch
    from ._lib import lib, check
    box = [None]
    if rank == src:
        if tensor.dtype != torch.int64 or not tensor.is_contiguous():
            raise ValueError("peer_view exports contiguous int64 tensors")
        handle = ctypes.create_string_buffer(64)
        offset = ctypes.c_uint64()
        check(lib.pfhe_ipc_export(ctypes.c_void_p(tensor.data_ptr()), handle, ctypes.byref(offset)))
        box[0] = (handle.raw, int(offset.value), tuple(tensor.shape))
    dist.broadcast_object_list(box, src)
    if rank == src:
        return tensor, None
    raw, offset, shape = box[0]
    mapped = ctypes.c_void_p()
    with torch.cuda.device(device):
        check(lib.pfhe_ipc_open(raw, offset, ctypes.byref(mapped)))
        m = _Mapped(mapped.value, offset, shape)
        view = torch.as_tensor(m, device=device)
    return view, m


@dataclass
class Piece:
    """a run of consecutive units [lo, hi) (global ids) that live on rank `home` and are computed by one rank in one tick"""
    home: int
    lo: int
    hi: int
    off: int = 0   # first unit's position inside the computing rank's staging slot (remote pieces only)

    def __len__(self):
        return self.hi - self.lo


@dataclass
class ExchangePlan:
    total: int
    world: int
    chunk: int
    kind: str
    slot: int = 0                               # units a staging slot holds (>= the largest tick)
    home: list = field(default_factory=list)    # per rank: (lo, hi) of the units it stores
    ticks: list = field(default_factory=list)   # per rank: list of ticks, a tick = list of Piece

    @staticmethod
    def _finish(plan):
        plan.slot = max(plan.slot, plan.chunk)
        for per_rank in plan.ticks:
            for c, tick in enumerate(per_rank):
                off = 0
                for p in tick:
                    p.off = off
                    off += len(p)
                if off > plan.slot:
                    raise AssertionError("tick larger than the staging slot")
        return plan

    @staticmethod
    def _check(total, world, chunk):
        if total < 0 or world < 1 or chunk < 1:
            raise ValueError("total / world / chunk is invalid")

    @classmethod
    def rooted(cls, total, world, chunk, root=0):
        """the batch lives on `root`; rank r computes shard_range(total, r, world)"""
        cls._check(total, world, chunk)
        if not (0 <= root < world):
            raise ValueError("root is invalid")
        plan = cls(total, world, chunk, "rooted")
        plan.home = [(0, total) if r == root else (0, 0) for r in range(world)]
        for r in range(world):
            lo, hi = shard_range(total, r, world)
            plan.ticks.append([[Piece(root, b, min(hi, b + chunk))] for b in range(lo, hi, chunk)])
        return cls._finish(plan)

    @classmethod
    def local(cls, total, world, chunk):
        """rank r stores and computes shard_range(total, r, world)"""
        cls._check(total, world, chunk)
        plan = cls(total, world, chunk, "local")
        plan.home = [shard_range(total, r, world) for r in range(world)]
        for r in range(world):
            lo, hi = plan.home[r]
            plan.ticks.append([[Piece(r, b, min(hi, b + chunk))] for b in range(lo, hi, chunk)])
        return cls._finish(plan)

    @classmethod
    def spread(cls, total, world, chunk):
        """rank h stores shard_range(total, h, world); rank c computes the c-th 1/world sub-slice of every home's slice"""
        cls._check(total, world, chunk)
        plan = cls(total, world, chunk, "spread")
        plan.home = [shard_range(total, r, world) for r in range(world)]
        per_home = max(1, chunk // world)   # units a tick takes from each home
        plan.slot = per_home * world
        for c in range(world):
            subs = []
            for h in range(world):
                lo, hi = plan.home[h]
                a, b = shard_range(hi - lo, c, world)
                subs.append((lo + a, lo + b))
            n_ticks = max((-(-(b - a) // per_home) for a, b in subs), default=0)
            ticks = []
            for t in range(n_ticks):
                tick = []
                for h, (a, b) in enumerate(subs):
                    s, e = a + t * per_home, min(b, a + (t + 1) * per_home)
                    if s < e:
                        tick.append(Piece(h, s, e))
                ticks.append(tick)
            plan.ticks.append(ticks)
        return cls._finish(plan)

    def n_ticks(self):
        return max((len(t) for t in self.ticks), default=0)

    def computed_by(self, rank):
        return sum(len(p) for tick in self.ticks[rank] for p in tick)

    def bytes_moved(self, rank, in_bytes, out_bytes):
        """(bytes this rank sends, bytes it receives) over the interconnect in one pass"""
        sent = recv = 0
        for c in range(self.world):
            for tick in self.ticks[c]:
                for p in tick:
                    if p.home == c:
                        continue
                    if c == rank:
                        recv += len(p) * in_bytes
                        sent += len(p) * out_bytes
                    if p.home == rank:
                        sent += len(p) * in_bytes
                        recv += len(p) * out_bytes
        return sent, recv


class Exchange:
    """Runs an ExchangePlan.  `store_in` / `store_out`: this rank's home storage, tensors [n_home, in_words] and
    [n_home, out_words] (n_home = hi - lo of plan.home[rank]; may be empty).  Staging for remote units is allocated here."""

    def __init__(self, plan, rank, store_in, store_out, dist=None, group=None, depth=2):
        import torch
        self.torch = torch
        self.plan, self.rank, self.dist, self.group = plan, rank, dist, group
        self.store_in, self.store_out = store_in, store_out
        lo, hi = plan.home[rank]
        if store_in.shape[0] != hi - lo or store_out.shape[0] != hi - lo:
            raise ValueError("home storage does not match the plan")
        self.in_words, self.out_words = store_in.shape[1], store_out.shape[1]
        remote = any(p.home != rank for tick in plan.ticks[rank] for p in tick)
        self.depth = depth
        slots = depth if remote else 0
        self.stage_in = [store_in.new_empty((plan.slot, self.in_words)) for _ in range(slots)]
        self.stage_out = [store_out.new_empty((plan.slot, self.out_words)) for _ in range(slots)]
        if plan.world > 1 and dist is None:
            raise ValueError("a multi-rank plan needs torch.distributed")
        self._views = {}

    # -- views -----------------------------------------------------------------------------------------------
    def tick_views(self, t):
        """[(in_view [k, in_words], out_view [k, out_words])] of the units this rank computes at tick t, in order"""
        if t in self._views:
            return self._views[t]
        mine = self.plan.ticks[self.rank]
        views = []
        if 0 <= t < len(mine):
            lo = self.plan.home[self.rank][0]
            for p in mine[t]:
                if p.home == self.rank:
                    views.append((self.store_in[p.lo - lo:p.hi - lo], self.store_out[p.lo - lo:p.hi - lo]))
                else:
                    views.append((self.stage_in[t % self.depth][p.off:p.off + len(p)], self.stage_out[t % self.depth][p.off:p.off + len(p)]))
        self._views[t] = views
        return views

    def _needs_wait(self, t):
        """does compute(t) depend on the group posted at tick t (received inputs, or a staging slot it is about to reuse)"""
        mine = self.plan.ticks[self.rank]
        for u in (t, t - 2):
            if 0 <= u < len(mine) and any(p.home != self.rank for p in mine[u]):
                return True
        return False

    # -- one tick's point-to-point group -------------------------------------------------------------------------
    def _post(self, t):
        plan, me, dist = self.plan, self.rank, self.dist
        if plan.world == 1:
            return []
        ops = []
        lo = plan.home[me][0]
        P2P = dist.P2POp
        # canonical order between any two ranks: inputs of tick t (by computing rank, then piece), then results of tick t-2
        for c in range(plan.world):
            tick = plan.ticks[c][t] if 0 <= t < len(plan.ticks[c]) else []
            for p in tick:
                if p.home == c:
                    continue
                if c == me:
                    ops.append(P2P(dist.irecv, self.stage_in[t % 2][p.off:p.off + len(p)], p.home, self.group))
                elif p.home == me:
                    ops.append(P2P(dist.isend, self.store_in[p.lo - lo:p.hi - lo], c, self.group))
        u = t - 2
        for c in range(plan.world):
            tick = plan.ticks[c][u] if 0 <= u < len(plan.ticks[c]) else []
            for p in tick:
                if p.home == c:
                    continue
                if c == me:
                    ops.append(P2P(dist.isend, self.stage_out[u % 2][p.off:p.off + len(p)], p.home, self.group))
                elif p.home == me:
                    ops.append(P2P(dist.irecv, self.store_out[p.lo - lo:p.hi - lo], c, self.group))
        if not ops:
            return []
        return dist.batch_isend_irecv(ops)

    def run(self, compute):
        """One pass over the batch.  compute(t, views) enqueues the arithmetic of tick t on the current stream
        (views = tick_views(t)).  Returns when everything is enqueued (CUDA) or done (CPU); the caller synchronises."""
        T = self.plan.n_ticks()
        pending = {}
        for t in range(T + 2):
            works = self._post(t)
            if works:
                pending[t] = works
            u = t - 1
            if 0 <= u < len(self.plan.ticks[self.rank]):
                if self._needs_wait(u):
                    for k in (u, u - 1):   # group u brought the inputs; group u - 1 ... u carried the sends of slot u % 2
                        for w in pending.pop(k, []):
                            w.wait()
                compute(u, self.tick_views(u))
        for works in pending.values():
            for w in works:
                w.wait()


class PullExchange(Exchange):
    """One-sided form of the same pipeline for ranks that can address each other's home storage (peer_view mappings of
    one NVLink node): the computing rank PULLS the operands of tick t+1 into its staging slot and PUSHES the results of
    tick t-1 back with plain asynchronous copies (copy engines: no communication kernel competes with the arithmetic
    for SMs, no rendezvous with the owner) while it computes tick t.  homes_in / homes_out: per rank, that rank's home
    storage as addressable from here (own tensors for `rank`, mappings for the others; entries of ranks this rank never
    touches may be None).  The caller brackets a pass with a barrier on both sides (operands must be final before, results
    are visible to their owners after)."""

    def __init__(self, plan, rank, homes_in, homes_out, depth=2):
        super().__init__(plan, rank, homes_in[rank], homes_out[rank], dist=_NoDist if plan.world > 1 else None, depth=depth)
        self.homes_in, self.homes_out = homes_in, homes_out
        t = self.torch
        self.cuda = homes_in[rank].is_cuda
        if self.cuda and self.stage_in:
            self.s_in, self.s_out = t.cuda.Stream(), t.cuda.Stream()
            self.ev_in = [t.cuda.Event() for _ in range(depth)]
            self.ev_comp = [t.cuda.Event() for _ in range(depth)]
            self.ev_out = [t.cuda.Event() for _ in range(depth)]

    def run(self, compute):
        t = self.torch
        plan, me = self.plan, self.rank
        mine = plan.ticks[me]
        staged = self.cuda and bool(self.stage_in)
        if staged:
            cur = t.cuda.current_stream()
            self.s_in.wait_stream(cur)
            self.s_out.wait_stream(cur)
        used = set()
        for u, tick in enumerate(mine):
            slot = u % self.depth
            remote = [p for p in tick if p.home != me]
            if remote:
                if staged:
                    if slot in used:
                        self.s_in.wait_event(self.ev_comp[slot])      # compute(u - 2) has read this slot
                    with t.cuda.stream(self.s_in):
                        for p in remote:
                            lo = plan.home[p.home][0]
                            self.stage_in[slot][p.off:p.off + len(p)].copy_(self.homes_in[p.home][p.lo - lo:p.hi - lo],
                                                                          non_blocking=True)
                        self.ev_in[slot].record(self.s_in)
                    cur.wait_event(self.ev_in[slot])
                    if slot in used:
                        cur.wait_event(self.ev_out[slot])             # results of tick u - 2 have left this slot
                else:
                    for p in remote:
                        lo = plan.home[p.home][0]
                        self.stage_in[slot][p.off:p.off + len(p)].copy_(self.homes_in[p.home][p.lo - lo:p.hi - lo])
            compute(u, self.tick_views(u))
            if remote:
                if staged:
                    self.ev_comp[slot].record(cur)
                    self.s_out.wait_event(self.ev_comp[slot])
                    with t.cuda.stream(self.s_out):
                        for p in remote:
                            lo = plan.home[p.home][0]
                            self.homes_out[p.home][p.lo - lo:p.hi - lo].copy_(self.stage_out[slot][p.off:p.off + len(p)],
                                                                              non_blocking=True)
                        self.ev_out[slot].record(self.s_out)
                    used.add(slot)
                else:
                    for p in remote:
                        lo = plan.home[p.home][0]
                        self.homes_out[p.home][p.lo - lo:p.hi - lo].copy_(self.stage_out[slot][p.off:p.off + len(p)])
        if staged:
            cur.wait_stream(self.s_out)
            cur.wait_stream(self.s_in)


class _NoDist:
    """placeholder: PullExchange never posts a message"""
