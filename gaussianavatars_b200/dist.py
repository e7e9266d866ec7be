"""Frame-sharded data parallelism: every rank holds the full splat model, renders its own cameras,
and the per-splat parameter gradients are summed with ONE NCCL all-reduce over NVLink/NVSwitch.

The reference has no multi-GPU code at all; this is the new capability north_star asks for.  No
collective touches the data path of a frame: frames are independent given the parameters.  The fused backward
already writes all six parameter gradients into one flat fp32 buffer (59 floats/splat at SH3 = 236 B/splat), which
is the communication buffer itself -- no pack/unpack pass precedes the collective.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence

import torch
import torch.distributed as dist


def shard_frames(num_frames: int, rank: int, world_size: int) -> List[int]:
    """Camera indices rendered by `rank`: round-robin r, r+W, r+2W, ... (reference: one camera per step, train.py:113)."""
    return list(range(rank, num_frames, world_size))


def shard_frames_by_cost(costs: Sequence[float], rank: int, world_size: int) -> List[int]:
    """Cost-aware sharding: cameras sorted by a cost estimate (the instance count of a previous visit), then dealt out
    so that the cameras rendered in the SAME step -- positions j*W .. j*W + W-1 of the sorted list -- have neighbouring
    costs.  A synchronous data-parallel step lasts as long as its slowest rank; with frames of unequal cost, dealing
    neighbours to one step makes every step cost about its group's mean instead of the maximum over a random draw
    (the sequence-length bucketing of other domains).  Every camera still goes to exactly one rank."""
    order = sorted(range(len(costs)), key=lambda i: (costs[i], i))
    return order[rank::world_size]


def _flat_covers(flat: torch.Tensor, grads: Sequence[torch.Tensor]) -> bool:
    if flat is None:
        return False
    lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * flat.element_size()
    total = 0
    for g in grads:
        if g is None or not g.is_contiguous() or g.dtype != flat.dtype:
            return False
        if not (lo <= g.data_ptr() and g.data_ptr() + g.numel() * g.element_size() <= hi):
            return False
        total += g.numel()
    return total == flat.numel()


def allreduce_splat_grads(pc, params: Iterable[torch.Tensor] = None, group=None, average: bool = False) -> int:
    """Sum (or average) the splat-parameter gradients over ranks.  Returns the number of collectives issued
    (1 when the gradients still alias the fused backward's flat buffer, else one coalesced fallback)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    params = list(pc.parameters() if params is None else params)
    grads = [p.grad for p in params]
    flat = getattr(pc, "flat_grad", None)
    world = dist.get_world_size(group)
    if _flat_covers(flat, grads):
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat.div_(world)
        return 1
    live = [g for g in grads if g is not None]
    buf = torch.cat([g.reshape(-1) for g in live])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    if average:
        buf.div_(world)
    off = 0
    for g in live:
        g.copy_(buf[off:off + g.numel()].view_as(g))
        off += g.numel()
    return 1


def allreduce_densification_stats(xyz_gradient_accum, denom, max_radii2D, group=None):
    """The three per-splat statistics the densifier consumes (scene/gaussian_model.py:517-519, train.py:197):
    sum, sum, max."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    dist.all_reduce(xyz_gradient_accum, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(denom, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(max_radii2D, op=dist.ReduceOp.MAX, group=group)


class SymmetricGradBuffer:
    """The flat per-splat gradient buffer as NVLink SYMMETRIC memory with an NVLS multicast mapping.

    With it attached to the model (`pc.symm_grad = SymmetricGradBuffer(pc, group)`), the fused backward does not
    store its parameter gradients locally and all-reduce them afterwards: preprocess-backward issues
    `multimem.red.add.f32` to the multicast address, so NVSwitch sums the ranks' contributions into EVERY rank's
    replica while the kernel is still running (include/gab200_rasterizer.h `grads_are_multicast`).  Protocol per step:

        buf.begin()      # switch to the pre-zeroed replica; zero the other one for the next step
        loss.backward()  # every rank's kernel reduces into all replicas
        buf.end()        # device-side group barrier: every replica now holds the sum

    Both barriers are stream-ordered device operations (symmetric-memory signal pads); the host never blocks.
    Falls back (attribute `.enabled` False) when the fabric has no multicast support; callers then use
    `allreduce_splat_grads` (one NCCL all-reduce)."""

    def __init__(self, pc, group=None, mode: str = "push"):
        """mode "push": the backward kernel emits multimem.red into every replica (above).
        mode "two_shot": the backward stores its gradients into the LOCAL replica with plain stores; `end()` then runs
        the two-shot NVLS all-reduce of csrc/nvls.cu in place (multimem.ld_reduce of this rank's slice + multimem.st
        to every replica) between two device-side group barriers.  (N-1)/N of the buffer crosses NVLink each way per
        GPU instead of N x: this is the form that scales, and it is what bench.py uses for N > 1 when the fabric has
        multicast.
        mode "plain": an ordinary device buffer, reduced by ncclAllReduce -- the same caller-owned-buffer protocol
        (begin / end / reduce) on a fabric without multicast.

        In the "two_shot" and "plain" modes the buffer is PERSISTENT and caller-owned, which is what lets the reduction
        be deferred: `end(reduce=False)` adopts the local gradients, and `reduce()` can be issued later -- on another
        stream, or as a forked branch of the NEXT step's CUDA graph (graph.GraphedFrame side_work), where it runs under
        that step's forward + backward."""
        import torch.distributed._symmetric_memory as symm_mem

        if mode not in ("push", "two_shot", "plain"):
            raise ValueError("mode must be 'push', 'two_shot' or 'plain'")
        self.enabled = False
        self.pc = pc
        self.mode = mode
        self.collective = "nvls" if mode == "two_shot" else "nccl"   # what reduce() issues (two_shot may use either)
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        group = group if group is not None else dist.group.WORLD
        params = list(pc.parameters())
        self.numel = sum(p.numel() for p in params)
        device = params[0].device
        if mode == "plain":
            self.flats = [torch.zeros(self.numel, dtype=torch.float32, device=device)]
            self.handles, self.mc_ptrs = [None], [0]
        elif not self._alloc_symmetric(symm_mem, device, group):
            return
        self.enabled = True
        self.group = group
        self.params = params
        # the layout of the fused backward's flat buffer: parameters in pc.parameters() order, each contiguous
        self.all_views = []
        for f in self.flats:
            views, off = [], 0
            for p in params:
                views.append(f[off:off + p.numel()].view_as(p))
                off += p.numel()
            self.all_views.append(views)
        for f in self.flats:
            f.zero_()
        if self.handles[0] is not None:
            self.handles[0].barrier(channel=0)
        self.cur = 1 if mode == "push" else 0  # push: begin() flips first
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def _alloc_symmetric(self, symm_mem, device, group) -> bool:
        try:
            try:
                symm_mem.set_backend("CUDA")
            except Exception:
                pass
            # push: two replicas used alternately: the one for step i+1 is zeroed during step i, so that only ONE group
            # barrier per step (end) sits on the critical path -- see begin()
            self.flats = [symm_mem.empty(self.numel, dtype=torch.float32, device=device)
                          for _ in range(2 if self.mode == "push" else 1)]
            self.handles = [symm_mem.rendezvous(f, group) for f in self.flats]
            self.mc_ptrs = [int(h.multicast_ptr) for h in self.handles]
        except Exception as e:  # pragma: no cover - fabric / build dependent
            self.error = repr(e)
            return False
        if 0 in self.mc_ptrs:
            self.error = "no NVLS multicast support on this fabric"
            return False
        return True

    @property
    def flat(self):
        return self.flats[self.cur]

    @property
    def mc_ptr(self):
        return self.mc_ptrs[self.cur]

    def matches(self, pc=None) -> bool:
        """True while the model still has the parameter OBJECTS and sizes this buffer was laid out for (densification
        and pruning replace the nn.Parameters, scene/gaussian_model.py:334-419: call `rebuild()` after them)."""
        params = list((pc if pc is not None else self.pc).parameters())
        return (len(params) == len(self.params) and all(a is b for a, b in zip(params, self.params))
                and sum(p.numel() for p in params) == self.numel)

    def rebuild(self, group=None):
        """New symmetric buffers for the model's current parameters (collective: every rank must call it)."""
        self.__init__(self.pc, group if group is not None else getattr(self, "group", None), mode=self.mode)
        if self.enabled:
            self.pc.symm_grad = self
        return self

    def _two_shot(self):
        import ctypes as C

        from . import _native as N

        h = self.handles[0]
        h.barrier(channel=0)     # every rank's backward has written its replica
        with torch.cuda.device(self.flats[0].device):
            stream = torch.cuda.current_stream(self.flats[0].device).cuda_stream
            N.check(N.lib().gab200_nvls_allreduce(C.c_void_p(self.mc_ptrs[0]), self.numel, self.rank, self.world,
                                                  C.c_void_p(stream)), "gab200_nvls_allreduce")
        h.barrier(channel=1)     # every slice has been stored into every replica

    def reduce(self):
        """Sum the buffer over the ranks, in place, on the current stream ("two_shot" / "plain" modes).  Stream-ordered,
        no host wait; capturable.  The caller orders it after the backward that filled the buffer."""
        if self.mode == "push":
            raise RuntimeError("push mode reduces inside the backward kernel")
        if self.mode == "two_shot" and self.collective == "nvls":
            self._two_shot()
        else:
            dist.all_reduce(self.flats[0], op=dist.ReduceOp.SUM, group=self.group)

    def self_test(self) -> bool:
        """Collective (every rank calls it).  Reduces a known pattern through the two-shot kernel and checks the sums
        on this rank; leaves the buffer zeroed.  bench.py falls back to NCCL when any rank reports False."""
        if not self.enabled or self.mode != "two_shot":
            return False
        f = self.flats[0]
        idx = torch.arange(self.numel, device=f.device, dtype=torch.float32)
        f.copy_((idx % 13.0) * float(self.rank + 1))
        self._two_shot()
        want = (idx % 13.0) * float(self.world * (self.world + 1) // 2)
        ok = bool(torch.equal(f, want))
        self.handles[0].barrier(channel=0)   # nobody zeroes before everybody has compared
        f.zero_()
        self.handles[0].barrier(channel=1)
        return ok

    def begin(self):
        """Call before backward.  Switches to the replica that every rank zeroed before the previous step's end()
        barrier, and zeroes the other one (whose sums the optimizer has consumed by now) for the step after."""
        self.pc._gab200_mc_used = False  # the backward sets it when its gradients really went into this buffer
        if self.mode != "push":
            return                       # the backward overwrites every element of the local buffer: nothing to zero
        prev = self.cur
        self.cur ^= 1
        self.flats[prev].zero_()

    def end(self, reduce: bool = True):
        """Call after backward.  reduce=False ("two_shot" / "plain"): adopt the LOCAL gradients only -- the caller
        issues `reduce()` later (deferred reduction).  Group barrier; then the parameters' .grad are pointed at the reduced replica -- but
        only if this step's backward took the multicast path for exactly these parameters.  Otherwise (override_color,
        a model whose parameters were replaced or resized since the buffer was built) the locally stored gradients are
        the valid ones: they are summed with one NCCL all-reduce instead, and the zeroed replica is left alone."""
        took = getattr(self.pc, "_gab200_mc_used", False) and self.matches()
        if self.mode != "push":
            if took and reduce:
                self.reduce()
        else:
            self.handles[self.cur].barrier(channel=1)
        if took:
            # autograd may have CLONED the gradient views while the reduction was still in flight (it only adopts a
            # tensor it holds the sole reference to): point .grad at the reduced buffer itself.  With
            # zero_grad(set_to_none=False) autograd would accumulate INTO a replica that must read zero at its next
            # turn: these views are replaced every step, and begin() re-zeroes the replica that is about to rest.
            for p, v in zip(self.params, self.all_views[self.cur]):
                p.grad = v
            return True
        if reduce:
            allreduce_splat_grads(self.pc, group=self.group)
        return False
