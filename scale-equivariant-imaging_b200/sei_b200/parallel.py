"""Data parallelism for the training step: one process per GPU, replicated network and physics constants,
the batch split across ranks, and ONE exchange per step -- a bucketed all-reduce (average) of the parameter
gradients (NCCL over NVLink on the B200 box; gloo in the CPU unit tests).  This replaces the reference's
in-process torch.nn.DataParallel wrap of the network (src/models/__init__.py:142-145).

Every operator of the hot path acts per image, so no other collective exists: no halo exchange, no
all-to-all.  Caveats carried over from the reference's arithmetic: SURE's constant
sigma^2 / B sees the per-rank batch (a constant offset of the logged loss, not of the gradients), and equal
shard sizes are required for the mean of per-rank means to equal the global mean."""
import os

import torch
import torch.distributed as dist


def init_distributed(backend=None):
    """Initialise torch.distributed from the torchrun environment.  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def shard_batch(t, rank, world):
    """Rank's equal share of a global batch (dimension 0); the global batch must divide evenly."""
    n = t.shape[0]
    if n % world != 0:
        raise ValueError(f"global batch {n} does not split evenly over {world} ranks "
                         "(equal shards are required for the mean of means to equal the global mean)")
    per = n // world
    return t[rank * per:(rank + 1) * per]


def make_buckets(params, max_elems=64 * 2 ** 20):
    """Greedy partition of the parameter list, in order, into buckets of at most ~max_elems elements."""
    buckets, cur, n = [], [], 0
    for p in params:
        cur.append(p)
        n += p.numel()
        if n >= max_elems:
            buckets.append(cur)
            cur, n = [], 0
    if cur:
        buckets.append(cur)
    return buckets


def completion_groups(module, max_elems):
    """Partition a module tree into the sub-modules whose parameter gradients become final together: descend from the
    root until a sub-module holds at most `max_elems` parameters, has parameters of its own, or declares itself atomic
    (`_sei_atomic_group = True`: a module that runs as ONE autograd node, whose children are never called -- the
    restoration CNN's ConvBlock).  For the CNN at its default flags this yields the deepest ConvBlock (537 M parameters,
    83 % of the gradient bytes, final half-way through the last backward pass) as a bucket of its own and one bucket per
    shallower block / resampling layer."""
    out = []

    def walk(m):
        n = sum(p.numel() for p in m.parameters())
        if n == 0:
            return
        own = any(True for _ in m.parameters(recurse=False))
        if own or n <= max_elems or getattr(m, "_sei_atomic_group", False) or not any(True for _ in m.children()):
            out.append(m)
            return
        for c in m.children():
            walk(c)

    walk(module)
    return out


_ALIGN = 32          # elements: every gradient view starts on a 128-byte boundary of its flat bucket


class _GradReady(torch.autograd.Function):
    """Identity on a group's input.  Its backward runs once every gradient INSIDE the group has been produced for that
    pass of the network (they all feed the gradient of the group's input), which is when the bucket may be reduced."""

    @staticmethod
    def forward(ctx, x, reducer, index):
        ctx.reducer, ctx.index = reducer, index
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        ctx.reducer._group_backward_done(ctx.index)
        return g, None, None


class GradAllReducer:
    """Average the gradients of `params` over all ranks, one all-reduce per bucket, overlapped with the backward pass.

    The gradients live in one flat buffer per bucket (every `p.grad` is a 128-byte-aligned view into it, installed here
    and kept by in-place accumulation / `zero_grad(set_to_none=False)`), so a bucket is reduced in place with a single
    NCCL all-reduce (ReduceOp.AVG): no flatten / unflatten copies and no separate division kernel.  Gradients that were
    re-created elsewhere (e.g. `zero_grad(set_to_none=True)`) fall back to the copying path.

    Two ways to use it:
      * `reducer()` after `backward()`: every bucket is reduced then (nothing overlaps);
      * `GradAllReducer(params, module=net)`: buckets follow `completion_groups(net)`; call `reducer.arm()` before
        `loss.backward()` and `reducer.finish()` after it.  An identity node on each group's input counts the passes of
        the network that go through the group (a `proposed` step runs it three times); when the backward pass has
        produced the group's last contribution its bucket is all-reduced asynchronously (torch runs NCCL collectives on
        the process group's own stream) while the rest of the backward pass keeps the SMs busy.  `finish()` reduces
        what is left (groups whose input carries no gradient, such as the first layer) and joins the streams.
        Under CUDA-graph capture the hooks only record an event per finished group; `finish()` (still capturing, on the
        capturing thread) then enqueues the collectives on a communication stream that waits for those events.  A graph
        orders its nodes by dependencies, not by issue order, so on replay every all-reduce starts as soon as its group's
        gradients exist and runs beside the rest of the backward pass."""

    def __init__(self, params, max_elems=2 ** 40, group=None, module=None, group_max_elems=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.modules = []
        if module is not None:
            total = sum(p.numel() for p in self.params)
            self.modules = completion_groups(module, group_max_elems or max(total // 8, 1 << 20))
            seen, self.buckets = set(), []
            wanted = {id(p) for p in self.params}
            for m in self.modules:
                b = [p for p in m.parameters() if id(p) in wanted and id(p) not in seen]
                seen.update(id(p) for p in b)
                self.buckets.append(b)
            rest = [p for p in self.params if id(p) not in seen]
            if rest:
                self.buckets.append(rest)
        else:
            self.buckets = make_buckets(self.params, max_elems)
        self.flat = []
        self.offsets = []
        if self.world > 1:
            for bucket in self.buckets:
                same = len({(p.dtype, p.device) for p in bucket}) == 1
                if not same or not bucket:
                    self.flat.append(None)
                    self.offsets.append(None)
                    continue
                offs, off = [], 0
                for p in bucket:
                    offs.append(off)
                    off += -(-p.numel() // _ALIGN) * _ALIGN
                flat = torch.zeros(off, dtype=bucket[0].dtype, device=bucket[0].device)
                for p, o in zip(bucket, offs):
                    view = flat[o:o + p.numel()].view_as(p)
                    if p.grad is not None:
                        view.copy_(p.grad)
                    p.grad = view
                self.flat.append(flat)
                self.offsets.append(offs)
        # overlap bookkeeping
        self._armed = False
        self._fwd = [0] * len(self.buckets)
        self._bwd = [0] * len(self.buckets)
        self._done = [False] * len(self.buckets)
        self._works = []
        self._ready = []            # (bucket index, event) recorded by the hooks while a CUDA graph is being captured
        self._deferred = False
        self._comm_stream = None
        self._hooks = []
        if self.world > 1:
            for i, m in enumerate(self.modules):
                self._hooks.append(m.register_forward_pre_hook(self._make_pre_hook(i)))

    # ---- overlap protocol
    def _make_pre_hook(self, index):
        def hook(mod, args):
            if not args or not torch.is_tensor(args[0]) or not torch.is_grad_enabled() or not args[0].requires_grad:
                return None
            self._fwd[index] += 1
            return (_GradReady.apply(args[0], self, index),) + tuple(args[1:])
        return hook

    def _group_backward_done(self, index):
        self._bwd[index] += 1
        if self._armed and not self._done[index] and self._bwd[index] >= self._fwd[index]:
            if self._deferred:
                self._done[index] = True
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream())
                self._ready.append((index, ev))
            else:
                self._reduce_bucket(index, async_op=True)

    def arm(self):
        """call right before loss.backward(): the forward passes since the last finish() are what the backward covers"""
        self._armed = True
        self._bwd = [0] * len(self.buckets)
        self._done = [False] * len(self.buckets)
        self._ready = []
        self._deferred = bool(self.world > 1 and torch.cuda.is_available() and torch.cuda.is_current_stream_capturing())

    def finish(self):
        """call right after loss.backward(): reduce the buckets the backward pass did not release, wait for all"""
        if self.world > 1 and self._deferred:
            cur = torch.cuda.current_stream()
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream()
            comm = self._comm_stream
            for index, ev in self._ready:                      # buckets the backward pass released, in that order
                comm.wait_event(ev)
                with torch.cuda.stream(comm):
                    self._done[index] = False
                    self._reduce_bucket(index, async_op=False)
            comm.wait_stream(cur)                              # the rest needs the whole backward pass
            with torch.cuda.stream(comm):
                for i in range(len(self.buckets)):
                    if not self._done[i]:
                        self._reduce_bucket(i, async_op=False)
            cur.wait_stream(comm)
            self._ready = []
        elif self.world > 1:
            for i in range(len(self.buckets)):
                if not self._done[i]:
                    self._reduce_bucket(i, async_op=True)
            for w in self._works:
                w.wait()               # the compute stream waits for the collective's stream (no host block on NCCL)
        self._works = []
        self._armed = False
        self._fwd = [0] * len(self.buckets)

    # ---- one bucket
    def _in_place(self, index):
        flat, offs = self.flat[index], self.offsets[index]
        if flat is None:
            return False
        for p, o in zip(self.buckets[index], offs):
            if p.grad is None or p.grad.data_ptr() != flat.data_ptr() + o * flat.element_size():
                return False
        return True

    def _reduce_bucket(self, index, async_op):
        self._done[index] = True
        bucket = self.buckets[index]
        if not bucket:
            return
        avg = dist.get_backend(self.group) == "nccl"
        if self._in_place(index):
            flat = self.flat[index]
            if avg:
                w = dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group, async_op=async_op)
                if async_op:
                    self._works.append(w)
            else:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
                flat.div_(self.world)
            return
        grads = [p.grad for p in bucket if p.grad is not None]
        if not grads:
            return
        tmp = torch._utils._flatten_dense_tensors(grads)
        dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=self.group)
        tmp.div_(self.world)
        for g, f in zip(grads, torch._utils._unflatten_dense_tensors(tmp, grads)):
            g.copy_(f)

    def __call__(self):
        if self.world == 1:
            return
        self._done = [False] * len(self.buckets)
        self.finish()


def broadcast_parameters(module, src=0):
    """Make every rank start from rank `src`'s weights.  The parameters themselves are broadcast (not `.data`), so
    their version counters advance and low-precision weight caches keyed on them are rebuilt."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        with torch.no_grad():
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t, src=src)
                inval = getattr(t, "_sei_invalidate", None)
                if inval is not None:
                    inval()
                if getattr(t, "_sei_maintained", False):
                    t._sei_maintained = False
