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


class GradAllReducer:
    """Average the gradients of `params` over all ranks, one all-reduce per bucket.

    The gradients live in one flat buffer per bucket (every `p.grad` is a view into it, installed here and kept by
    in-place accumulation / `zero_grad(set_to_none=False)`), so a bucket is reduced in place with a single NCCL
    all-reduce (ReduceOp.AVG): no flatten / unflatten copies and no separate division kernel.  Gradients that were
    re-created elsewhere (e.g. `zero_grad(set_to_none=True)`) fall back to the copying path."""

    def __init__(self, params, max_elems=2 ** 40, group=None):
        # one bucket by default: nothing overlaps the reduction (it runs between backward and the optimizer step), so a
        # single large all-reduce gets the best NVLink bandwidth; pass a smaller max_elems to split it
        self.params = [p for p in params if p.requires_grad]
        self.buckets = make_buckets(self.params, max_elems)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.flat = []
        if self.world > 1:
            for bucket in self.buckets:
                same = len({(p.dtype, p.device) for p in bucket}) == 1
                if not same:
                    self.flat.append(None)
                    continue
                flat = torch.zeros(sum(p.numel() for p in bucket), dtype=bucket[0].dtype, device=bucket[0].device)
                off = 0
                for p in bucket:
                    view = flat[off:off + p.numel()].view_as(p)
                    if p.grad is not None:
                        view.copy_(p.grad)
                    p.grad = view
                    off += p.numel()
                self.flat.append(flat)

    def _in_place(self, bucket, flat):
        if flat is None:
            return False
        off = 0
        for p in bucket:
            if p.grad is None or p.grad.data_ptr() != flat.data_ptr() + off * flat.element_size():
                return False
            off += p.numel()
        return True

    def __call__(self):
        if self.world == 1:
            return
        avg = dist.get_backend(self.group) == "nccl"
        for bucket, flat in zip(self.buckets, self.flat):
            if self._in_place(bucket, flat):
                if avg:
                    dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
                else:
                    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
                    flat.div_(self.world)
                continue
            grads = [p.grad for p in bucket if p.grad is not None]
            if not grads:
                continue
            tmp = torch._utils._flatten_dense_tensors(grads)
            dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=self.group)
            tmp.div_(self.world)
            for g, f in zip(grads, torch._utils._unflatten_dense_tensors(tmp, grads)):
                g.copy_(f)


def broadcast_parameters(module, src=0):
    """Make every rank start from rank `src`'s weights."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src)
