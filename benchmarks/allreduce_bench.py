#!/usr/bin/env python
"""All-reduce of the CNN's 2.6 GB of fp32 gradients: one flat call vs 256 MB buckets (torchrun, one rank per GPU)."""
import os
import sys
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scale-equivariant-imaging_b200"))


def main():
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    n = 645_063_043
    flat = torch.randn(n, device=dev)
    chunks = list(flat.split(64 * 2 ** 20))

    def timed(fn, reps=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    one = timed(lambda: dist.all_reduce(flat, op=dist.ReduceOp.AVG))
    many = timed(lambda: [dist.all_reduce(c, op=dist.ReduceOp.AVG) for c in chunks])
    half = flat.bfloat16()
    bf = timed(lambda: dist.all_reduce(half, op=dist.ReduceOp.AVG))
    if dist.get_rank() == 0:
        w = dist.get_world_size()
        gb = n * 4 / 1e9
        print(f"world {w}: one flat fp32 call {one:.2f} ms ({gb * 2 * (w - 1) / w / one * 1e3:.0f} GB/s bus), "
              f"{len(chunks)} x 256 MB buckets {many:.2f} ms, one bf16 call {bf:.2f} ms")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
