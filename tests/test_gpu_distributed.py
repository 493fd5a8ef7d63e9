"""Data parallelism on real GPUs (needs >= 2, skipped otherwise): two ranks over NCCL, the restoration CNN's kernels, the
bucketed all-reduce launched from inside backward() (sei_b200.parallel.GradAllReducer with module=...).  The averaged
in-place bucket gradients must equal the gradients of the full batch computed by one process, and the buckets released
during the backward pass must be the ones whose groups had finished."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(os.path.dirname(here), "scale-equivariant-imaging_b200"))
    import torch.distributed as dist
    import models.convolutional as mc
    from sei_b200 import parallel
    parallel.init_distributed(backend="nccl")
    dev = torch.device("cuda", rank)
    torch.manual_seed(5)
    kw = dict(in_channels=3, upsampling_rate=1, residual=True, inner_residual=True, num_conv_blocks=1, hidden_channels=16,
              inout_convs=True, scales=3)
    net = mc.ConvolutionalModel(**kw).to(dev)
    parallel.broadcast_parameters(net, src=0)
    torch.manual_seed(9)                                    # the same global batch on every rank
    xg, tg = torch.rand(8, 3, 64, 64, device=dev), torch.rand(8, 3, 64, 64, device=dev)
    x, t = parallel.shard_batch(xg, rank, world), parallel.shard_batch(tg, rank, world)

    def loss_of(model, a, b):                               # three passes through the network, like a `proposed` step
        return sum(torch.nn.functional.mse_loss(model(a + 0.05 * k), b) for k in range(3))

    red = parallel.GradAllReducer(net.parameters(), module=net, group_max_elems=20000)
    assert all(p.grad.data_ptr() % 16 == 0 for p in net.parameters())
    fired = []
    orig = red._reduce_bucket
    red._reduce_bucket = lambda i, async_op: (fired.append(red._armed), orig(i, async_op))[1]
    for p in net.parameters():
        p.grad.zero_()
    loss = loss_of(net, x, t)
    red.arm()
    loss.backward()
    n_bwd = len(fired)
    red.finish()
    torch.cuda.synchronize()
    ok = True
    worst = 0.0
    if rank == 0:
        ref = mc.ConvolutionalModel(**kw).to(dev)
        ref.load_state_dict(net.state_dict())
        loss_of(ref, xg, tg).backward()
        for (k, p), q in zip(net.named_parameters(), ref.parameters()):
            a, b = p.grad.double().flatten(), q.grad.double().flatten()
            if a.numel() > 16 and float(b.norm()) > 1e-12:
                cos = float(a @ b / (a.norm() * b.norm() + 1e-300))
                worst = max(worst, 1 - cos)
                ok = ok and cos > 0.999 and abs(float(a.norm() / b.norm()) - 1) < 0.02
    results[rank] = (ok, worst, n_bwd, len(fired), len(red.buckets))
    dist.barrier()
    dist.destroy_process_group()


def test_nccl_overlapped_gradient_average_equals_full_batch():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
        ok, worst, n_bwd, n_all, n_buckets = results[0]
        assert ok, worst
        assert n_all == n_buckets and 1 <= n_bwd < n_buckets
        print(f"NCCL DP gradients vs full batch: worst 1 - cos = {worst:.2e}; {n_bwd} of {n_buckets} buckets reduced inside backward()")
