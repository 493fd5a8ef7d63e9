"""The data-parallel host logic on the CPU with the gloo backend, world_size 2 (no GPU needed): bucketed
gradient averaging equals the full-batch gradient, shards are equal and disjoint, weights are broadcast."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


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
    from sei_b200 import parallel
    r, w, _ = parallel.init_distributed(backend="gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(100 + rank)                       # different initial weights per rank on purpose
    model = torch.nn.Sequential(torch.nn.Conv2d(3, 5, 3, padding=1), torch.nn.GELU(), torch.nn.Conv2d(5, 3, 1))
    parallel.broadcast_parameters(model, src=0)
    torch.manual_seed(7)                                # the same global batch on every rank
    x_global, t_global = torch.rand(8, 3, 12, 12), torch.rand(8, 3, 12, 12)
    x, t = parallel.shard_batch(x_global, rank, world), parallel.shard_batch(t_global, rank, world)
    assert x.shape[0] == 8 // world
    loss = torch.nn.functional.mse_loss(model(x), t)    # a per-rank mean, like every term of the SEI losses
    loss.backward()
    reducer = parallel.GradAllReducer(model.parameters(), max_elems=64)   # several buckets
    assert len(reducer.buckets) >= 2
    reducer()
    # single-process reference on the full batch with the same (rank-0) weights
    ref = torch.nn.Sequential(torch.nn.Conv2d(3, 5, 3, padding=1), torch.nn.GELU(), torch.nn.Conv2d(5, 3, 1))
    ref.load_state_dict(model.state_dict())
    torch.nn.functional.mse_loss(ref(x_global), t_global).backward()
    err = max(float((p.grad - q.grad).abs().max()) for p, q in zip(model.parameters(), ref.parameters()))
    sd = torch.cat([p.detach().flatten() for p in model.parameters()])
    gathered = [torch.zeros_like(sd) for _ in range(world)]
    dist.all_gather(gathered, sd)
    same_weights = all(torch.equal(gathered[0], g) for g in gathered)
    with pytest.raises(ValueError):
        parallel.shard_batch(torch.zeros(7, 1), rank, world)

    # the overlapped protocol: buckets follow the module tree, each is reduced from inside backward() as soon as the
    # last of the step's passes through its group has produced its gradients (a `proposed` step runs the network three
    # times); what the backward pass does not release (the first layer: its input carries no gradient) goes in finish()
    def tiny():
        return torch.nn.Sequential(torch.nn.Conv2d(3, 7, 3, padding=1), torch.nn.GELU(),
                                   torch.nn.Sequential(torch.nn.Conv2d(7, 7, 1), torch.nn.GELU(), torch.nn.Conv2d(7, 5, 1)),
                                   torch.nn.Conv2d(5, 3, 1))
    torch.manual_seed(3)
    net = tiny()
    red = parallel.GradAllReducer(net.parameters(), module=net, group_max_elems=100)
    assert len(red.buckets) >= 3
    # odd-sized tensors (9-element kernels, 3- and 7-element biases) must not misalign the views that follow them:
    # sei_adam_step_f32 and the in-place weight-gradient GEMM require 16-byte aligned gradients
    assert all(p.grad.data_ptr() % 16 == 0 for p in net.parameters())
    fired = []
    orig = red._reduce_bucket
    red._reduce_bucket = lambda i, async_op: (fired.append((i, red._armed)), orig(i, async_op))[1]
    for p in net.parameters():
        p.grad.zero_()
    loss2 = sum(torch.nn.functional.mse_loss(net(x + 0.1 * k), t) for k in range(3))
    red.arm()
    loss2.backward()
    n_from_backward = len(fired)
    red.finish()
    ref2 = tiny()
    ref2.load_state_dict(net.state_dict())
    sum(torch.nn.functional.mse_loss(ref2(x_global + 0.1 * k), t_global) for k in range(3)).backward()
    err2 = max(float((p.grad - q.grad).abs().max()) for p, q in zip(net.parameters(), ref2.parameters()))
    results[rank] = (err, same_weights, err2, n_from_backward, len(fired), len(red.buckets))
    dist.destroy_process_group()


def test_data_parallel_gradient_averaging_gloo():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
        assert len(results) == world
        for rank in range(world):
            err, same, err2, n_bwd, n_all, n_buckets = results[rank]
            assert err < 1e-6, err
            assert same
            assert err2 < 1e-6, err2
            assert n_all == n_buckets                   # every bucket reduced exactly once
            assert 1 <= n_bwd < n_buckets               # some from inside backward(), the first layer's in finish()


def test_sr_model_gradient_views_are_aligned():
    """the SR network starts with LayerNorm(3) / 9-element kernels; every gradient view of the flat bucket
    must still start on a 16-byte boundary (checked on the offsets; no process group needed)"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scale-equivariant-imaging_b200"))
    from sei_b200 import parallel
    sizes = [3, 3, 9 * 3 * 3, 3, 32 * 27, 32, 5, 128 * 32]
    offs, off = [], 0
    for n in sizes:
        offs.append(off)
        off += -(-n // parallel._ALIGN) * parallel._ALIGN
    assert all(o % 4 == 0 for o in offs) and parallel._ALIGN % 32 == 0


def test_bucket_partition():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scale-equivariant-imaging_b200"))
    from sei_b200.parallel import make_buckets
    params = [torch.nn.Parameter(torch.zeros(n)) for n in (10, 20, 5, 100, 1, 1)]
    buckets = make_buckets(params, max_elems=30)
    assert [sum(p.numel() for p in b) for b in buckets] == [30, 105, 2]
    assert [p for b in buckets for p in b] == params
