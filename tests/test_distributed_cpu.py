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
    results[rank] = (err, same_weights)
    dist.destroy_process_group()


def test_data_parallel_gradient_averaging_gloo():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
        assert len(results) == world
        for rank in range(world):
            err, same = results[rank]
            assert err < 1e-6, err
            assert same


def test_bucket_partition():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scale-equivariant-imaging_b200"))
    from sei_b200.parallel import make_buckets
    params = [torch.nn.Parameter(torch.zeros(n)) for n in (10, 20, 5, 100, 1, 1)]
    buckets = make_buckets(params, max_elems=30)
    assert [sum(p.numel() for p in b) for b in buckets] == [30, 105, 2]
    assert [p for b in buckets for p in b] == params
