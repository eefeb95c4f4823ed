"""N > 1 host logic on CPU with the gloo backend (world size 2): ensemble sharding and the
gradient-mean exchange (reference era5_data/utils_dist.py:125-134) against the oracle."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pangu_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pangu_pytorch_b200 import dist as pdist
        from pangu_pytorch_b200 import ensemble
        # --- gradient mean: many small tensors + two "bias table"-sized ones, tiny buckets
        g = torch.Generator().manual_seed(1000 + rank)
        shapes = [(7,), (3, 5), (1, 2, 3, 16, 16), (64, 64), (1,), (1, 4, 3, 32, 32)]
        params = []
        for s in shapes:
            p = torch.nn.Parameter(torch.zeros(s))
            p.grad = torch.randn(s, generator=g)
            params.append(p)
        frozen = torch.nn.Parameter(torch.zeros(3))          # no grad: must be skipped
        mine = [p.grad.clone() for p in params]
        pdist.gather_grad(params + [frozen], bucket_bytes=4096)
        # --- overlapped reducer: the same gradients handed over in three groups (one with a None = frozen parameter)
        red = pdist.GradReducer(bucket_bytes=4096)
        again = [t.clone() for t in mine]
        red.ready(again[0:2] + [None]); red.ready(again[2:3]); red.ready(again[3:])
        red.finish()
        assert all(torch.equal(a, p.grad) for a, p in zip(again, params)), "GradReducer != gather_grad"
        # --- ensemble sharding + metadata gather
        idx = ensemble.member_indices(11, rank, world)
        meta = ensemble.gather_metadata({k: float(k) * 2 for k in idx})
        # plain python payloads: tensors in an mp.Queue would travel by shared-memory handles
        q.put((rank, [t.tolist() for t in mine], [p.grad.tolist() for p in params], idx, meta, frozen.grad))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_gloo_world2_grad_mean_and_ensemble_sharding():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=240) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = O.mean_of_grads([[torch.tensor(t) for t in res[0][1]], [torch.tensor(t) for t in res[1][1]]])
    for r in range(world):
        for got, w in zip(res[r][2], want):
            assert torch.allclose(torch.tensor(got), w, rtol=0, atol=1e-7)
        assert res[r][5] is None
    assert sorted(res[0][3] + res[1][3]) == list(range(11)) and not set(res[0][3]) & set(res[1][3])
    assert res[0][4] == res[1][4] == {k: float(k) * 2 for k in range(11)}


def test_member_indices_single_rank_and_errors():
    from pangu_pytorch_b200 import ensemble
    assert ensemble.member_indices(5, 0, 1) == [0, 1, 2, 3, 4]
    assert ensemble.member_indices(3, 3, 8) == []
    with pytest.raises(ValueError):
        ensemble.member_indices(4, 2, 2)


def test_gather_grad_is_noop_without_process_group():
    from pangu_pytorch_b200 import dist as pdist
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    pdist.gather_grad([p])
    assert torch.equal(p.grad, torch.ones(4))
