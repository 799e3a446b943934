"""world_size-2 CPU test (gloo) of the batch-sharding host logic (no kernels involved)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_rows, q):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from torchsde_b200 import parallel
    import torchsde_b200 as tsde
    lo, hi = parallel.row_range(n_rows)
    full = torch.arange(n_rows * 3, dtype=torch.float32).reshape(n_rows, 3)
    y_local, bm = parallel.shard_rows(full, lambda n: tsde.BrownianInterval(0., 1., size=(n, 3), device='cuda',
                                                                            entropy=7))
    assert y_local.shape[0] == hi - lo and bm._row_offset == lo and bm.shape == (hi - lo, 3)
    # series gathered along the batch axis (dim 1 of (T, B, D))
    series = torch.stack([y_local, y_local + 1000.0])
    gathered = parallel.all_gather_rows(series, n_rows, dim=1)
    ok_gather = torch.equal(gathered, torch.stack([full, full + 1000.0]))
    p = torch.nn.Parameter(torch.zeros(2))
    p.grad = torch.tensor([float(rank + 1), 1.0])
    parallel.all_reduce_grads([p])
    ok_grad = torch.equal(p.grad, torch.tensor([float(sum(range(1, world + 1))), float(world)]))
    q.put((rank, lo, hi, ok_gather, ok_grad))
    dist.destroy_process_group()


def test_sharding_gather_and_grad_reduce_world2():
    world, n_rows = 2, 7
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_rows, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 4), (4, 7)]
    assert all(r[3] and r[4] for r in res)


def test_row_range_partition():
    from torchsde_b200 import parallel
    for n in (1, 7, 8, 65536):
        for world in (1, 2, 3, 8):
            spans = [parallel.row_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
