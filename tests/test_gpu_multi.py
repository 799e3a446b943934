"""Two-GPU test (NCCL): batch-sharded sdeint_adjoint == single-GPU solve.  Skipped on boxes with one GPU
(run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    import torchsde_b200 as tsde
    from torchsde_b200 import parallel
    from tests import problems
    dev = torch.device('cuda', rank)
    B, d = 64, 8
    sde = problems.GBMDiagonal(d, 'stratonovich', seed=3, dtype=torch.float64).to(dev)
    ts = torch.tensor([0.0, 0.25, 0.5], dtype=torch.float64, device=dev)
    y_full = (torch.arange(B * d, dtype=torch.float64, device=dev).reshape(B, d) % 7) * 0.1 + 0.2

    def loss_of(ys):
        return ys.pow(2).sum()

    # sharded: every rank integrates its rows of the SAME Brownian motion (global row keys)
    y_loc, bm = parallel.shard_rows(y_full, lambda n: tsde.BrownianInterval(0.0, 0.5, size=(n, d), dtype=torch.float64,
                                                                            device=dev, entropy=99))
    y_loc = y_loc.clone().requires_grad_(True)
    ys_loc = tsde.sdeint_adjoint(sde, y_loc, ts, bm=bm, method='reversible_heun', dt=2.0 ** -4)
    loss_of(ys_loc).backward()
    parallel.all_reduce_grads(sde.parameters())          # the one collective of a sharded training step
    ys_all = parallel.all_gather_rows(ys_loc.detach(), B, dim=1)
    gy_all = parallel.all_gather_rows(y_loc.grad, B, dim=0)
    sharded = [p.grad.clone() for p in sde.parameters()]
    ok = True
    if rank == 0:
        for p in sde.parameters():
            p.grad = None
        y_ref = y_full.clone().requires_grad_(True)
        bm_ref = tsde.BrownianInterval(0.0, 0.5, size=(B, d), dtype=torch.float64, device=dev, entropy=99)
        ys_ref = tsde.sdeint_adjoint(sde, y_ref, ts, bm=bm_ref, method='reversible_heun', dt=2.0 ** -4)
        loss_of(ys_ref).backward()
        ok = torch.equal(ys_ref.detach(), ys_all) and torch.equal(y_ref.grad, gy_all)
        for a, p in zip(sharded, sde.parameters()):
            ok = ok and torch.allclose(a, p.grad, rtol=1e-12, atol=1e-12)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_sharded_adjoint_equals_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res), res
