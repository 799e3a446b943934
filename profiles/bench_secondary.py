"""Secondary measurements (BASELINE.json configs[3] and configs[4]); one JSON line each.

  cfg4  sdeint_adjoint reversible_heun / adjoint_reversible_heun, Stratonovich diagonal latent-SDE-like
        model (f = MLP(D+1 -> 128 -> D) softplus, g = 0.1 sigmoid(w*y + b)), B=32768, D=128, T=256 steps,
        loss = ys[-1].pow(2).sum(1).mean(); forward + backward.
  cfg5  BrownianInterval(0, 1, size=(B,16), dt=2^-6), 64 sequential queries then a random permutation,
        levy in {none, space-time, foster}, B in 2^10 .. 2^20.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import torchsde_b200 as tsde  # noqa: E402
from tests import problems  # noqa: E402

dev = torch.device('cuda')


def cfg4(B=32768, D=128, T=256, reps=3, graph=True):
    dt = 2.0 ** -8
    sde = problems.LatentLike(D, hidden=128, seed=0).to(dev)
    ts = torch.tensor([0.0, T * dt], device=dev)
    y0 = torch.full((B, D), 0.1, device=dev)

    def step(entropy):
        bm = tsde.BrownianInterval(0.0, T * dt, size=(B, D), dtype=torch.float32, device=dev, entropy=entropy)
        ys = tsde.sdeint_adjoint(sde, y0, ts, bm=bm, method='reversible_heun', dt=dt,
                                 options={'cuda_graph': graph}, adjoint_options={'cuda_graph': graph})
        loss = ys[-1].pow(2).sum(1).mean()
        for p in sde.parameters():
            p.grad = None
        loss.backward()
        return loss

    for i in range(2):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        loss = step(10 + i)
    e1.record()
    torch.cuda.synchronize()
    el = e0.elapsed_time(e1) * 1e-3 / reps
    print(json.dumps({"workload": "cfg4", "metric": "trajectory-steps/s (fwd+bwd)", "value": B * T / el,
                      "ms_per_step": el * 1e3, "B": B, "D": D, "T": T, "cuda_graph": graph,
                      "loss": float(loss), "grad_norm": float(sum(p.grad.norm() ** 2 for p in sde.parameters()) ** .5)}),
          flush=True)


def cfg5():
    M, n = 16, 64
    dt = 2.0 ** -6
    rng = np.random.RandomState(0)
    for levy in ('none', 'space-time', 'foster'):
        for logB in (10, 14, 17, 20):
            B = 2 ** logB
            if levy == 'foster' and logB > 17:
                continue
            out = {}
            for order in ('sequential', 'random'):
                idx = np.arange(n) if order == 'sequential' else rng.permutation(n)
                best = None
                for rep in range(3):
                    bm = tsde.BrownianInterval(0.0, 1.0, size=(B, M), dtype=torch.float32, device=dev,
                                               entropy=1147481649 + rep, dt=dt, levy_area_approximation=levy)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    for k in idx:
                        r = bm(k * dt, (k + 1) * dt, return_U=levy != 'none', return_A=levy == 'foster')
                    torch.cuda.synchronize()
                    el = time.perf_counter() - t0
                    best = el if best is None else min(best, el)
                out[order] = B * n / best
            written = {'none': M * 4, 'space-time': 2 * M * 4, 'foster': (2 * M + M * M) * 4}[levy]
            print(json.dumps({"workload": "cfg5", "levy": levy, "B": B, "M": M, "queries": n,
                              "row_queries_per_s_sequential": out['sequential'],
                              "row_queries_per_s_random": out['random'],
                              "GBps_written_sequential": out['sequential'] * written / 1e9}), flush=True)


if __name__ == '__main__':
    which = sys.argv[1:] or ['cfg4', 'cfg5']
    if 'cfg4' in which:
        cfg4(graph=False)
        cfg4(graph=True)
    if 'cfg5' in which:
        cfg5()
