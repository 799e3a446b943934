"""A few eager steps of cfg4 (sdeint_adjoint, reversible Heun, latent-SDE-like MLP drift, B=32768 D=128) for an ncu
launch list: which kernels the forward + backward step consists of and what share of the time is this library's.

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_cfg4.csv \
        python profiles/cfg4_eager.py
    python profiles/launch_shares.py gpurun_out/r02_launches_cfg4.csv
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import torchsde_b200 as tsde  # noqa: E402
from tests import problems  # noqa: E402

dev = torch.device('cuda')
B, D, T, dt = 32768, 128, int(os.environ.get('CFG4_T', 6)), 2.0 ** -10
sde = problems.LatentLike(D, hidden=128, seed=0).to(dev)
ts = (torch.arange(T + 1, dtype=torch.float32) * dt).to(dev)
y0 = torch.full((B, D), 0.1, device=dev)
for rep in range(2):
    bm = tsde.BrownianInterval(0.0, T * dt, size=(B, D), dtype=torch.float32, device=dev, entropy=3 + rep)
    ys = tsde.sdeint_adjoint(sde, y0, ts, bm=bm, method='reversible_heun', adjoint_method='adjoint_reversible_heun', dt=dt)
    ys[-1].pow(2).sum(1).mean().backward()
torch.cuda.synchronize()
print('ok')
