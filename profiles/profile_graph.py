"""cfg2 solved from the CUDA graph (few steps) — for ncu --graph-profiling node --cache-control none."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import torchsde_b200 as tsde  # noqa: E402

w = dict(bench.WORKLOADS['cfg2'])
n = int(os.environ.get('NSTEPS', '10'))
dev = torch.device('cuda')
sde = bench.build_sde(w, dev)
ts = (torch.arange(n + 1, dtype=torch.float32) * w['dt']).to(dev)
y0 = torch.full((w['B'], w['D']), 0.1, device=dev)
for rep in range(2):
    bm = tsde.BrownianInterval(0.0, n * w['dt'], size=(w['B'], w['D']), dtype=torch.float32, device=dev, entropy=rep)
    with torch.no_grad():
        ys = tsde.sdeint(sde, y0, ts, bm=bm, method='milstein', dt=w['dt'], options={'cuda_graph': True})
torch.cuda.synchronize()
print('ok')
