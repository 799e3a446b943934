"""A few eager steps of the cfg2 workload (Milstein Ito/diagonal, B=65536, D=64, fp32), for ncu.

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python profiles/profile_step.py
    ncu --set full --clock-control none --import-source on -k regex:ew_kernel -s 8 -c 4 \
        -o gpurun_out/prof python profiles/profile_step.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import torchsde_b200 as tsde  # noqa: E402

w = dict(bench.WORKLOADS[os.environ.get('WORKLOAD', 'cfg2')])
n = int(os.environ.get('NSTEPS', '12'))
dev = torch.device('cuda')
sde = bench.build_sde(w, dev)
ts = (torch.arange(n + 1, dtype=torch.float32) * w['dt']).to(dev)
y0 = torch.full((w['B'], w['D']), 0.1, device=dev)
M = w['D'] if w.get('kind', 'gbm') == 'gbm' else w['M']
bm = tsde.BrownianInterval(0.0, n * w['dt'], size=(w['B'], M), dtype=torch.float32, device=dev, entropy=1,
                           levy_area_approximation=w.get('levy', 'none'))
with torch.no_grad():
    ys = tsde.sdeint(sde, y0, ts, bm=bm, method=w['method'], dt=w['dt'], options=w.get('options'))
torch.cuda.synchronize()
print('ok', float(ys[-1].mean()))
