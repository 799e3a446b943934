"""A few eager (no CUDA graph) steps of a cfg3 workload for an ncu launch list: how many kernels one step is, which
of them are the user's PyTorch ops, and what share is this library's.  At B=8192, D=32, M=16 the state is 1 MiB and
the diffusion 16 MiB: every kernel lasts a few microseconds and the step is a chain of dependent launches.

    CFG3=srk_additive_expand ncu --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/r02_launches_cfg3.csv python profiles/cfg3_eager.py
    python profiles/launch_shares.py gpurun_out/r02_launches_cfg3.csv
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import torchsde_b200 as tsde  # noqa: E402

name = 'cfg3_' + os.environ.get('CFG3', 'srk_additive_expand')
w = bench.WORKLOADS[name]
T = int(os.environ.get('CFG3_T', 8))
dev = torch.device('cuda')
sde = bench.build_sde(w, dev)
ts = (torch.arange(T + 1, dtype=torch.float32) * w['dt']).to(dev)
y0 = torch.full((w['B'], w['D']), 0.1, device=dev)
bm = tsde.BrownianInterval(0.0, T * w['dt'], size=(w['B'], w['M']), dtype=torch.float32, device=dev, entropy=5,
                           levy_area_approximation=w.get('levy', 'none'))
with torch.no_grad():
    ys = tsde.sdeint(sde, y0, ts, bm=bm, method=w['method'], dt=w['dt'], options={'cuda_graph': False})
torch.cuda.synchronize()
print(name, T, 'steps ok', float(ys[-1].abs().mean()))
