import sys, time, cProfile, pstats
sys.path.insert(0, '/root/repo')
import torch, bench
import torchsde_b200 as tsde
w = dict(bench.WORKLOADS['cfg2'])
dev = torch.device('cuda')
sde = bench.build_sde(w, dev)
B, D, T, dt = w['B'], w['D'], w['T'], w['dt']
ts = (torch.arange(T + 1, dtype=torch.float32) * dt).to(dev)
y0 = torch.full((B, D), 0.1, device=dev)
def solve(e):
    bm = tsde.BrownianInterval(0.0, T * dt, size=(B, D), dtype=torch.float32, device=dev, entropy=e)
    with torch.no_grad():
        return tsde.sdeint(sde, y0, ts, bm=bm, method='milstein', dt=dt, options={'cuda_graph': True})
for i in range(3): solve(i)
torch.cuda.synchronize()
for i in range(3):
    t0 = time.perf_counter(); solve(10+i); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"host {1e3*(t1-t0):.1f} ms, +wait {1e3*(t2-t1):.1f} ms")
pr = cProfile.Profile(); pr.enable(); solve(99); pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('cumulative').print_stats(14)
