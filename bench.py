#!/usr/bin/env python
"""Benchmark of the hot path on BASELINE.json's metric: trajectory-steps/s (batch x t_steps / s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]

A "step" is one full solve of the workload (one pass of the hot path over one batch).
Workload cfg2 (BASELINE.json configs[1], the configuration the metric is quoted on):
    Milstein Ito/diagonal, batch=65536, state=64, t=1000 steps (dt=2^-10), fp32, full output series
    ts = arange(1001)*dt, SDE = per-channel GBM (f = mu*y, g = sigma*y) as ordinary torch callables,
    Brownian motion = torchsde_b200.BrownianInterval (counter-based, regenerated in registers).

Output: ONE JSON line (rank 0) — see the keys below; `roofline` and `cpu_baseline` as the tier
contract asks; `--impl reference` times the CPU oracle port (the Python reference cannot travel to
the GPU box) on the host cores with the same config/metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[1] — the configuration the metric is quoted on (the headline)
    'cfg2': dict(method='milstein', sde_type='ito', kind='gbm', B=65536, D=64, M=64, T=1000, dt=2.0 ** -10,
                 E_bytes_per_traj_step=13 * 64 * 4, S_tableau_bytes_per_traj_step=5 * 64 * 4),
    'cfg2_small': dict(method='milstein', sde_type='ito', kind='gbm', B=4096, D=64, M=64, T=100, dt=2.0 ** -10,
                       E_bytes_per_traj_step=13 * 64 * 4, S_tableau_bytes_per_traj_step=5 * 64 * 4),
    # honest HBM-bound point of SURVEY §8(d): working set 320 MiB >> L2
    'cfg2_b262144': dict(method='milstein', sde_type='ito', kind='gbm', B=262144, D=64, M=64, T=100, dt=2.0 ** -10,
                         E_bytes_per_traj_step=13 * 64 * 4, S_tableau_bytes_per_traj_step=5 * 64 * 4),
    # BASELINE.json configs[2], substituted (reference SRK rejects general noise, srk.py:35): parity/measurement cases
    'cfg3_srk_additive': dict(method='srk', sde_type='ito', kind='additive', B=8192, D=32, M=16, T=500,
                              dt=2.0 ** -10, levy='space-time', E_bytes_per_traj_step=(11 * 32 + 5 * 32 * 16) * 4),
    'cfg3_euler_general': dict(method='euler', sde_type='ito', kind='general', B=8192, D=32, M=16, T=500,
                               dt=2.0 ** -10, E_bytes_per_traj_step=(6 * 32 + 2 * 32 * 16) * 4),
    'cfg3_heun_general': dict(method='heun', sde_type='stratonovich', kind='general', B=8192, D=32, M=16, T=500,
                              dt=2.0 ** -10, E_bytes_per_traj_step=2 * (6 * 32 + 2 * 32 * 16) * 4),
    # cfg2 with per-trajectory parameters (mu, sigma of shape (B, D)): f/g/vjp are contiguous element-wise
    # products, i.e. PyTorch's vectorised kernels instead of its broadcasting ones.  E = 7 (solver) + 3*3 = 16 D s.
    'cfg2_pertraj': dict(method='milstein', sde_type='ito', kind='gbm_pertraj', B=65536, D=64, M=64, T=1000,
                         dt=2.0 ** -10, E_bytes_per_traj_step=16 * 64 * 4, S_tableau_bytes_per_traj_step=5 * 64 * 4),
    # general noise in the HBM-bound regime (g = 256 MiB per evaluation)
    'cfg3_euler_general_large': dict(method='euler', sde_type='ito', kind='general', B=65536, D=64, M=16, T=100,
                                     dt=2.0 ** -10, E_bytes_per_traj_step=(6 * 64 + 2 * 64 * 16) * 4),
    # other diagonal tableaus at the cfg2 size
    'cfg2_euler': dict(method='euler', sde_type='ito', kind='gbm', B=65536, D=64, M=64, T=200, dt=2.0 ** -10,
                       E_bytes_per_traj_step=8 * 64 * 4),
    'cfg2_srk': dict(method='srk', sde_type='ito', kind='gbm', B=65536, D=64, M=64, T=200, dt=2.0 ** -10,
                     levy='space-time', E_bytes_per_traj_step=41 * 64 * 4),
    'cfg2_milstein_gf': dict(method='milstein', sde_type='ito', kind='gbm', B=65536, D=64, M=64, T=200,
                             dt=2.0 ** -10, options={'grad_free': True}, E_bytes_per_traj_step=15 * 64 * 4),
}
METRIC = "trajectory-steps/s (batch x t_steps / s)"


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (profiling recipe)."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.proc = None
        self.index = index
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            p = [x.strip() for x in ln.split(',')]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for n, v in zip(names, p[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_sde(w, device, dtype=torch.float32):
    from tests import problems
    torch.manual_seed(1147481649)
    kind = w.get('kind', 'gbm')
    if kind == 'gbm':
        sde = problems.GBMDiagonal(w['D'], w['sde_type'], seed=1147481649 % 1000, dtype=dtype)
    elif kind == 'gbm_pertraj':
        sde = problems.GBMPerTrajectory(w['B'], w['D'], w['sde_type'], seed=649, dtype=dtype)
    else:
        sde = problems.make(kind, w['D'], w['M'], w['sde_type'], dtype=dtype, seed=649)
    return sde.to(device)


# -------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm, on the host cores
# -------------------------------------------------------------------------------------------------
def cpu_port_run(w, n_steps, threads, B=None):
    """Oracle Milstein (oracle/solvers.py) + oracle counter-based Brownian cells on `threads` host
    threads (rows are independent; numpy releases the GIL inside ufuncs).  Returns traj-steps/s."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import brownian as obm
    from oracle import solvers
    from tests import problems
    B = B or w['B']
    D = w['D']
    dt = w['dt']
    sde_t = build_sde(w, 'cpu')
    mu = sde_t.mu.detach().numpy()
    sigma = sde_t.sigma.detach().numpy()

    class NpGBM:  # numpy callables of the same synthetic SDE (f = mu*y, g = sigma*y; gdg = vjp)
        noise_type, sde_type = 'diagonal', w['sde_type']
        f = staticmethod(lambda t, y: mu * y)
        g = staticmethod(lambda t, y: sigma * y)
        gdg = staticmethod(lambda t, y, v2: (sigma * y) * v2 * sigma)

    ts = np.array([0.0, n_steps * dt], dtype=np.float32)
    chunks = np.array_split(np.arange(B), threads)

    def work(rows):
        r0, n = int(rows[0]), len(rows)

        def bm(ta, tb, return_U=False):
            k = int(round(float(ta) / dt))
            W, _ = obm.cell(1234567, 1000 + k, float(tb) - float(ta), n, D, np.float32, False, row_offset=r0)
            return W
        y0 = np.full((n, D), 0.1, dtype=np.float32)
        solvers.make('milstein', NpGBM, bm, dt).integrate(y0, ts)

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(work, chunks))
    el = time.perf_counter() - t0
    return B * n_steps / el, el


def run_reference(args, w, rank, world):
    if rank != 0:
        return
    threads, tried = best_cpu_threads(w)
    n_steps = 8
    vals = []
    for i in range(args.warmup + args.steps):
        v, el = cpu_port_run(w, n_steps, threads)
        if i >= args.warmup:
            vals.append((v, el))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([el for _, el in vals]) * 1e3)
    sample = f"B={w['B']} D={w['D']} first {n_steps} of {w['T']} steps per bench step"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "traj-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, **{k: w[k] for k in ('method', 'sde_type', 'B', 'D', 'T')},
                       "cpu_sample": sample},
            "cpu_baseline": {"value": value, "unit": "traj-steps/s", "cores": threads, "kind": "port",
                             "sample": sample, "host_cores": os.cpu_count(),
                             "threads_tried": {str(k): round(v) for k, v in tried.items()}},
            "e2e": {"value": value, "unit": "traj-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


_CPU_THREADS = {}


def best_cpu_threads(w):
    """The numpy port stops scaling once its per-thread slices get small (interpreter overhead under the GIL), so
    'all the host threads' is not its fastest setting: try a ladder of thread counts on one step each and keep the
    fastest, so that the CPU arm is reported at its best.  Returns (threads, {threads: traj-steps/s})."""
    key = (w['B'], w['D'])
    if key not in _CPU_THREADS:
        n = os.cpu_count() or 1
        tried = {}
        for c in sorted({min(c, n) for c in (4, 8, 16, 32, 64, 128, n)}):
            cpu_port_run(w, 1, c)                      # warm-up (thread start, allocations)
            tried[c] = max(cpu_port_run(w, 1, c)[0], cpu_port_run(w, 1, c)[0])
        _CPU_THREADS[key] = (max(tried, key=tried.get), tried)
    return _CPU_THREADS[key]


# -------------------------------------------------------------------------------------------------
# GPU arm
# -------------------------------------------------------------------------------------------------
def run_ours(args, w, rank, world, local_rank):
    import torch.distributed as dist
    import torchsde_b200 as tsde
    from torchsde_b200 import _cabi
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    _cabi.lib()  # fail loudly if the CUDA library is missing: there is no fallback
    B, D, T, dt = w['B'], w['D'], w['T'], w['dt']
    sde = build_sde(w, dev)
    ts = (torch.arange(T + 1, dtype=torch.float32) * dt).to(dev)
    y0_host = torch.full((B, D), 0.1, dtype=torch.float32).pin_memory()
    y0_dev = y0_host.to(dev)
    out_host = torch.empty((B, D), dtype=torch.float32).pin_memory()
    opts = {'cuda_graph': not args.no_graph}
    if args.row_split > 1:
        opts['row_split'] = args.row_split
    row_offset = rank * B  # weak scaling: every rank integrates its own B trajectories of one global batch

    M = D if w.get('kind', 'gbm').startswith('gbm') else w['M']
    opts.update(w.get('options', {}))

    def solve(y0, entropy):
        bm = tsde.BrownianInterval(0.0, T * dt, size=(B, M), dtype=torch.float32, device=dev, entropy=entropy,
                                   levy_area_approximation=w.get('levy', 'none'))
        bm.shard_rows(row_offset)
        with torch.no_grad():
            return tsde.sdeint(sde, y0, ts, bm=bm, method=w['method'], dt=dt, options=dict(opts))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident throughput (`value`) ----
    for i in range(args.warmup):
        solve(y0_dev, 1000 + i)
    barrier()
    launches0 = _cabi.LAUNCHES
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        ys = solve(y0_dev, 2000 + i)
    ev1.record()
    barrier()
    elapsed = ev0.elapsed_time(ev1) * 1e-3
    clocks = sampler.stop() if rank == 0 else None
    eager_launches = _cabi.LAUNCHES - launches0
    finite = bool(torch.isfinite(ys[-1]).all().item())

    # ---- end to end through the public API with HOST buffers (`e2e`) ----
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gathered = torch.empty((world * B, D), dtype=torch.float32, device=dev) if world > 1 else None
    e0.record()
    for i in range(args.steps):
        y0 = y0_host.to(dev, non_blocking=True)           # H2D of the step's inputs
        ys = solve(y0, 3000 + i)
        if world > 1:
            # the one collective of the path: gather the row shards' terminal states over NVLink
            dist.all_gather_into_tensor(gathered, ys[-1].contiguous())
        out_host.copy_(ys[-1], non_blocking=True)          # D2H of the step's result (terminal states)
    e1.record()
    barrier()
    e2e_elapsed = e0.elapsed_time(e1) * 1e-3

    per_rank_ms = [elapsed / args.steps * 1e3]
    if world > 1:
        mine = torch.tensor([elapsed], device=dev, dtype=torch.float64)
        every = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        per_rank_ms = [float(x.item()) / args.steps * 1e3 for x in every]
        t = torch.tensor([elapsed, e2e_elapsed], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed, e2e_elapsed = t.tolist()
    total_traj_steps = world * B * T * args.steps
    value = total_traj_steps / elapsed
    e2e_value = total_traj_steps / e2e_elapsed

    # ---- roofline of the dominant kernel (the fused Milstein tableau), timed in situ ----
    headline = args.workload.startswith('cfg2') and w['method'] == 'milstein' and not w.get('options') \
        and w.get('kind', 'gbm') == 'gbm'
    roof = tableau_roofline(w, sde, dev) if (rank == 0 and headline) else None
    if rank != 0:
        return
    peak, peak_src = peaks()
    per_solve_kernels = {'milstein': 2, 'euler': 1, 'heun': 2, 'srk': 4 if w.get('kind') != 'additive' else 2}.get(
        w['method'], 2) * T  # solver-owned kernel launches per solve (aligned outputs)
    if w.get('options', {}).get('grad_free'):
        per_solve_kernels = 2 * T
    cpu_threads, cpu_tried = best_cpu_threads(w) if headline else (1, {})
    cpu_val, cpu_el = cpu_port_run(w, 4, cpu_threads) if headline else (None, None)
    E = w['E_bytes_per_traj_step']
    line = {
        "metric": METRIC, "value": value, "unit": "traj-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "method": w['method'], "sde_type": w['sde_type'], "noise": "diagonal",
                   "batch_per_gpu": B, "state": D, "t_steps": T, "dt": dt, "output": "full series (T+1,B,D)",
                   "cuda_graph": opts['cuda_graph'], "row_split": args.row_split, "l2_policy": "inputs larger than L2: per-step working set "
                   "80 MiB + 16 MiB ys row streamed into a 16.8 GB series", "parallelism": f"batch-sharded x{world}",
                   "finite": finite},
        "clocks": clocks,
        "per_rank_ms_per_step": [round(x, 3) for x in per_rank_ms],
        "e2e": {"value": e2e_value, "unit": "traj-steps/s", "h2d_bytes_per_step": int(y0_host.numel() * 4),
                "d2h_bytes_per_step": int(out_host.numel() * 4), "ms_per_step": e2e_elapsed / args.steps * 1e3,
                "result_copied": "ys[-1] (terminal states)",
                "collective": None if world == 1 else "one NCCL all_gather of the terminal states per solve"},
        "gpu_launches": int(per_solve_kernels * args.steps),
        "host_launch_calls_in_timed_region": int(eager_launches),
        "roofline": None if roof is None else {
            "bound": "hbm", "achieved": roof['gbs'], "peak": peak, "unit": "GB/s",
            "frac": roof['gbs'] / peak, "traffic": roof.get('traffic'),
            "kernel": "ew_fast_kernel<float, MilsteinOp, COUNTER> (tsde_step_milstein)",
            "algorithmic_bytes_per_launch": roof['bytes'], "avg_launch_us": roof['us'], "peak_source": peak_src,
            "timing": "CUDA events around a graph replay of back-to-back launches of the kernel (how the solver "
                      "issues them), cfg2 tensor sizes, rotating buffer sets larger than L2; median of 7"},
        "roofline_whole_step": {"E_bytes_per_traj_step": E, "achieved": value / world * E / 1e9, "peak": peak,
                                "unit": "GB/s", "frac": value / world * E / 1e9 / peak,
                                "note": "SURVEY §8(d) E-bytes: solver kernels + the synthetic SDE's own f/g/vjp"},
        "cpu_baseline": None if cpu_val is None else {
            "value": cpu_val, "unit": "traj-steps/s", "cores": cpu_threads, "kind": "port",
            "sample": f"oracle Milstein + oracle Philox cells, B={B} D={D}, first 4 of {T} steps",
            "host_cores": os.cpu_count(), "threads_tried": {str(k): round(v) for k, v in cpu_tried.items()}},
    }
    print(json.dumps(line), flush=True)


def tableau_roofline(w, sde, dev):
    """Average duration of the dominant solver kernel (fused Milstein tableau) at the workload's tensor
    sizes: CUDA events on the launching stream around back-to-back launches through the C ABI, each launch
    on a different buffer set (12 sets x 5 tensors x 16 MiB = 960 MiB, larger than the 126 MB L2)."""
    import ctypes
    from torchsde_b200 import _cabi
    B, D, dt = w['B'], w['D'], w['dt']
    lib = _cabi.lib()
    nset = max(2, min(12, int(2e9 // (5 * B * D * 4))))
    sets = [[torch.rand(B, D, device=dev) for _ in range(5)] for _ in range(nset)]
    key = torch.tensor([987654321], dtype=torch.int64, device=dev)
    L = _cabi.make_launch(torch.float32, _cabi.NOISE_DIAGONAL, B, D, D)
    nz = _cabi.Noise()
    nz.source, nz.key, nz.cell_id, nz.n_cells, nz.h, nz.h_total = _cabi.SRC_COUNTER, key.data_ptr(), 7, 1, dt, dt

    def launch(s):
        _cabi.check(lib.tsde_step_milstein(ctypes.byref(L), ctypes.byref(nz), s[0].data_ptr(), s[1].data_ptr(),
                                           s[2].data_ptr(), s[3].data_ptr(), dt, s[4].data_ptr()), "tsde_step_milstein")
    for s_ in sets:
        launch(s_)
    torch.cuda.synchronize(dev)
    # the solver replays its launches from a CUDA graph, so time them the same way: nset launches
    # (each on a different buffer set) captured once, replayed and bracketed by CUDA events
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        L.stream = torch.cuda.current_stream(dev).cuda_stream  # launch on the capturing stream
        for s_ in sets:
            launch(s_)
    L.stream = torch.cuda.current_stream(dev).cuda_stream
    graph.replay()
    torch.cuda.synchronize(dev)
    times = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize(dev)
        times.append(e0.elapsed_time(e1) * 1e3 / nset)
    us = float(np.median(times))
    nbytes = w['S_tableau_bytes_per_traj_step'] * B
    # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this shape from the committed
    # `ncu --set full` capture (profiles/r01_ncu_full_all_kernels.csv): 67.15 MB + 3.33 MB (the remaining
    # writes were still in L2 when the kernel retired)
    traffic = 70.48e6 if (B, D) == (65536, 64) else None
    return {"us": us, "bytes": nbytes, "gbs": nbytes / (us * 1e-6) / 1e9, "traffic": traffic}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg2', choices=sorted(WORKLOADS))
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--row-split', type=int, default=1)
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if os.environ.get('TSDE_BENCH_B'):  # experiments only: override the batch size
        w['B'] = int(os.environ['TSDE_BENCH_B'])
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, w, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        torch.cuda.set_device(local_rank)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    try:
        run_ours(args, w, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
