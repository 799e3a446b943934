"""Summarise an `ncu --set full` report of profiles/kernels_for_ncu.py (read here, no GPU needed):

    python profiles/ncu_extract.py gpurun_out/r02_full.ncu-rep profiles/r02_ncu_full_kernels.csv

writes the per-kernel summary CSV (committed evidence) and profiles/ncu_traffic.json, which bench.py reads for
`roofline.traffic` (dram__bytes_read.sum + dram__bytes_write.sum per launch at the cfg2 shape).
"""
import csv
import io
import json
import os
import subprocess
import sys

rep, out_csv = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, body = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'sm__inst_executed.sum', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'lts__t_bytes.sum']
NAMES = {'MilsteinSeedOp': 'milstein_vjp_seed', 'MilsteinOp': 'step_milstein', 'EulerOp<': 'step_euler',
         'SrkDiagFinalOp': 'step_srk_diag', 'CellsOp<float, false>': 'brownian_cells_W',
         'CellsOp<float, true>': 'brownian_cells_WU', 'levy_area': 'brownian_levy_area', 'bmm_ga': 'bmm_ga'}


def to_bytes(v, unit):
    v = float(v.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)


seen = {}
with open(out_csv, 'w', newline='') as f:
    w = csv.writer(f)
    w.writerow(['kernel', 'grid'] + [f"{m} [{units[col[m]]}]" for m in WANT if m in col])
    for r in body:
        name = r[col['Kernel Name']]
        w.writerow([name[:110], r[col['Grid Size']] if 'Grid Size' in col else ''] + [r[col[m]] for m in WANT if m in col])
        for key, short in NAMES.items():
            if key in name and 'dram__bytes_read.sum' in col:
                rd = to_bytes(r[col['dram__bytes_read.sum']], units[col['dram__bytes_read.sum']])
                wr = to_bytes(r[col['dram__bytes_write.sum']], units[col['dram__bytes_write.sum']])
                seen[short] = rd + wr   # last launch of each kernel (the script launches each three times)
seen['source'] = f"{os.path.basename(out_csv)} (ncu --set full of profiles/kernels_for_ncu.py: dram__bytes_read.sum + dram__bytes_write.sum of the last launch of each kernel)"
json.dump(seen, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ncu_traffic.json'), 'w'), indent=1)
print(json.dumps(seen, indent=1))
