"""Static SASS evidence for the shipped library (no GPU needed): per kernel family, registers / spills and the
counts of the instructions that carry the design — 128-bit global loads/stores, evict-first loads, bulk copies
(UBLKCP) and mbarrier ops (SYNCS) of the TMA-staged tile kernel, SFU ops of the in-register Box-Muller, and the
absence of FFMA in the tableau kernels that follow the reference's separately rounded op order (-fmad=false).

    python profiles/sass_check.py > profiles/r02_sass_evidence.txt
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'torchsde_b200', 'lib', 'libtorchsde_b200.so')
PICK = [  # (label, regex on the demangled kernel name)
    ('Milstein tableau, fp32, counter noise (headline)', r'ew_fast_kernel<float, tsde::MilsteinOp<float>, 1>'),
    ('Milstein vjp seed, fp32, counter noise', r'ew_fast_kernel<float, tsde::MilsteinVjpSeedOp<float>, 1>|ew_fast_kernel<float, tsde::MilsteinSeedOp<float>, 1>'),
    ('SRK srid2 final stage, fp32', r'ew_fast_kernel<float, tsde::SrkDiagFinalOp<float>, 1>'),
    ('Euler tableau, fp64, counter noise', r'ew_fast_kernel<double, tsde::EulerOp<double>, 1>'),
    ('general Euler tile, per-thread loads, fp32', r'gen_cta_kernel<float, tsde::GEulerOp<float>, 1>'),
    ('general Heun tile, per-thread loads, fp32', r'gen_cta_kernel<float, tsde::GHeunOp<float>, 1>'),
    ('general Euler tile, TMA-staged, fp32, m=16', r'gen_tma_kernel<float, tsde::GEulerOp<float>, 1, 2>'),
    ('general Heun tile, TMA-staged, fp32, m=64', r'gen_tma_kernel<float, tsde::GHeunOp<float>, 1, 4>'),
    ('Brownian cells W (materialised queries), fp32', r'ew_fast_kernel<float, tsde::CellsOp<float, false>, 1>|ew_fast_kernel<float, tsde::CellsOp<float>, 1>'),
    ('Brownian bridge', r'bridge_kernel<float'),
    ('Levy area (Davie / Foster), fp32, m = 16: one normal per pair, packed pair arithmetic', r'levy_tile_kernel<float, false, 16>|levy_tile_kernel<float, \(bool\)0, 16>'),
    ('fused cell query W, U, A (drawn in the kernel), fp32, m = 16', r'levy_tile_kernel<float, true, 16>|levy_tile_kernel<float, \(bool\)1, 16>'),
    ('Levy area, fp64, m = 16 (scalar pair arithmetic, separately rounded)', r'levy_tile_kernel<double, false, 16>|levy_tile_kernel<double, \(bool\)0, 16>'),
    ('bmm(g, A) of the log-ODE correction, fp32, m = 16', r'bmm_ga_kernel<float, 16>'),
    ('logqp KL-integrand augmentation, fp32', r'logqp_augment_kernel<float>'),
]
COUNT = ['LDG.E.128', 'LDG.E.EF.128', 'STG.E.128', 'LDS.128', 'UBLKCP', 'SYNCS', 'MUFU', 'FFMA2', 'FMUL2', 'FADD2', 'FFMA', 'FMUL', 'FADD', 'DFMA',
         'SHFL', 'BAR.SYNC', 'LDL', 'STL', 'IMAD.WIDE']


def main():
    if not os.path.exists(LIB):
        sys.exit("build the library first: python -c 'import __graft_entry__ as g; g.build()'")
    res = subprocess.run(['cuobjdump', '-res-usage', LIB], capture_output=True, text=True).stdout
    usage = {}
    for m in re.finditer(r'Function (\S+):\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)', res):
        usage[m.group(1)] = tuple(int(m.group(i)) for i in range(2, 6))
    names = sorted(usage)
    dem = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True).stdout.splitlines()
    framed = [(n, d) for n, d in zip(names, dem) if usage[n][1] or usage[n][3]]
    f32 = [d for n, d in framed if '<float' in d]
    print(f"library: {os.path.relpath(LIB, ROOT)}   kernels: {len(names)}   max registers: {max(u[0] for u in usage.values())}")
    print(f"kernels with a stack frame: {len(framed)} (fp64 libm slow paths of log / sincospi), of which fp32: {len(f32)}")
    fam = {}
    for d in f32:
        k = d.split('(')[0].split('<')[0].replace('void ', '')
        fam[k] = fam.get(k, 0) + 1
    for k, v in sorted(fam.items()):
        print(f"   fp32 with a stack frame: {k} x{v}")
    print("   (small frames: generic fallback kernels, and the increment staging of gen_tma_kernel's producer warp — none on a"
          " hot loop; the specialised ew_fast_kernel / gen_cta_kernel instantiations have none)")
    print()
    for label, pat in PICK:
        hit = [(n, d) for n, d in zip(names, dem) if re.search(pat, d)]
        if not hit:
            print(f"## {label}\n   (no kernel matches /{pat}/)\n")
            continue
        n, d = hit[0]
        sass = subprocess.run(['cuobjdump', '-sass', '-fun', n, LIB], capture_output=True, text=True).stdout
        ops = [ln.split(';')[0] for ln in sass.splitlines() if re.match(r'\s+/\*[0-9a-f]{4,}\*/', ln)]
        reg, stack, shared, local = usage[n]
        print(f"## {label}\n   {d.split('(')[0]}\n   instructions {len(ops)}  REG {reg}  STACK {stack}  SHARED(static) {shared}  LOCAL {local}")
        cnt = {k: sum(1 for o in ops if re.search(r'\b' + re.escape(k) + r'\b', o)) for k in COUNT}
        print('   ' + '  '.join(f"{k}={v}" for k, v in cnt.items() if v) + '\n')


if __name__ == '__main__':
    main()
