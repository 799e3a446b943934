"""A/B of the general-noise tile kernels: per-thread loads (`gen_cta_kernel`, TSDE_GEN_TMA=0) against the
TMA-staged persistent kernel (`gen_tma_kernel`, TSDE_GEN_TMA=2).  CUDA events, rotating operand sets larger
than L2, outputs compared bit for bit.   python profiles/gen_tma_ab.py
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from torchsde_b200 import _cabi  # noqa: E402

dev = torch.device('cuda')
lib = _cabi.lib()
key = torch.tensor([123456789], dtype=torch.int64, device=dev)
PEAK = 6569.6


def run(B, D, M, nset, op):
    L = _cabi.make_launch(torch.float32, _cabi.NOISE_GENERAL, B, D, M)
    nz = _cabi.Noise()
    nz.source = _cabi.SRC_COUNTER
    nz.key = key.data_ptr()
    nz.cell_id = 42
    nz.n_cells = 1
    nz.h = 2.0 ** -10
    nz.h_total = 2.0 ** -10
    sets = [dict(y=torch.rand(B, D, device=dev), f=torch.rand(B, D, device=dev), f1=torch.rand(B, D, device=dev),
                 g=torch.rand(B, D, M, device=dev), g1=torch.rand(B, D, M, device=dev) if op == 'heun' else None,
                 o=torch.empty(B, D, device=dev)) for _ in range(nset)]
    dt = 2.0 ** -10

    def call(s):
        if op == 'euler':
            rc = lib.tsde_step_euler(ctypes.byref(L), ctypes.byref(nz), s['y'].data_ptr(), s['f'].data_ptr(),
                                     s['g'].data_ptr(), dt, s['o'].data_ptr())
        else:
            rc = lib.tsde_step_heun(ctypes.byref(L), ctypes.byref(nz), s['y'].data_ptr(), s['f'].data_ptr(),
                                    s['f1'].data_ptr(), s['g'].data_ptr(), s['g1'].data_ptr(), dt, s['o'].data_ptr())
        assert rc == 0, rc

    ng = 1 if op == 'euler' else 2
    ne = 2 if op == 'euler' else 3
    nbytes = (ng * B * D * M + (ne + 1) * B * D) * 4
    outs = {}
    for mode, kb in (('0', None), ('2', '8'), ('2', '16'), ('2', '24'), ('2', '32'), ('2', '48')):
        os.environ['TSDE_GEN_TMA'] = mode
        if kb:
            os.environ['TSDE_GEN_TMA_KB'] = kb
        for s in sets:
            call(s)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for s in sets:
                call(s)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / nset * 1e3)
        if mode == '2' and '2' in outs:
            assert torch.equal(outs['2'], sets[0]['o'])
        outs[mode] = sets[0]['o'].clone()
        print(f"{op:6s} B={B} D={D} M={M} TSDE_GEN_TMA={mode} stage_kb={kb}: {best:8.2f} us  {nbytes / best / 1e3:7.1f} GB/s "
              f"({nbytes / best / 1e3 / PEAK * 100:.1f} % of {PEAK})", flush=True)
    print("   bit-identical:", torch.equal(outs['0'], outs['2']), flush=True)


run(65536, 64, 16, 3, 'euler')
run(65536, 64, 16, 2, 'heun')
run(8192, 32, 16, 12, 'euler')
run(65536, 32, 64, 2, 'euler')
