"""Shared helpers of the parity tests."""
import glob
import os

import numpy as np
import torch

from . import problems

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def golden_files(prefix):
    return sorted(glob.glob(os.path.join(GOLDEN, prefix + '*.npz')))


def load(path):
    z = np.load(path, allow_pickle=False)
    return {k: z[k] for k in z.files}


def case_id(path):
    return os.path.basename(path)[:-4]


def build_problem(case, dtype=None, device='cpu'):
    tdt = {'f64': torch.float64, 'f32': torch.float32}[str(case['dtype'])] if dtype is None else dtype
    sde = problems.make(str(case['kind']), int(case['d']), int(case['m']), str(case['sde_type']), dtype=tdt,
                        seed=int(case['seed']))
    return sde.to(device)


def replay_numpy(case):
    levy = 'space-time' if 'U' in case else 'none'
    return problems.ReplayBM(case['ta'], case['tb'], case['W'], case.get('U'), levy=levy)


def replay_torch(case, device):
    levy = 'space-time' if 'U' in case else 'none'
    Ws = [torch.from_numpy(w).to(device) for w in case['W']]
    Us = [torch.from_numpy(u).to(device) for u in case['U']] if 'U' in case else None
    bm = problems.ReplayBM(case['ta'], case['tb'], Ws, Us, levy=levy)
    bm.dtype = Ws[0].dtype
    bm.device = Ws[0].device
    return bm


def tol_for(dtype_str, diag):
    """Tolerances of the parity tests (stated per BASELINE north_star: <= 1e-5 rel in fp32).
    Diagonal/scalar noise follows the reference's op order exactly -> far tighter in practice."""
    if dtype_str == 'f64':
        return dict(rtol=1e-12, atol=1e-13)
    return dict(rtol=1e-5, atol=1e-6)
