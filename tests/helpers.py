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


# ---- oracle view of a grid-bound torchsde_b200.BrownianInterval on a SAMPLE of rows ------------------
MASK64 = (1 << 64) - 1


def sample_rows(n_rows, n_sample, seed):
    """Sorted random sample of global row indices that always contains the first and the last row."""
    rng = np.random.default_rng(seed)
    n_sample = min(n_sample, n_rows)
    rows = set(rng.choice(n_rows, size=n_sample, replace=False).tolist()) | {0, n_rows - 1}
    return np.array(sorted(rows), dtype=np.int64)


def oracle_grid_bm(bm, row_ids, m, npdt, have_h):
    """numpy `bm(ta, tb, return_U=False)` reproducing — for the global rows `row_ids` only — the path of a
    BrownianInterval whose root is a GRID node (what a fixed-step solve binds): every query must be a run of
    whole primary cells.  Rows are independent Philox streams (oracle/philox.py), so a full-size solve can be
    checked on a sample of its trajectories."""
    from oracle import brownian as obm
    grid, key = bm._root, bm._key
    assert grid.kind == 2, "the Brownian motion is not grid-bound"
    index = {b: i for i, b in enumerate(grid.bounds)}
    row_ids = np.asarray(row_ids, dtype=np.int64) + int(bm._row_offset)

    def query(ta, tb, return_U=False):
        i, j = index[float(ta)], index[float(tb)]
        lengths = [grid.bounds[k + 1] - grid.bounds[k] for k in range(i, j)]
        W, H = obm.cells(key, (grid.cell_base + i) & MASK64, lengths, len(row_ids), m, npdt, have_h,
                         row_ids=row_ids)
        if return_U:
            return W, obm.h_to_u(W, H, float(tb) - float(ta))
        return W
    return query


def rel_err(got, ref, floor=1e-6):
    """max |got - ref| / max(|ref|, floor) over all elements."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), floor)))
