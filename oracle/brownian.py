"""Brownian-source arithmetic on the CPU: restatement of the reference's formulas.

Test infrastructure only (see oracle/__init__.py).  Every function cites the reference lines
(torchsde/_brownian/brownian_interval.py) it follows.  Arrays are numpy, dtype float32/float64;
python-float coefficients multiply arrays exactly as `python_float * torch.Tensor` does (the
scalar is rounded once to the array dtype).
"""
import math

import numpy as np

from . import philox

_rsqrt3 = 1 / math.sqrt(3)
_r12 = 1 / 12


def _s(x, dtype):
    """python scalar -> array dtype (what torch does for `python_float * tensor`)."""
    return np.dtype(dtype).type(x)


def bridge(W, H, start, mid, end, is_left, X1, X2=None):
    """Child (W, H) from the parent's (W, H): brownian_interval.py:188-241.
    With H (:199-225) or W only (H is None, :226-237)."""
    dt = W.dtype
    h_reciprocal = 1 / (end - start)
    left_diff = mid - start
    right_diff = end - mid
    if H is not None:
        left_diff_squared = left_diff ** 2
        right_diff_squared = right_diff ** 2
        left_diff_cubed = left_diff * left_diff_squared
        right_diff_cubed = right_diff * right_diff_squared
        v = 0.5 * math.sqrt(left_diff * right_diff / (left_diff_cubed + right_diff_cubed))
        a = v * left_diff_squared * h_reciprocal
        b = v * right_diff_squared * h_reciprocal
        c = v * _rsqrt3
        third_coeff = 2 * (a * left_diff + b * right_diff) * h_reciprocal
        if is_left:
            first_coeff = left_diff * h_reciprocal
            second_coeff = 6 * first_coeff * right_diff * h_reciprocal
            out_W = _s(first_coeff, dt) * W + _s(second_coeff, dt) * H + _s(third_coeff, dt) * X1
            out_H = _s(first_coeff ** 2, dt) * H - _s(a, dt) * X1 + _s(c * right_diff, dt) * X2
        else:
            first_coeff = right_diff * h_reciprocal
            second_coeff = 6 * first_coeff * left_diff * h_reciprocal
            out_W = _s(first_coeff, dt) * W - _s(second_coeff, dt) * H - _s(third_coeff, dt) * X1
            out_H = _s(first_coeff ** 2, dt) * H - _s(b, dt) * X1 - _s(c * left_diff, dt) * X2
        return out_W, out_H
    mean = _s(left_diff, dt) * W * _s(h_reciprocal, dt)
    var = left_diff * right_diff * h_reciprocal
    left_W = mean + _s(math.sqrt(var), dt) * X1
    return (left_W if is_left else W - left_W), None


def merge(W, H, Wi, Hi, ta, start_i, end_i):
    """Aggregate the running (W, H) over [ta, start_i] with interval i = [start_i, end_i]:
    brownian_interval.py:649-658, 672."""
    dt = W.dtype
    if H is not None:
        term1 = _s(end_i - start_i, dt) * (Hi + _s(0.5, dt) * W)
        term2 = _s(start_i - ta, dt) * (H - _s(0.5, dt) * Wi)
        H = (term1 + term2) / _s(end_i - ta, dt)
    return W + Wi, H


def merge_area(A, Ai, W, Wi):
    """brownian_interval.py:671 (uses W before it is updated)."""
    dt = W.dtype
    return A + Ai + _s(0.5, dt) * (W[..., :, None] * Wi[..., None, :] - Wi[..., :, None] * W[..., None, :])


def h_to_u(W, H, h):
    """_H_to_U, brownian_interval.py:102-103."""
    dt = W.dtype
    return _s(h, dt) * (_s(.5, dt) * W + H)


def davie_foster(W, H, h, foster, noise):
    """_davie_foster_approximation for W.ndim >= 2, brownian_interval.py:85-99."""
    dt = W.dtype
    A = H[..., :, None] * W[..., None, :] - W[..., :, None] * H[..., None, :]
    noise = noise - np.swapaxes(noise, -1, -2)
    if foster:
        tenth_h = _s(0.1 * h, dt)
        H_squared = H ** 2
        std = np.sqrt(tenth_h * (tenth_h + H_squared[..., :, None] + H_squared[..., None, :]))
    else:
        std = _s(math.sqrt(_r12 * h ** 2), dt)
    return A + std * noise


# ---- this repository's counter-based source (spec restated; see philox.py) ----------------------
def cell(key, cell_id, h, rows, m, dtype, have_h, row_offset=0, row_ids=None):
    """Direct draw of a primary cell: W = sqrt(h) N_W, H = sqrt(h/12) N_H — the law of the top
    interval in the reference (brownian_interval.py:553-558)."""
    W = philox.normals(key, cell_id, philox.STREAM_W, rows, m, dtype, row_offset, row_ids) * _s(math.sqrt(h), dtype)
    H = None
    if have_h:
        H = philox.normals(key, cell_id, philox.STREAM_H, rows, m, dtype, row_offset, row_ids) * _s(math.sqrt(h / 12), dtype)
    return W, H


def cells(key, cell_id, lengths, rows, m, dtype, have_h, row_offset=0, row_ids=None):
    """Merge of consecutive primary cells (ids cell_id, cell_id+1, ...), left to right with `merge`
    (elapsed time accumulated in float64, as csrc/ew.cuh counter_noise does)."""
    W, H = cell(key, cell_id, lengths[0], rows, m, dtype, have_h, row_offset, row_ids)
    elapsed = lengths[0]
    for c in range(1, len(lengths)):
        Wi, Hi = cell(key, (cell_id + c) & ((1 << 64) - 1), lengths[c], rows, m, dtype, have_h, row_offset, row_ids)
        # merge() with ta = 0, start_i = elapsed, end_i = elapsed + len  (term coefficients: len, elapsed, total)
        dt = W.dtype
        if have_h:
            term1 = _s(lengths[c], dt) * (Hi + _s(0.5, dt) * W)
            term2 = _s(elapsed, dt) * (H - _s(0.5, dt) * Wi)
            H = (term1 + term2) / _s(elapsed + lengths[c], dt)
        W = W + Wi
        elapsed += lengths[c]
    return W, H


def bridge_chain(key, W, H, levels, row_offset=0):
    """Descend binary splits. levels: list of (parent_id, is_left, start, mid, end)."""
    rows, m = W.shape
    for parent_id, is_left, start, mid, end in levels:
        X1 = philox.normals(key, parent_id, philox.STREAM_X1, rows, m, W.dtype, row_offset)
        X2 = philox.normals(key, parent_id, philox.STREAM_X2, rows, m, W.dtype, row_offset) if H is not None else None
        W, H = bridge(W, H, start, mid, end, is_left, X1, X2)
    return W, H


def levy_noise(key, a_id, rows, m, dtype, row_offset=0, row_ids=None):
    """(rows, m, m) matrix N whose antisymmetric part N - N^T is the Levy-area noise of `davie_foster` (the reference
    draws m^2 iid normals and antisymmetrises, brownian_interval.py:88-90; only N - N^T enters, with independent
    N(0, 2) entries per pair).  This repository's source draws one normal z per pair i < j — channel p of stream
    STREAM_A, p counting the pairs in row-major upper-triangular order — and sets N_ij = z / sqrt(2), N_ji = -N_ij,
    N_ii = 0 (csrc/brownian.cu)."""
    npairs = m * (m - 1) // 2
    z = philox.normals(key, a_id, philox.STREAM_A, rows, max(npairs, 1), dtype, row_offset, row_ids)
    n = np.zeros((z.shape[0], m, m), dtype=dtype)
    iu, ju = np.triu_indices(m, k=1)
    half = (z[:, :npairs] * _s(0.70710678118654752440, dtype)).astype(dtype)
    n[:, iu, ju] = half
    n[:, ju, iu] = -half
    return n
