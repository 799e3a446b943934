"""Fixed-step time grid of a solve, computed once on the host.

Restates the control flow of ``BaseSDESolver.integrate`` (reference:
torchsde/_core/base_solver.py:92-149, fixed-step branch :143-147) *without* running it on the
device: the reference evaluates ``while curr_t < out_t`` and ``min(curr_t + dt, ts[-1])`` on 0-d
device tensors, i.e. with two or three host<->device syncs per step.  Here the same expressions
are evaluated once with 0-d CPU tensors of the same dtype (so rounding — e.g. the accumulation
of ``curr_t + dt`` in fp32, which gives 1001 steps for dt=1e-3 on [0,1] — is bit-identical), and
the resulting plan is what the CUDA-graph time loop executes.
"""
import weakref

import torch

_VALUES = {}     # (id(ts), version) -> (weakref(ts), tuple of python floats)
_SCHEDULES = {}  # (id(ts), version, dt) -> (weakref(ts), Schedule)
_MAX_CACHED = 16


def _remember(cache, key, value):
    while len(cache) >= _MAX_CACHED:
        cache.pop(next(iter(cache)))
    cache[key] = value


def ts_values(ts):
    """The evaluation times as python floats.  Cached per tensor object (identity + in-place version
    counter), so that repeated solves on the same `ts` tensor do not synchronise with the device."""
    key = (id(ts), ts._version)
    hit = _VALUES.get(key)
    if hit is not None and hit[0]() is ts:
        return hit[1]
    vals = tuple(ts.detach().to('cpu').tolist())
    _remember(_VALUES, key, (weakref.ref(ts), vals))
    return vals


def get_schedule(ts, dt):
    """`build_schedule` with a cache: planning a 1000-step grid costs ~10 ms of host time (0-d tensor
    arithmetic in ts' dtype, deliberately identical to the reference's), which would otherwise be paid —
    serialised with the GPU by the device->host read of `ts` — on every solve of a training loop."""
    # dt is keyed by VALUE (and dtype: it takes part in the grid's type promotion), never by identity: a training loop
    # that passes a fresh `torch.tensor(dt)` per call gets recycled object ids, and an identity key would then serve the
    # grid of an earlier, different step size.  (`ts` is keyed by identity + version and guarded by a weak reference.)
    dkey = (float(dt), str(dt.dtype)) if torch.is_tensor(dt) else float(dt)
    key = (id(ts), ts._version, dkey)
    hit = _SCHEDULES.get(key)
    if hit is not None and hit[0]() is ts:
        return hit[1]
    sched = build_schedule(ts, dt)
    sched.values = ts_values(ts)
    _remember(_SCHEDULES, key, (weakref.ref(ts), sched))
    return sched


class Output:
    """Row `index` of ys is produced after step `step`:
    aligned -> ys[index] is that step's y1 itself; else the linear interpolation
    w0 * prev_y + w1 * curr_y of _core/interp.py:15-18 (weights computed in ts' dtype)."""
    __slots__ = ('index', 'step', 'aligned', 'w0', 'w1')

    def __init__(self, index, step, aligned, w0, w1):
        self.index = index
        self.step = step
        self.aligned = aligned
        self.w0 = w0
        self.w1 = w1


class Schedule:
    """steps[k] = (t0, t1) as 0-d CPU tensors in ts' dtype; outputs = list of Output."""

    def __init__(self, ts_cpu, steps, outputs):
        self.ts = ts_cpu
        self.steps = steps
        self.outputs = outputs
        self.n_steps = len(steps)
        # python floats (exact) of the step boundaries, used by the Brownian grid binding
        self.bounds = [float(steps[0][0])] + [float(s[1]) for s in steps] if steps else [float(ts_cpu[0])]
        outs_after = {}
        for o in outputs:
            outs_after.setdefault(o.step, []).append(o)
        self.outputs_after = outs_after
        self._aligned = {o.step: o.index for o in outputs if o.aligned}

    def aligned_row(self, k):
        """Row of ys that coincides with the end of step k (or None)."""
        return self._aligned.get(k)


def build_schedule(ts, dt):
    """ts: 1-D tensor (any device); dt: python float or 0-d tensor.  base_solver.py:107-147."""
    ts_cpu = ts.detach().to('cpu')
    step_size = dt.detach().to('cpu') if torch.is_tensor(dt) else dt
    curr_t = ts_cpu[0]
    prev_t = curr_t
    end_t = ts_cpu[-1]
    steps = []
    outputs = []
    for i in range(1, ts_cpu.numel()):
        out_t = ts_cpu[i]
        while curr_t < out_t:
            next_t = min(curr_t + step_size, end_t)
            if not (next_t > curr_t):
                raise ValueError("Step size `dt` is too small to advance time in the dtype of `ts`.")
            prev_t = curr_t
            steps.append((curr_t, next_t))
            curr_t = next_t
        k = len(steps) - 1
        if k < 0:
            raise ValueError("Evaluation times `ts` must be strictly increasing.")
        aligned = bool(curr_t == out_t)
        if aligned:
            w0, w1 = 0.0, 1.0
        else:
            # interp.py:17 — (t1 - t)/(t1 - t0) * y0 + (t - t0)/(t1 - t0) * y1
            w0 = float((curr_t - out_t) / (curr_t - prev_t))
            w1 = float((out_t - prev_t) / (curr_t - prev_t))
        outputs.append(Output(i, k, aligned, w0, w1))
    return Schedule(ts_cpu, steps, outputs)
