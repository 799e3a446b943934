"""ctypes binding of the C ABI declared in ``include/torchsde_b200.h``.

This is the only place where Python meets the CUDA library.  Tensors cross as raw device
pointers (``Tensor.data_ptr()``) plus sizes; the stream is torch's current CUDA stream, so every
launch is ordered with the surrounding torch ops and can be captured into a CUDA graph.

There is deliberately no fallback: if the shared library is missing, importing the solver
raises (`LibraryNotBuilt`) telling the user to run ``python __graft_entry__.py`` (build()).
"""
import ctypes
import os

import torch

_LIB_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'lib')
# TORCHSDE_B200_LIB: load another build of the same ABI instead (A/B measurements of kernel changes on one box)
LIB_PATH = os.environ.get('TORCHSDE_B200_LIB') or os.path.join(_LIB_DIR, 'libtorchsde_b200.so')

F32, F64 = 0, 1
NOISE_DIAGONAL, NOISE_GENERAL = 0, 1
SRC_MEMORY, SRC_COUNTER, SRC_UNIT = 0, 1, 2
EINVAL = -22
FLAG_G_BROADCAST = 1


class LibraryNotBuilt(RuntimeError):
    pass


class Launch(ctypes.Structure):
    _fields_ = [('dtype', ctypes.c_int32), ('noise_type', ctypes.c_int32), ('rows', ctypes.c_int64),
                ('d', ctypes.c_int64), ('m', ctypes.c_int64), ('stream', ctypes.c_void_p)]


class Noise(ctypes.Structure):
    _fields_ = [('source', ctypes.c_int32), ('want_u', ctypes.c_int32), ('w', ctypes.c_void_p),
                ('u', ctypes.c_void_p), ('key', ctypes.c_void_p), ('cell_id', ctypes.c_uint64),
                ('row_offset', ctypes.c_int64), ('n_cells', ctypes.c_int32), ('flags', ctypes.c_int32),
                ('h', ctypes.c_double), ('cell_h', ctypes.c_void_p), ('h_total', ctypes.c_double)]


_P = ctypes.c_void_p
_D = ctypes.c_double
_I = ctypes.c_int32
_L = ctypes.POINTER(Launch)
_N = ctypes.POINTER(Noise)

# name -> argtypes after (launch, [noise]).  Mirrors include/torchsde_b200.h one to one.
SIGNATURES = {
    'tsde_brownian_cells': [_L, _N, _P, _P, _P],
    'tsde_brownian_bridge': [_L, _P, ctypes.c_int64, _I, ctypes.POINTER(ctypes.c_uint64),
                             ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_double), _P, _P, _P, _P],
    'tsde_brownian_merge': [_L, _P, _P, _P, _P, _D, _D, _D],
    'tsde_brownian_h_to_u': [_L, _P, _P, _D, _P],
    'tsde_brownian_levy_area': [_L, _P, ctypes.c_int64, ctypes.c_uint64, _P, _P, _D, _I, _P],
    'tsde_brownian_merge_area': [_L, _P, _P, _P, _P],
    'tsde_bmm_ga': [_L, _P, _P, _P],
    'tsde_logqp_augment': [_L, _P, _P, _P, _D, _P, _P],
    'tsde_brownian_cell_levy': [_L, _N, ctypes.c_uint64, _I, _P, _P, _P],
    'tsde_step_euler': [_L, _N, _P, _P, _P, _D, _P],
    'tsde_milstein_vjp_seed': [_L, _N, _P, _D, _I, _P],
    'tsde_step_milstein': [_L, _N, _P, _P, _P, _P, _D, _P],
    'tsde_milstein_gf_predict': [_L, _P, _P, _P, _D, _D, _I, _P],
    'tsde_step_milstein_gf': [_L, _N, _P, _P, _P, _P, _D, _D, _I, _P],
    'tsde_step_heun': [_L, _N, _P, _P, _P, _P, _P, _D, _P],
    'tsde_midpoint_predict': [_L, _N, _P, _P, _P, _D, _P],
    'tsde_euler_heun_predict': [_L, _N, _P, _P, _P],
    'tsde_step_euler_heun': [_L, _N, _P, _P, _P, _P, _D, _P],
    'tsde_reversible_heun_z': [_L, _N, _P, _P, _P, _P, _D, _P],
    'tsde_step_reversible_heun': [_L, _N, _P, _P, _P, _P, _P, _D, _P],
    'tsde_srk_diag_stage1': [_L, _P, _P, _P, _D, _D, _P, _P],
    'tsde_srk_diag_stage2': [_L, _N, _P, _P, _P, _P, _P, _D, _D, _D, _P, _P],
    'tsde_srk_diag_stage3': [_L, _P, _P, _P, _P, _P, _D, _D, _P],
    'tsde_step_srk_diag': [_L, _N, _P, _P, _P, _P, _P, _P, _P, _P, _D, _D, _D, _D, _P],
    'tsde_srk_additive_stage': [_L, _N, _P, _P, _P, _D, _D, _P],
    'tsde_step_srk_additive': [_L, _N, _P, _P, _P, _P, _P, _D, _D, _P],
    'tsde_linear_interp': [_L, _P, _P, _D, _D, _P],
    'tsde_adaptive_error_sumsq': [_L, _P, _P, _D, _D, _D, _P, _P],
    'tsde_adjoint_reversible_heun_a': [_L, _N, _P, _P, _P, _P, _P, _P, _P, _D, _D, _P, _P, _P],
    'tsde_adjoint_reversible_heun_b': [_L, _N, _P, _P, _P, _P, _P, _P, _P, _P, _D, _D, _P, _P, _P, _P, _P],
}

_lib = None


def lib():
    """Load the shared library (once) and attach prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryNotBuilt(
            f"torchsde_b200: CUDA library not found at {LIB_PATH}. Build it with "
            f"`python -c 'import __graft_entry__ as g; g.build()'` from the repository root. "
            f"There is no CPU or PyTorch fallback.")
    handle = ctypes.CDLL(LIB_PATH)
    handle.tsde_abi_version.restype = ctypes.c_int
    handle.tsde_error_string.restype = ctypes.c_char_p
    handle.tsde_error_string.argtypes = [ctypes.c_int]
    handle.tsde_kernel_launches.restype = ctypes.c_int64
    handle.tsde_kernel_launches.argtypes = [ctypes.c_int32]
    for name, argtypes in SIGNATURES.items():
        fn = getattr(handle, name)  # AttributeError here = header/library mismatch: fail loudly
        fn.restype = ctypes.c_int
        fn.argtypes = argtypes
    if handle.tsde_abi_version() != 1:
        raise LibraryNotBuilt("torchsde_b200: ABI version mismatch, rebuild the library.")
    _lib = handle
    return _lib


LAUNCHES = 0  # number of C-ABI launch calls issued by this process (bench.py reports it)


def check(code, what):
    global LAUNCHES
    LAUNCHES += 1
    if code != 0:
        msg = lib().tsde_error_string(code).decode()
        raise RuntimeError(f"torchsde_b200: {what} failed: {msg} (code {code})")


def dtype_code(dtype):
    if dtype == torch.float32:
        return F32
    if dtype == torch.float64:
        return F64
    raise ValueError(f"torchsde_b200 supports float32 and float64 tensors, got {dtype}.")


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "torchsde_b200 is a CUDA (sm_100a) implementation: tensors must live on a CUDA device. "
                "There is no CPU path; use the reference torchsde for CPU solves.")


def make_launch(dtype, noise_type, rows, d, m, stream=None, device=None):
    if stream is None:
        stream = torch.cuda.current_stream(device).cuda_stream  # the stream of the tensors' device, not of the current one
    return Launch(dtype_code(dtype), noise_type, rows, d, m, stream)


def device_guard(device):
    """Make `device` the current CUDA device for the duration of a solve / query: the C ABI launches on the stream
    it is handed and never calls cudaSetDevice, and a launch on a stream of another device than the current one
    is an invalid resource handle.  No-op for non-CUDA devices (the host-side dry runs of the test-suite)."""
    import contextlib
    device = torch.device(device) if device is not None else None
    if device is None or device.type != 'cuda':
        return contextlib.nullcontext()
    return torch.cuda.device(device)


class nvtx_range:
    """NVTX range around a host-side phase (solve, graph capture, replay, backward sweep) when TSDE_NVTX=1 — what
    shows up as named spans on an Nsight Systems timeline (tracing).  Free when the variable is unset."""
    _on = os.environ.get('TSDE_NVTX', '0') not in ('0', '')

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if self._on:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if self._on:
            torch.cuda.nvtx.range_pop()


def ptr(t):
    """Device pointer of a contiguous tensor (or NULL)."""
    if t is None:
        return None
    if not t.is_contiguous():
        raise RuntimeError("torchsde_b200: internal error, non-contiguous tensor at the C ABI.")
    return t.data_ptr()
