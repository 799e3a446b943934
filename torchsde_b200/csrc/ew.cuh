// Fused row-wise tableau kernel framework (diagonal noise, and every purely element-wise
// stage).  One thread owns one "quad": 4 consecutive channels of one trajectory, which is
// exactly what one Philox4x32 call yields in fp32 — so the Brownian increment of the quad is
// produced in registers and never touches HBM (north star: "dW in registers").
//
// HBM plan (B200: 148 SMs, ~6.5 TB/s measured copy bandwidth): each input tensor is read
// once with 128-bit loads, each output written once with 128-bit stores; every thread issues
// all of its loads before the first dependent use (NIN independent LDG.128 in flight per
// thread); the grid is one resident wave (SM count x occupancy), each CTA owning an equal
// contiguous slice of the quads.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <mutex>
#include <type_traits>
#include <unordered_map>

#include "../../include/torchsde_b200.h"
#include "philox.cuh"

namespace tsde {

constexpr int kThreads = 256;
constexpr int kSMs = 148;        // B200
constexpr int kBlocksPerSM = 8;  // 2048 threads / SM

inline bool stream_loads_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TSDE_STREAM");
    v = (e && e[0] == '0') ? 0 : 1;  // on by default; TSDE_STREAM=0 for A/B measurements
  }
  return v == 1;
}

inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TSDE_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// TSDE_EW_CTAS=n: use at most n resident CTAs per SM for the persistent grids of the row-wise kernels (experiments:
// measured on cfg2, fewer resident CTAs only lose — 47.6 ms per solve at 4, 49.0 at 3, 51.6 at 2).
inline int ctas_per_sm_limit() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TSDE_EW_CTAS");
    v = e ? atoi(e) : 0;
  }
  return v;
}

inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kSMs;
  }
  return n;
}

// Resident CTAs per SM of one kernel instantiation at kThreads threads, queried once per kernel.
// (Keyed by the kernel's address: instantiations that differ only in a non-type template argument
// share a function-pointer type, so a function-local static would be shared between them.)
template <typename K>
inline int resident_ctas(K kernel) {
  static std::mutex mu;
  static std::unordered_map<const void*, int> cache;
  const void* id = reinterpret_cast<const void*>(kernel);
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(id);
  if (it != cache.end()) return it->second;
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kThreads, 0) != cudaSuccess || n < 1) n = 1;
  cache.emplace(id, n);
  return n;
}

template <typename T>
struct NoiseP {
  const T* w;         // MEMORY
  const T* u;         // MEMORY
  const void* key;    // COUNTER
  uint64_t cell_id;
  int64_t row_offset;
  int32_t n_cells;
  int32_t bcast;      // noise has a single channel shared by all d (scalar noise, squeezed g)
  double h;           // uniform cell length
  const double* cell_h;  // device, or nullptr
  double h_total;     // tb - ta of the whole query (for U)
  int64_t m;          // channels of the noise tensor
  T sqrt_h;           // (T)sqrt(h), (T)sqrt(h/12), (T)h_total: host-rounded once (single-cell path)
  T sqrt_h12;
  T ht;
};

template <int NIN, int NOUT>
struct EwP {
  const void* in[NIN > 0 ? NIN : 1];
  void* out[NOUT];
  int64_t rows;
  int64_t d;
  int64_t qpr;     // quads per row
  int64_t nquads;  // rows * qpr
  int32_t vec;     // all pointers 16B-aligned and d % 4 == 0
  int32_t qshift;  // log2(qpr) if qpr is a power of two, else -1
  int32_t small;   // nquads < 2^31: 32-bit index arithmetic
  uint64_t qmagic; // ceil(2^40 / qpr): row = (Q * qmagic) >> 40, exact while Q * qpr < 2^40
};

// ---- vector load / store helpers ---------------------------------------------------------
__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void ld4(const double* p, double (&v)[4]) {
  const double2 a = *reinterpret_cast<const double2*>(p);
  const double2 b = *reinterpret_cast<const double2*>(p + 2);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void st4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void st4(double* p, const double (&v)[4]) {
  *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
  *reinterpret_cast<double2*>(p + 2) = make_double2(v[2], v[3]);
}

// streaming variants (ld.global.cs: evict-first) for operands that are dead after this kernel
__device__ __forceinline__ void ld4cs(const float* p, float (&v)[4]) {
  const float4 t = __ldcs(reinterpret_cast<const float4*>(p));
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void ld4cs(const double* p, double (&v)[4]) {
  const double2 a = __ldcs(reinterpret_cast<const double2*>(p));
  const double2 b = __ldcs(reinterpret_cast<const double2*>(p + 2));
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

template <typename Op, typename = void>
struct streams_inputs { static constexpr bool value = false; };
template <typename Op>
struct streams_inputs<Op, decltype((void)Op::STREAM_INPUTS)> { static constexpr bool value = Op::STREAM_INPUTS; };

template <typename T>
__device__ __forceinline__ void load_quad(const T* p, int64_t base, bool vec, int nvalid,
                                          T (&v)[4]) {
  if (vec) {
    ld4(p + base, v);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = j < nvalid ? p[base + j] : T(0);
  }
}
template <typename T>
__device__ __forceinline__ void store_quad(T* p, int64_t base, bool vec, int nvalid,
                                           const T (&v)[4]) {
  if (vec) {
    st4(p + base, v);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < nvalid) p[base + j] = v[j];
  }
}

// ---- Brownian increment of one quad --------------------------------------------------------
// Counter mode: merge of n_cells primary cells, left to right, with the reference's
// aggregation rule (brownian_interval.py:643-672):
//     H <- ( len_i (H_i + W/2) + (start_i - ta)(H - W_i/2) ) / (end_i - ta) ;  W <- W + W_i
// Lengths are host doubles rounded once to T (python-float * tensor semantics).
constexpr int kSrcCounterMulti = 3;  // internal: COUNTER source merging several primary cells

template <typename T, bool WANT_U, bool MULTI = true>
__device__ __forceinline__ void counter_noise(const NoiseP<T>& nz, Key key, uint32_t row,
                                              uint32_t q, T (&w)[4], T (&u)[4]) {
  T hh[4];
  if (!MULTI || nz.n_cells == 1) {
    // the solver's own grid: one primary cell per step, scales rounded once on the host
    T n[4];
    normal4(key, nz.cell_id, STREAM_W, row, q, n);
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = n[j] * nz.sqrt_h;
    if (WANT_U) {
      normal4(key, nz.cell_id, STREAM_H, row, q, n);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        hh[j] = n[j] * nz.sqrt_h12;
        u[j] = nz.ht * (T(0.5) * w[j] + hh[j]);  // _H_to_U :102-103
      }
    }
    return;
  }
  if (!MULTI) return;  // (unreachable; lets the compiler drop the merge loop from single-cell kernels)
  double len0 = nz.cell_h ? nz.cell_h[0] : nz.h;
  {
    T n[4];
    normal4(key, nz.cell_id, STREAM_W, row, q, n);
    const T s = (T)sqrt(len0);
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = n[j] * s;
    if (WANT_U) {
      normal4(key, nz.cell_id, STREAM_H, row, q, n);
      const T s12 = (T)sqrt(len0 / 12.0);
#pragma unroll
      for (int j = 0; j < 4; ++j) hh[j] = n[j] * s12;
    }
  }
  double elapsed = len0;  // start_i - ta
  for (int c = 1; c < nz.n_cells; ++c) {
    const double len = nz.cell_h ? nz.cell_h[c] : nz.h;
    T n[4], wi[4];
    normal4(key, nz.cell_id + (uint64_t)c, STREAM_W, row, q, n);
    const T s = (T)sqrt(len);
#pragma unroll
    for (int j = 0; j < 4; ++j) wi[j] = n[j] * s;
    if (WANT_U) {
      normal4(key, nz.cell_id + (uint64_t)c, STREAM_H, row, q, n);
      const T s12 = (T)sqrt(len / 12.0);
      const T tl = (T)len, te = (T)elapsed, tt = (T)(elapsed + len);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const T hi = n[j] * s12;
        const T term1 = tl * (hi + T(0.5) * w[j]);
        const T term2 = te * (hh[j] - T(0.5) * wi[j]);
        hh[j] = (term1 + term2) / tt;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = w[j] + wi[j];
    elapsed += len;
  }
  if (WANT_U) {
    const T ht = (T)nz.h_total;
#pragma unroll
    for (int j = 0; j < 4; ++j) u[j] = ht * (T(0.5) * w[j] + hh[j]);  // _H_to_U :102-103
  }
}

template <typename T, int SRC, bool WANT_U>
__device__ __forceinline__ void quad_noise(const NoiseP<T>& nz, Key key, int64_t row, int64_t q,
                                           bool vec, int nvalid, T (&w)[4], T (&u)[4]) {
  if (SRC == TSDE_SRC_UNIT) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { w[j] = T(1); u[j] = T(0); }
  } else if (SRC == TSDE_SRC_MEMORY) {
    if (nz.bcast) {
      const T a = nz.w[row];
      const T b = WANT_U ? nz.u[row] : T(0);
#pragma unroll
      for (int j = 0; j < 4; ++j) { w[j] = a; u[j] = b; }
    } else {
      const int64_t base = row * nz.m + 4 * q;
      load_quad(nz.w, base, vec, nvalid, w);
      if (WANT_U) load_quad(nz.u, base, vec, nvalid, u);
    }
  } else {
    constexpr bool MULTI = SRC == kSrcCounterMulti;
    const uint32_t grow = (uint32_t)(row + nz.row_offset);
    if (nz.bcast) {
      T w4[4], u4[4];
      counter_noise<T, WANT_U, MULTI>(nz, key, grow, 0u, w4, u4);
#pragma unroll
      for (int j = 0; j < 4; ++j) { w[j] = w4[0]; u[j] = WANT_U ? u4[0] : T(0); }
    } else {
      counter_noise<T, WANT_U, MULTI>(nz, key, grow, (uint32_t)q, w, u);
    }
  }
}

// ---- the kernel -----------------------------------------------------------------------------
// Op: struct with  static constexpr int NIN, NOUT; static constexpr bool USES_NOISE, WANT_U;
//     template<T> __device__ void operator()(const T (&in)[NIN], T w, T u, T (&out)[NOUT]) const
template <typename T, typename Op, int SRC>
__global__ void __launch_bounds__(kThreads, 4)
ew_kernel(const EwP<Op::NIN, Op::NOUT> p, const NoiseP<T> nz, const Op op) {
  constexpr int NIN = Op::NIN, NOUT = Op::NOUT;
  Key key{0u, 0u};
  if (Op::USES_NOISE && (SRC == TSDE_SRC_COUNTER || SRC == kSrcCounterMulti)) key = load_key(nz.key);
  const bool vec = p.vec != 0;
  // Each CTA owns one contiguous slice of the quads, slices differ by at most one quad: with
  // gridDim = SMs x resident CTAs every SM gets the same amount of work (no tail wave, no
  // ceil(trip count) imbalance between SMs).
  const int64_t q_begin = (p.nquads * (int64_t)blockIdx.x) / gridDim.x;
  const int64_t q_end = (p.nquads * ((int64_t)blockIdx.x + 1)) / gridDim.x;
  for (int64_t Q = q_begin + threadIdx.x; Q < q_end; Q += kThreads) {
    int64_t row, q;
    if (p.qshift >= 0) {
      row = Q >> p.qshift;
      q = Q & ((1 << p.qshift) - 1);
    } else if (p.small) {
      const uint32_t r32 = (uint32_t)Q / (uint32_t)p.qpr;
      row = r32;
      q = (uint32_t)Q - r32 * (uint32_t)p.qpr;
    } else {
      row = Q / p.qpr;
      q = Q - row * p.qpr;
    }
    const int64_t base = row * p.d + 4 * q;
    const int64_t rem = p.d - 4 * q;
    const int nvalid = rem < 4 ? (int)rem : 4;
    T in[NIN > 0 ? NIN : 1][4];
#pragma unroll
    for (int i = 0; i < NIN; ++i) load_quad(reinterpret_cast<const T*>(p.in[i]), base, vec, nvalid, in[i]);
    T w[4], u[4];
    if (Op::USES_NOISE) {
      quad_noise<T, SRC, Op::WANT_U>(nz, key, row, q, vec, nvalid, w, u);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) { w[j] = T(0); u[j] = T(0); }
    }
    T out[NOUT][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      T a[NIN > 0 ? NIN : 1], b[NOUT];
#pragma unroll
      for (int i = 0; i < NIN; ++i) a[i] = in[i][j];
      op(a, w[j], u[j], b);
#pragma unroll
      for (int i = 0; i < NOUT; ++i) out[i][j] = b[i];
    }
#pragma unroll
    for (int i = 0; i < NOUT; ++i) store_quad(reinterpret_cast<T*>(p.out[i]), base, vec, nvalid, out[i]);
  }
}

// ---- fast variant: the shapes the headline path uses -------------------------------------------
// Preconditions checked on the host: 128-bit aligned tensors with d % 4 == 0 (vector path), noise not
// broadcast, quads-per-row a power of two (or divisible by multiply-shift), fewer than 2^31 quads.
// Everything is 32-bit index arithmetic and there is no per-quad branching.
//
// (Tried and rejected, r02: software-pipelining the integer half of iteration i+1's noise (Philox) next to the float
// half of iteration i (Box-Muller) — 10-15 % SLOWER on every RNG-bound kernel (seed 8.5 -> 9.7 us, cells 6.6 -> 7.4):
// the carried Philox state costs more than the pipe mixing gains.)
// Light ops (at most 3 tensors per quad: the Milstein vjp seed, the Brownian materialisation, the
// predictor stages) are NOT HBM-bound per thread: the Philox + Box-Muller chain (~150 dependent-ish
// instructions per quad) dominates and one 128-bit load per thread does not cover the HBM latency-bandwidth
// product (ncu r02: issue active 57 %, 6 eligible warps per scheduler, every pipe below 40 %).  Those ops
// process U = 2 quads per thread per iteration: both loads are issued first, the two independent Philox
// chains interleave (2x ILP, 2x bytes in flight per thread).  Heavy ops (>= 4 tensors) keep U = 1: their
// loads already cover the latency and the extra registers would cost occupancy.
template <typename T, typename Op>
struct quads_per_iter { static constexpr int value = (sizeof(T) == 4 && Op::NIN + Op::NOUT <= 3) ? 2 : 1; };

template <typename T, typename Op>
struct FastCtx {
  const EwP<Op::NIN, Op::NOUT>& p;
  const NoiseP<T>& nz;
  const Op& op;
  Key key;
  uint32_t q_end, qshift, qmask, qpr32, row_off;
  uint64_t qmagic;
  bool pow2, stream_hint;
  __device__ __forceinline__ uint32_t row_of(uint32_t Q) const {
    return pow2 ? (Q >> qshift) : (uint32_t)(((uint64_t)Q * qmagic) >> 40);
  }
  __device__ __forceinline__ uint32_t quad_of(uint32_t Q, uint32_t row) const {
    return pow2 ? (Q & qmask) : (Q - row * qpr32);
  }
  __device__ __forceinline__ void rng(uint32_t Q, T (&w)[4], T (&u)[4]) const {
    const uint32_t r = row_of(Q);
    counter_noise<T, Op::WANT_U, false>(nz, key, r + row_off, quad_of(Q, r), w, u);
  }
};

// One iteration: U quads Q, Q + kThreads, ... (each warp access stays one contiguous 512-byte run).
// FIRST: the counter noise was produced ahead of the dependency wait and is passed in (w0, u0).
template <typename T, typename Op, int SRC, int U, bool FIRST>
__device__ __forceinline__ void ew_fast_body(const FastCtx<T, Op>& c, uint32_t Q, const T (&w0)[U][4],
                                             const T (&u0)[U][4]) {
  constexpr int NIN = Op::NIN, NOUT = Op::NOUT;
  bool ok[U];
  T in[U][NIN > 0 ? NIN : 1][4];
#pragma unroll
  for (int k = 0; k < U; ++k) {
    const uint32_t Qk = Q + (uint32_t)k * kThreads;
    ok[k] = k == 0 || Qk < c.q_end;
    const size_t base = (size_t)Qk * 4;  // d == 4 * qpr: quads are laid out contiguously
    if (ok[k]) {
#pragma unroll
      for (int i = 0; i < NIN; ++i) {
        if (streams_inputs<Op>::value && c.stream_hint)
          ld4cs(reinterpret_cast<const T*>(c.p.in[i]) + base, in[k][i]);
        else
          ld4(reinterpret_cast<const T*>(c.p.in[i]) + base, in[k][i]);
      }
    }
  }
  T w[U][4], u[U][4];
#pragma unroll
  for (int k = 0; k < U; ++k) {
    const uint32_t Qk = Q + (uint32_t)k * kThreads;
    const size_t base = (size_t)Qk * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) { w[k][j] = T(0); u[k][j] = T(0); }
    if (Op::USES_NOISE) {
      if (SRC == TSDE_SRC_COUNTER) {
        if (FIRST) {
#pragma unroll
          for (int j = 0; j < 4; ++j) { w[k][j] = w0[k][j]; u[k][j] = Op::WANT_U ? u0[k][j] : T(0); }
        } else {
          c.rng(Qk, w[k], u[k]);  // unconditional (a quad past the slice end costs nothing observable): the U
                                  // Philox chains must stay in one basic block to interleave
        }
      } else if (SRC == TSDE_SRC_MEMORY) {
        if (ok[k]) {
          ld4(c.nz.w + base, w[k]);
          if (Op::WANT_U) ld4(c.nz.u + base, u[k]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) { w[k][j] = T(1); u[k][j] = T(0); }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < U; ++k) {
    if (!ok[k]) continue;
    const size_t base = (size_t)(Q + (uint32_t)k * kThreads) * 4;
    T out[NOUT][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      T a[NIN > 0 ? NIN : 1], b[NOUT];
#pragma unroll
      for (int i = 0; i < NIN; ++i) a[i] = in[k][i][j];
      c.op(a, w[k][j], u[k][j], b);
#pragma unroll
      for (int i = 0; i < NOUT; ++i) out[i][j] = b[i];
    }
#pragma unroll
    for (int i = 0; i < NOUT; ++i) st4(reinterpret_cast<T*>(c.p.out[i]) + base, out[i]);
  }
}

template <typename T, typename Op, int SRC>
__global__ void __launch_bounds__(kThreads, 4)
ew_fast_kernel(const EwP<Op::NIN, Op::NOUT> p, const NoiseP<T> nz, const Op op) {
  constexpr int U = quads_per_iter<T, Op>::value;
  constexpr bool COUNTER = Op::USES_NOISE && SRC == TSDE_SRC_COUNTER;
  const uint32_t nquads = (uint32_t)p.nquads;
  const uint32_t q_begin = (uint32_t)(((uint64_t)nquads * blockIdx.x) / gridDim.x);
  const uint32_t q_end = (uint32_t)(((uint64_t)nquads * (blockIdx.x + 1)) / gridDim.x);
  const bool pow2 = p.qshift >= 0;
  const uint32_t qshift = pow2 ? (uint32_t)p.qshift : 0u;
  const FastCtx<T, Op> c{p, nz, op, COUNTER ? load_key(nz.key) : Key{0u, 0u}, q_end, qshift, (1u << qshift) - 1u,
                         (uint32_t)p.qpr, (uint32_t)nz.row_offset, p.qmagic, pow2,
                         p.vec > 1 /* host sets vec = 2 to enable evict-first loads */};
  // Programmatic dependent launch: this grid may start while its predecessor in the stream/graph is
  // still draining.  Everything that does not touch the predecessor's outputs — the Philox/Box-Muller
  // work of the thread's first U quads — runs before `griddepcontrol.wait`; all loads and stores come after.
  T w0[U][4], u0[U][4];
  const uint32_t Q0 = q_begin + threadIdx.x;
  if (COUNTER) {
#pragma unroll
    for (int k = 0; k < U; ++k) {
      c.rng(Q0 + (uint32_t)k * kThreads, w0[k], u0[k]);
    }
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (Q0 >= q_end) return;
  ew_fast_body<T, Op, SRC, U, true>(c, Q0, w0, u0);
  for (uint32_t Q = Q0 + U * kThreads; Q < q_end; Q += U * kThreads) ew_fast_body<T, Op, SRC, U, false>(c, Q, w0, u0);
}

// ---- host-side launcher -----------------------------------------------------------------------
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T>
inline int fill_noise(const tsde_launch* L, const tsde_noise* nz, bool bcast, NoiseP<T>& out) {
  out = NoiseP<T>{};
  out.m = bcast ? 1 : L->m;
  out.bcast = bcast ? 1 : 0;
  if (!nz) return 0;
  out.w = reinterpret_cast<const T*>(nz->w);
  out.u = reinterpret_cast<const T*>(nz->u);
  out.key = nz->key;
  out.cell_id = nz->cell_id;
  out.row_offset = nz->row_offset;
  out.n_cells = nz->n_cells < 1 ? 1 : nz->n_cells;
  out.h = nz->h;
  out.cell_h = nz->cell_h;
  out.h_total = nz->h_total;
  out.sqrt_h = (T)sqrt(nz->h);
  out.sqrt_h12 = (T)sqrt(nz->h / 12.0);
  out.ht = (T)nz->h_total;
  if (nz->source == TSDE_SRC_MEMORY && !nz->w) return TSDE_EINVAL;
  if (nz->source == TSDE_SRC_MEMORY && nz->want_u && !nz->u) return TSDE_EINVAL;
  if (nz->source == TSDE_SRC_COUNTER && !nz->key) return TSDE_EINVAL;
  return 0;
}

template <typename T, typename Op>
inline int launch_ew(const tsde_launch* L, const tsde_noise* nz, bool bcast,
                     const void* const* ins, void* const* outs, const Op& op) {
  EwP<Op::NIN, Op::NOUT> p{};
  if (L->rows == 0) return 0;
  bool vec = (L->d % 4) == 0;
  for (int i = 0; i < Op::NIN; ++i) {
    if (!ins[i]) return TSDE_EINVAL;
    p.in[i] = ins[i];
    vec = vec && aligned16(ins[i]);
  }
  for (int i = 0; i < Op::NOUT; ++i) {
    if (!outs[i]) return TSDE_EINVAL;
    p.out[i] = outs[i];
    vec = vec && aligned16(outs[i]);
  }
  NoiseP<T> np;
  if (int e = fill_noise<T>(L, Op::USES_NOISE ? nz : nullptr, bcast, np)) return e;
  const int src = (Op::USES_NOISE && nz) ? nz->source : TSDE_SRC_UNIT;
  if (Op::USES_NOISE && !nz) return TSDE_EINVAL;
  if (src == TSDE_SRC_MEMORY && !bcast) {
    vec = vec && aligned16(np.w) && (!Op::WANT_U || aligned16(np.u));
    if (L->noise_type == TSDE_NOISE_DIAGONAL && L->m != L->d) return TSDE_EINVAL;
  }
  p.rows = L->rows;
  p.d = L->d;
  p.qpr = (L->d + 3) / 4;
  p.nquads = p.rows * p.qpr;
  p.vec = vec ? (stream_loads_enabled() ? 2 : 1) : 0;
  if (p.nquads == 0) return 0;
  if (L->rows + (nz ? nz->row_offset : 0) > 0xFFFFFFFFll) return TSDE_EINVAL;
  p.qshift = -1;
  if ((p.qpr & (p.qpr - 1)) == 0 && p.qpr < (1ll << 30)) {
    int sh = 0;
    while ((1ll << sh) < p.qpr) ++sh;
    p.qshift = sh;
  }
  p.small = p.nquads < (1ll << 31) ? 1 : 0;
  p.qmagic = ((1ull << 40) + (uint64_t)p.qpr - 1) / (uint64_t)p.qpr;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(L->stream);
  bool pdl = false;
  auto go = [&](auto kernel) -> int {
    // Persistent, balanced grid: one wave of resident CTAs, each owning an equal contiguous slice.
    int per_sm = resident_ctas(kernel);
    if (const int lim = ctas_per_sm_limit(); lim > 0 && lim < per_sm) per_sm = lim;
    const int64_t cap = (int64_t)sm_count() * per_sm;
    int64_t blocks = (p.nquads + kThreads - 1) / kThreads;  // small problems: one quad per thread
    if (blocks > cap) blocks = cap;                          // large: one resident wave, sliced evenly
    if (pdl) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((unsigned)blocks);
      cfg.blockDim = dim3(kThreads);
      cfg.dynamicSmemBytes = 0;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      return (int)cudaLaunchKernelEx(&cfg, kernel, p, np, op);
    }
    kernel<<<(unsigned)blocks, kThreads, 0, st>>>(p, np, op);
    return (int)cudaGetLastError();
  };
  const bool divisible = p.qshift >= 0 || (p.nquads * p.qpr < (1ll << 40) && p.qpr < (1ll << 20));
  const bool fast = p.vec && !bcast && divisible && p.small && np.n_cells == 1 &&
                    (L->rows + (nz ? nz->row_offset : 0)) < 0xFFFFFFFFll;
  pdl = fast && pdl_enabled();
  if constexpr (!Op::USES_NOISE) {
    if (fast) return go(ew_fast_kernel<T, Op, TSDE_SRC_UNIT>);
    return go(ew_kernel<T, Op, TSDE_SRC_UNIT>);
  } else {
    if (fast) {
      switch (src) {
        case TSDE_SRC_MEMORY: return go(ew_fast_kernel<T, Op, TSDE_SRC_MEMORY>);
        case TSDE_SRC_COUNTER: return go(ew_fast_kernel<T, Op, TSDE_SRC_COUNTER>);
        case TSDE_SRC_UNIT: return go(ew_fast_kernel<T, Op, TSDE_SRC_UNIT>);
        default: return TSDE_EINVAL;
      }
    }
    switch (src) {
      case TSDE_SRC_MEMORY:
        return go(ew_kernel<T, Op, TSDE_SRC_MEMORY>);
      case TSDE_SRC_COUNTER:
        if (np.n_cells > 1) return go(ew_kernel<T, Op, kSrcCounterMulti>);
        return go(ew_kernel<T, Op, TSDE_SRC_COUNTER>);
      case TSDE_SRC_UNIT:
        return go(ew_kernel<T, Op, TSDE_SRC_UNIT>);
      default:
        return TSDE_EINVAL;
    }
  }
}

// A launch descriptor every entry point can rely on: non-null, non-negative row count, positive widths.
inline bool launch_invalid(const tsde_launch* L) { return !L || L->rows < 0 || L->d <= 0 || L->m <= 0; }

// (an empty batch is a valid launch that does nothing: its tensors have no storage, so their pointers are null)
#define TSDE_DISPATCH_DTYPE(L, EXPR_F32, EXPR_F64)                                                   \
  (::tsde::launch_invalid(L) ? TSDE_EINVAL                                                           \
   : ((L)->dtype != TSDE_F32 && (L)->dtype != TSDE_F64) ? TSDE_EINVAL                                \
   : (L)->rows == 0          ? 0                                                                     \
   : (L)->dtype == TSDE_F32  ? (EXPR_F32)                                                            \
   : (L)->dtype == TSDE_F64  ? (EXPR_F64)                                                            \
                             : TSDE_EINVAL)

}  // namespace tsde
