// Counter-based normal generator used by every kernel of the library.
//
// Replaces the reference's per-node `torch.Generator(device).manual_seed(seed)` +
// `torch.randn` (torchsde/_brownian/brownian_interval.py:30-32, 243-255) whose seeds come
// from `numpy.random.SeedSequence` (:336-339, :551-552).  That stream is not pinned by any
// reference test, so the bit-level definition below *is* the specification;
// oracle/philox.py restates it in numpy and tests/ pins the two against each other and
// against the Random123 known-answer vectors.
//
// Definition.  For a Brownian object with 64-bit key K, a tree-node / cell id I (64 bit),
// a stream tag S (which normal of the node), global row r and channel c:
//     q    = c / 4,  lane = c % 4
//     ctr  = ( q | S << 24 | call << 31,  r_lo32,  I_lo32,  I_hi32 )        call = 0 (fp32)
//     x[4] = Philox4x32-10(ctr, K)
//   fp32:  u_j = fma(float(x_j), 2^-32, 2^-33)            (round-to-nearest conversion)
//          n0,n1 = BoxMuller(u_0, u_1) ; n2,n3 = BoxMuller(u_2, u_3);  normal = n[lane]
//          (evaluated with SFU approximations, see below; the oracle evaluates the same
//           formula in float64 — agreement ~1e-6 absolute per normal)
//   fp64:  two calls (call = 0,1); call k serves lanes 2k, 2k+1:
//          u_a = ((x_0 * 2^32 + x_1) >> 11 + 0.5) * 2^-53 ; u_b likewise from x_2,x_3
//          n_{2k}, n_{2k+1} = BoxMuller(u_a, u_b)
//   BoxMuller(a, b) = sqrt(-2 ln a) * (cos(2 pi b), sin(2 pi b))
// Rows are independent streams, so sharding the batch over GPUs (row_offset) cannot change
// any trajectory.
#pragma once
#include <stdint.h>

namespace tsde {

enum : uint32_t {
  STREAM_W = 0,   // cell increment normal          (brownian_interval.py:553-554)
  STREAM_H = 1,   // cell space-time Levy normal    (:555-558)
  STREAM_X1 = 2,  // bridge normal X1               (:211)
  STREAM_X2 = 3,  // bridge normal X2               (:212)
  STREAM_A = 4    // Davie/Foster Levy-area noise   (:88, 252-255)
};

struct Key {
  uint32_t lo, hi;
};

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
    k0 += W0;
    k1 += W1;
  }
  return c;
}

// ---- fp32 -------------------------------------------------------------------------------
// The Brownian kernels are HBM-bound only if the normals are cheap, so the fp32 path uses the
// SFU (MUFU.LG2 / MUFU.SQRT / MUFU.SIN / MUFU.COS) instead of libm:
//   * radius:  r^2 = -2 ln u.  MUFU.LG2 has ~2^-22 *absolute* error, which would hurt only for
//     u -> 1 (tiny r); there u = 1 - v with v = (~x + 0.5) 2^-32 exact, and
//     -ln(1 - v) = v (1 + v/2 + v^2/3 + ...) is used instead (x >= 0xFF000000, v < 2^-8).
//   * angle:   theta = 2 pi b - pi in (-pi, pi] where sin/cos.approx are accurate (~5e-7 abs);
//     cos(2 pi b) = -cos(theta), sin(2 pi b) = -sin(theta).
// Agreement with the float64 oracle: a few 1e-7 relative on r, <= ~1e-6 absolute on a normal.
__device__ __forceinline__ float u01(uint32_t x) {
  return fmaf(__uint2float_rn(x), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
}

__device__ __forceinline__ float mufu_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_sqrt(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_sin(float x) {
  float y;
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_cos(float x) {
  float y;
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// a = u01(xa) drives the radius, b = u01(xb) the angle
__device__ __forceinline__ void box_muller(uint32_t xa, uint32_t xb, float& n0, float& n1) {
  const float a = u01(xa);
  const float v = fmaf(__uint2float_rn(~xa), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
  const float series = 2.0f * v * fmaf(v, fmaf(v, 0.33333334f, 0.5f), 1.0f);
  const float viaLog = -1.3862943611198906f * mufu_lg2(a);  // -2 ln2 log2(a)
  const float r2 = xa >= 0xFF000000u ? series : viaLog;
  const float r = mufu_sqrt(r2);
  const float th = fmaf(u01(xb), 6.2831853071795865f, -3.1415926535897932f);
  n0 = -r * mufu_cos(th);
  n1 = -r * mufu_sin(th);
}

// four normals for channels 4q..4q+3 of (row, id, stream)
__device__ __forceinline__ void normal4(Key k, uint64_t id, uint32_t stream, uint32_t row,
                                        uint32_t q, float (&n)[4]) {
  const uint4 x = philox4x32_10(
      make_uint4(q | (stream << 24), row, (uint32_t)id, (uint32_t)(id >> 32)), k.lo, k.hi);
  box_muller(x.x, x.y, n[0], n[1]);
  box_muller(x.z, x.w, n[2], n[3]);
}

// ---- fp64 -------------------------------------------------------------------------------
__device__ __forceinline__ double u01d(uint32_t hi, uint32_t lo) {
  const uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;
  return ((double)v + 0.5) * 1.1102230246251565e-16;  // 2^-53
}

__device__ __forceinline__ void box_muller(double a, double b, double& n0, double& n1) {
  const double r = sqrt(-2.0 * log(a));
  double s, c;
  sincospi(2.0 * b, &s, &c);
  n0 = r * c;
  n1 = r * s;
}

__device__ __forceinline__ void normal4(Key k, uint64_t id, uint32_t stream, uint32_t row,
                                        uint32_t q, double (&n)[4]) {
#pragma unroll
  for (uint32_t call = 0; call < 2; ++call) {
    const uint4 x = philox4x32_10(
        make_uint4(q | (stream << 24) | (call << 31), row, (uint32_t)id, (uint32_t)(id >> 32)),
        k.lo, k.hi);
    box_muller(u01d(x.x, x.y), u01d(x.z, x.w), n[2 * call], n[2 * call + 1]);
  }
}

__device__ __forceinline__ Key load_key(const void* p) {
  const uint2 v = *reinterpret_cast<const uint2*>(p);
  return Key{v.x, v.y};
}

}  // namespace tsde
