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
//   fp32:  pairs (x_0, x_1) and (x_2, x_3) each give two normals by Box-Muller:
//            a = fp32( fma(float(x_a), 2^-32, 2^-33) )      radius uniform in (0, 1], 32-bit resolution
//            b = (x_b >> 9) * 2^-23                           angle fraction in [0, 1), 23-bit resolution
//            n_even = sqrt(-2 ln a) cos(2 pi b),  n_odd = sqrt(-2 ln a) sin(2 pi b);   normal = n[lane]
//          (evaluated with SFU approximations, see below; the oracle evaluates the same formula in
//           float64 — agreement ~1e-6 absolute per normal)
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
// The Brownian kernels are HBM-bound only if the normals are cheap (ncu, r02: the pure-RNG kernels run at 57-68 %
// issue utilisation with every pipe below 40 %: the instruction COUNT per normal is the limiter), so the fp32 path
//   * uses the SFU (MUFU.LG2 / MUFU.SQRT / MUFU.SIN / MUFU.COS) instead of libm;
//   * radius:  r^2 = -2 ln a.  MUFU.LG2 has ~2^-22 *absolute* error, which would hurt only for a -> 1 (tiny r);
//     there (x >= 0xFF000000, v = 1 - a < 2^-8, exact in fp32)  -ln(1 - v) = v (1 + v/2 + v^2/3 + ...) is used;
//   * angle:   the fraction b is built straight into a float's mantissa, fb = as_float(x >> 9 | 0x3f800000) in
//     [1, 2) — one funnel shift instead of an int->float conversion and a multiply — and
//     theta = 2 pi (fb - 1.5) = 2 pi b - pi lies in [-pi, pi) where sin/cos.approx are accurate (~5e-7 abs);
//     cos(2 pi b) = -cos(theta), sin(2 pi b) = -sin(theta);
//   * evaluates the two pairs of a quad with packed fp32x2 instructions (Blackwell FFMA2 / FMUL2 / FADD2): the
//     arithmetic between the SFU calls costs half the issue slots.
// Agreement with the float64 oracle: a few 1e-7 relative on r, <= ~1e-6 absolute on a normal.
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

// packed fp32x2 helpers (two independent IEEE lanes per instruction; same rounding as the scalar forms)
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// four normals from one Philox output: (x.x, x.y) -> n0, n1 ; (x.z, x.w) -> n2, n3
__device__ __forceinline__ void box_muller4(const uint4 x, float (&n)[4]) {
  const f32x2 k2m32 = pack2(2.3283064365386963e-10f, 2.3283064365386963e-10f);
  const f32x2 k2m33 = pack2(1.1641532182693481e-10f, 1.1641532182693481e-10f);
  // radius uniforms a = fma(float(x), 2^-32, 2^-33), both pairs at once
  const f32x2 a = fma2(pack2(__uint2float_rn(x.x), __uint2float_rn(x.z)), k2m32, k2m33);
  float a0, a1;
  unpack2(a, a0, a1);
  // r^2 through the SFU ...
  const f32x2 via_log = mul2(pack2(mufu_lg2(a0), mufu_lg2(a1)), pack2(-1.3862943611198906f, -1.3862943611198906f));
  // ... and through the series in v = 1 - a (exact), used where a is within 2^-8 of 1
  const f32x2 v = add2(pack2(1.0f, 1.0f), mul2(a, pack2(-1.0f, -1.0f)));
  const f32x2 poly = fma2(v, fma2(v, pack2(0.33333334f, 0.33333334f), pack2(0.5f, 0.5f)), pack2(1.0f, 1.0f));
  const f32x2 series = mul2(add2(v, v), poly);
  float l0, l1, s0, s1;
  unpack2(via_log, l0, l1);
  unpack2(series, s0, s1);
  const float r0 = mufu_sqrt(x.x >= 0xFF000000u ? s0 : l0);
  const float r1 = mufu_sqrt(x.z >= 0xFF000000u ? s1 : l1);
  // angles: mantissa construction, theta = 2 pi (fb - 1.5) in [-pi, pi); fb - 1.5 is exact, so theta carries one
  // rounding and the rounding of the constant (<= 2.7e-7 absolute)
  const f32x2 fb = pack2(__uint_as_float(__funnelshift_r(x.y, 0x7Fu, 9)), __uint_as_float(__funnelshift_r(x.w, 0x7Fu, 9)));
  const f32x2 th = mul2(add2(fb, pack2(-1.5f, -1.5f)), pack2(6.2831853071795865f, 6.2831853071795865f));
  float t0, t1;
  unpack2(th, t0, t1);
  const f32x2 p0 = mul2(pack2(mufu_cos(t0), mufu_sin(t0)), pack2(-r0, -r0));
  const f32x2 p1 = mul2(pack2(mufu_cos(t1), mufu_sin(t1)), pack2(-r1, -r1));
  unpack2(p0, n[0], n[1]);
  unpack2(p1, n[2], n[3]);
}

// four normals for channels 4q..4q+3 of (row, id, stream)
__device__ __forceinline__ void normal4(Key k, uint64_t id, uint32_t stream, uint32_t row,
                                        uint32_t q, float (&n)[4]) {
  const uint4 x = philox4x32_10(
      make_uint4(q | (stream << 24), row, (uint32_t)id, (uint32_t)(id >> 32)), k.lo, k.hi);
  box_muller4(x, n);
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
