// Counter-based Gaussian noise of the fused sim kernel (K_A): Philox4x32-10 + a branch-free fp64
// Box-Muller.  The definition of the noise (oracle/philox_np.py) is unchanged:
//   u1 = ((a >> 11) + 1) 2^-53, u2 = (b >> 11) 2^-53,  n1 + i n2 = sqrt(-2 ln u1) e^{2 pi i u2}
// with a, b the two 64-bit halves of the Philox output.  CUDA's log()/sincospi()/sqrt() cost
// ~290 issued instructions per pair in K_A (49 of them UMOVs that rebuild polynomial constants,
// plus special-case branches that stop the scheduler from interleaving independent pixels); here
// the same quantities take ~130 and the code is one basic block:
//   * -2 ln u1: u1 = m 2^(e-53), m in [1,2); c_i = 1 + i/128 nearest to m (129-entry table in shared
//     memory: {-2/c_i, -2 ln c_i}); w = fma(m, -2/c_i, 2) = -2 (m/c_i - 1), |w| <= 2^-7;
//     -2 ln u1 = (e-53)(-2 ln 2) + (-2 ln c_i) + sum_{k=1..6} w^k / (k 2^(k-1)).  c_0 = 1 and c_128 = 2
//     keep the relative accuracy as u1 -> 1 (the constant terms cancel exactly).
//   * sqrt: rsqrt.approx.f64 seed + one Goldschmidt step + one Newton correction.
//   * e^{2 pi i u2}: the nearest quarter turn comes from the integer bits of b (no range reduction),
//     the remainder |t| <= 1/8 turn goes through degree-6/7 polynomials in t^2 (tools/gen_rng_tables.py:
//     fit errors 4e-17 / 2e-17).
// Agreement with the numpy definition: a few ulp (tests/test_gpu_sim_power.py checks 1e-10 end to end).
#pragma once
#include <cuda_runtime.h>

#include "ox_rng_tables.h"

namespace oxrng {

__constant__ double c_sin[7] = OX_RNG_SIN_COEF;
__constant__ double c_cos[8] = OX_RNG_COS_COEF;
// l(w) = w + w^2 (1/4 + w/12 + w^2/32 + w^3/80 + w^4/192)
__constant__ double c_log[5] = {1.0 / 4, 1.0 / 12, 1.0 / 32, 1.0 / 80, 1.0 / 192};
__device__ const double g_log_table[258] = OX_RNG_LOG_TABLE;

constexpr int LOG_TABLE_ENTRIES = 129;

// copy the log table into shared memory (all threads of the CTA; caller synchronises)
__device__ __forceinline__ void load_log_table(double2 *tab, int tid, int nthreads) {
  for (int i = tid; i < LOG_TABLE_ENTRIES; i += nthreads) tab[i] = make_double2(g_log_table[2 * i], g_log_table[2 * i + 1]);
}

// Philox4x32-10 with the key schedule precomputed (the seed is the same for a whole CTA)
struct PhiloxKeys {
  unsigned k0[10], k1[10];
  __device__ __forceinline__ void init(unsigned long long seed) {
    unsigned a = (unsigned)seed, b = (unsigned)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; r++) {
      k0[r] = a;
      k1[r] = b;
      a += 0x9E3779B9u;
      b += 0xBB67AE85u;
    }
  }
};

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, const PhiloxKeys &k) {
  const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    unsigned hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    unsigned hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.k0[r], lo1, hi0 ^ c.w ^ k.k1[r], lo0);
  }
  return c;
}

// (n1, n2) from the four Philox words (x.x:x.y = a, x.z:x.w = b); tab = shared-memory log table
__device__ __forceinline__ void box_muller_fast(uint4 x, const double2 *__restrict__ tab, double &n1, double &n2) {
  // ---- L = -2 ln u1
  const unsigned long long n = ((((unsigned long long)x.x << 32) | x.y) >> 11) + 1ull;  // 1 .. 2^53
  const double d = (double)n;                                                          // exact
  const int hi = __double2hiint(d);
  const unsigned mant = (unsigned)hi & 0xfffffu;
  const double m = __hiloint2double((int)(mant | 0x3ff00000u), __double2loint(d));
  const double2 t = tab[(mant + 0x1000u) >> 13];
  const double ed = (double)((hi >> 20) - (1023 + 53));
  const double w = fma(m, t.x, 2.0);
  double p = fma(c_log[4], w, c_log[3]);
  p = fma(p, w, c_log[2]);
  p = fma(p, w, c_log[1]);
  p = fma(p, w, c_log[0]);
  double L = fma(ed, OX_RNG_NEG2LN2, t.y) + fma(p, w * w, w);
  // ---- r = sqrt(L).  L = 0 (u1 = 1, probability 2^-53) becomes 1e-300 so that rsqrt stays finite; every
  // other L (>= 2^-52) is unchanged by the addition.  rsqrt.approx seed (2^-22) -> one Goldschmidt step
  // (2^-43) -> one Newton correction with the half-reciprocal h (full precision).
  L += 1e-300;
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(L));
  double g = L * y, h = 0.5 * y;
  const double e = fma(-h, g, 0.5);
  g = fma(g, e, g);
  h = fma(h, e, h);
  g = fma(fma(-g, g, L), h, g);
  // ---- e^{2 pi i u2}: kb = b >> 11 is the angle in units of 2^-53 turn
  const unsigned kb_hi = x.z >> 11, kb_lo = (x.w >> 11) | (x.z << 21);
  const unsigned q = (kb_hi + (1u << 18)) >> 19;  // nearest quarter turn, 0..4
  const int r_hi = (int)kb_hi - (int)(q << 19);
  const long long ri = (long long)(((unsigned long long)(unsigned)r_hi << 32) | kb_lo);
  const double tt = (double)ri * 0x1p-53;  // |tt| <= 1/8 turn
  const double z = tt * tt;
  double s = fma(c_sin[6], z, c_sin[5]);
  double c = fma(c_cos[7], z, c_cos[6]);
  s = fma(s, z, c_sin[4]);
  c = fma(c, z, c_cos[5]);
  s = fma(s, z, c_sin[3]);
  c = fma(c, z, c_cos[4]);
  s = fma(s, z, c_sin[2]);
  c = fma(c, z, c_cos[3]);
  s = fma(s, z, c_sin[1]);
  c = fma(c, z, c_cos[2]);
  s = fma(s, z, c_sin[0]);
  c = fma(c, z, c_cos[1]);
  s *= tt;
  c = fma(c, z, c_cos[0]);
  // quarter turns: q=0 (c,s)  q=1 (-s,c)  q=2 (-c,-s)  q=3 (s,-c)  q=4 = q=0
  const bool swap = q & 1u;
  double co = swap ? s : c, si = swap ? c : s;
  co = __hiloint2double(__double2hiint(co) ^ (int)(((q + 1u) & 2u) << 30), __double2loint(co));
  si = __hiloint2double(__double2hiint(si) ^ (int)((q & 2u) << 30), __double2loint(si));
  n1 = g * co;
  n2 = g * si;
}

// Single-precision variant for the float32 pipeline (tolerance 1e-5): the same random bits and the same
// table-based reduction of u1 (two fp64 operations keep the relative accuracy of -2 ln u1 as u1 -> 1), then
// the series, the square root and the sine/cosine polynomials in fp32 -- the fp64 pipe, which bounds the
// double-precision version, is almost idle here.  Normals agree with the fp64 definition to ~3e-7.
__device__ __forceinline__ void box_muller_fast_f32(uint4 x, const double2 *__restrict__ tab, double &n1, double &n2) {
  const unsigned long long n = ((((unsigned long long)x.x << 32) | x.y) >> 11) + 1ull;
  const double d = (double)n;
  const int hi = __double2hiint(d);
  const unsigned mant = (unsigned)hi & 0xfffffu;
  const double m = __hiloint2double((int)(mant | 0x3ff00000u), __double2loint(d));
  const double2 t = tab[(mant + 0x1000u) >> 13];
  const float w = (float)fma(m, t.x, 2.0);
  const float ed = (float)((hi >> 20) - (1023 + 53));
  const float base = fmaf(ed, (float)OX_RNG_NEG2LN2, (float)t.y);  // cancels exactly for u1 in [1/2 (1 + 255/256), 1)
  const float L = base + fmaf(fmaf(w, 1.0f / 12.0f, 0.25f), w * w, w);
  const float g = __fsqrt_rn(L);
  const unsigned kb_hi = x.z >> 11, kb_lo = (x.w >> 11) | (x.z << 21);
  const unsigned q = (kb_hi + (1u << 18)) >> 19;
  const int r_hi = (int)kb_hi - (int)(q << 19);
  const long long ri = (long long)(((unsigned long long)(unsigned)r_hi << 32) | kb_lo);
  const float tt = (float)ri * 0x1p-53f;  // |tt| <= 1/8 turn
  const float z = tt * tt;
  // Taylor coefficients of sin(2 pi t)/t and cos(2 pi t) in t^2: truncation 2e-9 / 3e-8 at |t| = 1/8
  constexpr float S0 = 6.283185307179586f, S1 = -41.341702240399755f, S2 = 81.60524927607504f, S3 = -76.70585975306136f,
                  S4 = 42.05869394489765f;
  constexpr float C1 = -19.739208802178716f, C2 = 64.93939402266829f, C3 = -85.45681720669373f, C4 = 60.24464137187666f;
  float sn = fmaf(S4, z, S3), cs = fmaf(C4, z, C3);
  sn = fmaf(sn, z, S2);
  cs = fmaf(cs, z, C2);
  sn = fmaf(sn, z, S1);
  cs = fmaf(cs, z, C1);
  sn = fmaf(sn, z, S0) * tt;
  cs = fmaf(cs, z, 1.0f);
  const bool swap = q & 1u;
  float co = swap ? sn : cs, si = swap ? cs : sn;
  co = __int_as_float(__float_as_int(co) ^ (int)(((q + 1u) & 2u) << 30));
  si = __int_as_float(__float_as_int(si) ^ (int)((q & 2u) << 30));
  n1 = (double)(g * co);
  n2 = (double)(g * si);
}

// dispatch on the pipeline's real type
template <typename T>
__device__ __forceinline__ void box_muller(uint4 x, const double2 *__restrict__ tab, double &n1, double &n2) {
  if (sizeof(T) == 4) box_muller_fast_f32(x, tab, n1, n2);
  else box_muller_fast(x, tab, n1, n2);
}

}  // namespace oxrng
