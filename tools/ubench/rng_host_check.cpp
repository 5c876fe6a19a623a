// Host mirror of oxrng::box_muller_fast (ox_rng.cuh) to check the algebra, the table and the
// polynomial coefficients against libm on the CPU:  g++ -O2 -I../../orphics_b200/csrc rng_host_check.cpp
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <random>
#include "ox_rng_tables.h"
static const double c_sin[7] = OX_RNG_SIN_COEF, c_cos[8] = OX_RNG_COS_COEF;
static const double c_log[5] = {1.0 / 4, 1.0 / 12, 1.0 / 32, 1.0 / 80, 1.0 / 192};
static const double tab[258] = OX_RNG_LOG_TABLE;
static int hiint(double d) { uint64_t u; memcpy(&u, &d, 8); return (int)(u >> 32); }
static int loint(double d) { uint64_t u; memcpy(&u, &d, 8); return (int)(uint32_t)u; }
static double hilo(int hi, int lo) { uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double d; memcpy(&d, &u, 8); return d; }
static void bm(uint32_t xx, uint32_t xy, uint32_t xz, uint32_t xw, double &n1, double &n2, double &Lout) {
  const uint64_t n = ((((uint64_t)xx << 32) | xy) >> 11) + 1ull;
  const double d = (double)n;
  const int hi = hiint(d);
  const unsigned mant = (unsigned)hi & 0xfffffu;
  const double m = hilo((int)(mant | 0x3ff00000u), loint(d));
  const unsigned idx = (mant + 0x1000u) >> 13;
  const double t0 = tab[2 * idx], t1 = tab[2 * idx + 1];
  const double ed = (double)((hi >> 20) - (1023 + 53));
  const double w = fma(m, t0, 2.0);
  double p = fma(c_log[4], w, c_log[3]);
  p = fma(p, w, c_log[2]); p = fma(p, w, c_log[1]); p = fma(p, w, c_log[0]);
  double L = fma(ed, OX_RNG_NEG2LN2, t1) + fma(p, w * w, w);
  Lout = L;
  L += 1e-300;
  double y = hilo(hiint((double)(1.0f / sqrtf((float)L))), 0);   // ~22-bit seed, low word zero
  if (L < 1e-30) y = hilo(hiint(1.0 / sqrt(L)), 0);
  y *= (1.0 + 2.3e-7);                                            // worst-case seed error 2^-22
  double g = L * y, h = 0.5 * y;
  const double e = fma(-h, g, 0.5);
  g = fma(g, e, g); h = fma(h, e, h);
  g = fma(fma(-g, g, L), h, g);
  const unsigned kb_hi = xz >> 11, kb_lo = (xw >> 11) | (xz << 21);
  const unsigned q = (kb_hi + (1u << 18)) >> 19;
  const int r_hi = (int)kb_hi - (int)(q << 19);
  const long long ri = (long long)(((unsigned long long)(unsigned)r_hi << 32) | kb_lo);
  const double tt = (double)ri * 0x1p-53;
  const double z = tt * tt;
  double s = fma(c_sin[6], z, c_sin[5]);
  double c = fma(c_cos[7], z, c_cos[6]);
  for (int k = 4; k >= 0; k--) s = fma(s, z, c_sin[k]);
  for (int k = 5; k >= 0; k--) c = fma(c, z, c_cos[k]);
  s *= tt;
  const bool swap = q & 1u;
  double co = swap ? s : c, si = swap ? c : s;
  co = hilo(hiint(co) ^ (int)(((q + 1u) & 2u) << 30), loint(co));
  si = hilo(hiint(si) ^ (int)((q & 2u) << 30), loint(si));
  n1 = g * co; n2 = g * si;
}
// host mirror of oxrng::box_muller_fast_f32 (single-precision tail of the float32 pipeline)
static void bm32(uint32_t xx, uint32_t xy, uint32_t xz, uint32_t xw, double &n1, double &n2) {
  const uint64_t n = ((((uint64_t)xx << 32) | xy) >> 11) + 1ull;
  const double d = (double)n;
  const int hi = hiint(d);
  const unsigned mant = (unsigned)hi & 0xfffffu;
  const double m = hilo((int)(mant | 0x3ff00000u), loint(d));
  const unsigned idx = (mant + 0x1000u) >> 13;
  const float w = (float)fma(m, tab[2 * idx], 2.0);
  const float ed = (float)((hi >> 20) - (1023 + 53));
  const float base = fmaf(ed, (float)OX_RNG_NEG2LN2, (float)tab[2 * idx + 1]);
  const float L = base + fmaf(fmaf(w, 1.0f / 12.0f, 0.25f), w * w, w);
  const float g = sqrtf(L);
  const unsigned kb_hi = xz >> 11, kb_lo = (xw >> 11) | (xz << 21);
  const unsigned q = (kb_hi + (1u << 18)) >> 19;
  const int r_hi = (int)kb_hi - (int)(q << 19);
  const long long ri = (long long)(((unsigned long long)(unsigned)r_hi << 32) | kb_lo);
  const float tt = (float)ri * 0x1p-53f;
  const float z = tt * tt;
  const float S0 = 6.283185307179586f, S1 = -41.341702240399755f, S2 = 81.60524927607504f, S3 = -76.70585975306136f,
              S4 = 42.05869394489765f;
  const float C1 = -19.739208802178716f, C2 = 64.93939402266829f, C3 = -85.45681720669373f, C4 = 60.24464137187666f;
  float sn = fmaf(S4, z, S3), cs = fmaf(C4, z, C3);
  sn = fmaf(sn, z, S2); cs = fmaf(cs, z, C2);
  sn = fmaf(sn, z, S1); cs = fmaf(cs, z, C1);
  sn = fmaf(sn, z, S0) * tt; cs = fmaf(cs, z, 1.0f);
  const bool swap = q & 1u;
  float co = swap ? sn : cs, si = swap ? cs : sn;
  if ((q + 1u) & 2u) co = -co;
  if (q & 2u) si = -si;
  n1 = (double)(g * co); n2 = (double)(g * si);
}

int main(int argc, char **argv) {
  const long niter = argc > 1 ? atol(argv[1]) : 20000000;
  std::mt19937_64 gen(12345);
  double worst_n = 0, worst_L = 0, worst_32 = 0;
  for (long it = 0; it < niter; it++) {
    uint64_t a = gen(), b = gen();
    if (it < 64) a = ~0ull << (it);                 // u1 near 1 ... small
    if (it >= 64 && it < 128) a = (1ull << (it - 64)) - 1; // tiny u1
    if (it >= 128 && it < 200) b = (uint64_t)(it - 128) << 58;  // quadrant boundaries
    double n1, n2, L;
    bm(a >> 32, (uint32_t)a, b >> 32, (uint32_t)b, n1, n2, L);
    long double u1 = (long double)((a >> 11) + 1) * 0x1p-53L, u2 = (long double)(b >> 11) * 0x1p-53L;
    long double Lr = -2.0L * logl(u1), r = sqrtl(Lr);
    long double e1 = r * cosl(6.283185307179586476925286766559L * u2), e2 = r * sinl(6.283185307179586476925286766559L * u2);
    double dn = fmax(fabs((double)(n1 - e1)), fabs((double)(n2 - e2)));
    double dL = Lr > 0 ? fabs((double)((L - Lr) / Lr)) : fabs(L);
    if (dn > worst_n) worst_n = dn;
    double m1, m2;
    bm32(a >> 32, (uint32_t)a, b >> 32, (uint32_t)b, m1, m2);
    double dm = fmax(fabs((double)(m1 - e1)), fabs((double)(m2 - e2)));
    if (dm > worst_32) worst_32 = dm;
    if (dL > worst_L) { worst_L = dL; }
  }
  printf("max abs err of normals %.3g, max rel err of -2 ln u1 %.3g, max abs err of the float32 variant %.3g\n", worst_n, worst_L,
         worst_32);
  return !(worst_n < 5e-15 && worst_L < 1e-13 && worst_32 < 5e-6);
}
