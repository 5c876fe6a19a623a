// maps.MapGen.get_map on the device (maps.py:1576-1587).
//
// Reference algorithm: rand = N(0,1)+iN(0,1) on the FULL Fourier plane; k = covsqrt.rand
// (per-pixel ncomp x ncomp mat-vec); [EB->QU rotation]; unitary inverse c2c FFT; .real.
// Taking .real of a c2c inverse equals a c2r inverse of the Hermitian part
//     k_h(p) = 1/2 [ k(p) + conj(k(p')) ],   p' = (-iy mod Ny, -ix mod Nx),
// so one thread per half-plane pixel evaluates k(p) and k(p') -- each with ITS OWN covsqrt
// and rotation (they differ on Nyquist rows/columns and for unsymmetric user covsqrt) --
// and writes k_h(p)/sqrt(Npix); cuFFT Z2D then gives the map with half the bytes of c2c.
//
// Noise sources: OX_NOISE_HOST (numpy's legacy-RNG draws uploaded by the caller: seed
// parity with the reference), OX_NOISE_PHILOX (same algorithm, counter RNG keyed by
// (seed, component, full-plane pixel)), OX_NOISE_PHILOX_HERMITIAN (draws the Hermitian
// noise field directly on the half plane: R(p') = conj(R(p)), |R|^2 has unit mean; half the
// normals, identical statistics).
#include <math.h>

#include "ox_common.cuh"

using namespace ox;

namespace {

constexpr int SIM_THREADS = 256;

// ---- Philox4x32-10 -----------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    unsigned hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    unsigned hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

// two independent N(0,1) from one Philox block: Box-Muller on 53-bit uniforms
__device__ __forceinline__ void philox_normal2(unsigned long long seed, unsigned long long pix, unsigned comp,
                                               unsigned stream, double &n1, double &n2) {
  uint4 c = make_uint4((unsigned)pix, (unsigned)(pix >> 32), comp, stream);
  uint2 k = make_uint2((unsigned)seed, (unsigned)(seed >> 32));
  uint4 x = philox4x32_10(c, k);
  unsigned long long a = ((unsigned long long)x.x << 32) | x.y;
  unsigned long long b = ((unsigned long long)x.z << 32) | x.w;
  double u1 = (double)((a >> 11) + 1ull) * 0x1.0p-53;  // (0,1]
  double u2 = (double)(b >> 11) * 0x1.0p-53;           // [0,1)
  double r = sqrt(-2.0 * log(u1));
  double s, co;
  sincospi(2.0 * u2, &s, &co);
  n1 = r * co;
  n2 = r * s;
}

// cos/sin of a = sgn*2*atan2(-lx, ly) without trigonometry
__device__ __forceinline__ void rot_cs(double y, double x, double sgn, double &c, double &s) {
  double l2 = y * y + x * x;
  c = 1.0;
  s = 0.0;
  if (l2 > 0.0) {
    double inv = 1.0 / l2;
    c = (y * y - x * x) * inv;
    s = sgn * (-2.0 * x * y) * inv;
  }
}

template <typename T, int NC>
struct SimArgs {
  const T *covsqrt;        // [NC][NC][ny][nx]
  const double *noise;     // [nsim][2][NC][ny][nx] (host-noise mode)
  const long long *seeds;  // [nsim]
  const double *ly, *lx;
  int ny, nx, nxh;
  double scale;    // 1/sqrt(Npix)
  double rot_sgn;  // +1 iau, -1 healpix
};

// k(p) = [Rinv(p)] covsqrt(p) r(p) for one full-plane pixel
template <typename T, int NC, bool ROT>
__device__ __forceinline__ void apply_cov(const SimArgs<T, NC> &a, int iy, int ix, const double (&rr)[NC],
                                          const double (&ri)[NC], double (&kr)[NC], double (&ki)[NC]) {
  const long long n = (long long)a.ny * a.nx;
  const long long pix = (long long)iy * a.nx + ix;
#pragma unroll
  for (int i = 0; i < NC; i++) {
    double sr = 0.0, si = 0.0;
#pragma unroll
    for (int j = 0; j < NC; j++) {
      double c = (double)a.covsqrt[(long long)(i * NC + j) * n + pix];
      sr += c * rr[j];
      si += c * ri[j];
    }
    kr[i] = sr;
    ki[i] = si;
  }
  if (ROT && NC == 3) {
    double c, s;
    rot_cs(a.ly[iy], a.lx[ix], a.rot_sgn, c, s);
    // inverse queb_rotmat: [Q;U] = [[c, s],[-s, c]] [E;B]
    double qr = c * kr[1] + s * kr[2], qi = c * ki[1] + s * ki[2];
    double ur = -s * kr[1] + c * kr[2], ui = -s * ki[1] + c * ki[2];
    kr[1] = qr; ki[1] = qi; kr[2] = ur; ki[2] = ui;
  }
}

template <typename T, int NC, int MODE, bool ROT>
__global__ void __launch_bounds__(SIM_THREADS)
sim_fill_kernel(SimArgs<T, NC> a, typename Vec2<T>::type *__restrict__ kh /*[nsim][NC][ny][nxh]*/) {
  typedef typename Vec2<T>::type T2;
  const long long nh = (long long)a.ny * a.nxh;
  const long long n = (long long)a.ny * a.nx;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= nh) return;
  const int sim = blockIdx.y;
  const int iy = (int)(i / a.nxh), ix = (int)(i - (long long)iy * a.nxh);
  const int my = iy ? a.ny - iy : 0, mx = ix ? a.nx - ix : 0;
  const long long p = (long long)iy * a.nx + ix, q = (long long)my * a.nx + mx;
  double pr[NC], pi[NC], qr[NC], qi[NC];
  if (MODE == OX_NOISE_HOST) {
    const double *base = a.noise + (long long)sim * 2 * NC * n;
#pragma unroll
    for (int c = 0; c < NC; c++) {
      pr[c] = base[(long long)c * n + p];
      pi[c] = base[(long long)(NC + c) * n + p];
      qr[c] = base[(long long)c * n + q];
      qi[c] = base[(long long)(NC + c) * n + q];
    }
  } else if (MODE == OX_NOISE_PHILOX) {
    const unsigned long long seed = (unsigned long long)a.seeds[sim];
#pragma unroll
    for (int c = 0; c < NC; c++) {
      philox_normal2(seed, (unsigned long long)p, c, 0u, pr[c], pi[c]);
      if (q == p) {
        qr[c] = pr[c];
        qi[c] = pi[c];
      } else {
        philox_normal2(seed, (unsigned long long)q, c, 0u, qr[c], qi[c]);
      }
    }
  } else {  // Hermitian field drawn on the half plane; canonical pixel = the smaller linear index of {p,p'}
    const unsigned long long seed = (unsigned long long)a.seeds[sim];
    const bool conj_me = q < p;  // only possible in the self-conjugate columns ix==0, 2ix==nx
    const long long canon = conj_me ? q : p;
#pragma unroll
    for (int c = 0; c < NC; c++) {
      double n1, n2;
      philox_normal2(seed, (unsigned long long)canon, c, 1u, n1, n2);
      if (q == p) {
        pr[c] = n1;
        pi[c] = 0.0;
      } else {
        pr[c] = n1 * 0.70710678118654752440;
        pi[c] = (conj_me ? -n2 : n2) * 0.70710678118654752440;
      }
      qr[c] = pr[c];
      qi[c] = -pi[c];
    }
  }
  double kpr[NC], kpi[NC], kqr[NC], kqi[NC];
  apply_cov<T, NC, ROT>(a, iy, ix, pr, pi, kpr, kpi);
  apply_cov<T, NC, ROT>(a, my, mx, qr, qi, kqr, kqi);
  const double h = 0.5 * a.scale;
#pragma unroll
  for (int c = 0; c < NC; c++) {
    T2 z;
    z.x = (T)(h * (kpr[c] + kqr[c]));
    z.y = (T)(h * (kpi[c] - kqi[c]));
    kh[((long long)sim * NC + c) * nh + i] = z;
  }
}

// get_map(harm=True): full-plane covsqrt . rand, no rotation, no FFT (maps.py:1579-1582)
template <typename T, int NC, int MODE>
__global__ void __launch_bounds__(SIM_THREADS)
sim_harm_kernel(SimArgs<T, NC> a, typename Vec2<T>::type *__restrict__ out /*[nsim][NC][ny][nx]*/) {
  typedef typename Vec2<T>::type T2;
  const long long n = (long long)a.ny * a.nx;
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int sim = blockIdx.y;
  const int iy = (int)(p / a.nx), ix = (int)(p - (long long)iy * a.nx);
  double rr[NC], ri[NC], kr[NC], ki[NC];
  if (MODE == OX_NOISE_HOST) {
    const double *base = a.noise + (long long)sim * 2 * NC * n;
#pragma unroll
    for (int c = 0; c < NC; c++) {
      rr[c] = base[(long long)c * n + p];
      ri[c] = base[(long long)(NC + c) * n + p];
    }
  } else if (MODE == OX_NOISE_PHILOX) {
    const unsigned long long seed = (unsigned long long)a.seeds[sim];
#pragma unroll
    for (int c = 0; c < NC; c++) philox_normal2(seed, (unsigned long long)p, c, 0u, rr[c], ri[c]);
  } else {
    const unsigned long long seed = (unsigned long long)a.seeds[sim];
    const int my = iy ? a.ny - iy : 0, mx = ix ? a.nx - ix : 0;
    const long long q = (long long)my * a.nx + mx;
    const bool conj_me = q < p;
    const long long canon = conj_me ? q : p;
#pragma unroll
    for (int c = 0; c < NC; c++) {
      double n1, n2;
      philox_normal2(seed, (unsigned long long)canon, c, 1u, n1, n2);
      if (q == p) {
        rr[c] = n1;
        ri[c] = 0.0;
      } else {
        rr[c] = n1 * 0.70710678118654752440;
        ri[c] = (conj_me ? -n2 : n2) * 0.70710678118654752440;
      }
    }
  }
  apply_cov<T, NC, false>(a, iy, ix, rr, ri, kr, ki);
#pragma unroll
  for (int c = 0; c < NC; c++) {
    T2 z;
    z.x = (T)kr[c];
    z.y = (T)ki[c];
    out[((long long)sim * NC + c) * n + p] = z;
  }
}

template <typename T, int NC>
int launch_fill(ox_simplan *p, int nsim, int mode, const double *noise_dev, int flags, bool harm, void *out) {
  ox_geometry *g = p->g;
  SimArgs<T, NC> a;
  a.covsqrt = p->covsqrt.as<T>();
  a.noise = noise_dev;
  a.seeds = p->seeds.as<long long>();
  a.ly = g->ly.as<double>();
  a.lx = g->lx.as<double>();
  a.ny = g->ny; a.nx = g->nx; a.nxh = g->nxh;
  a.scale = 1.0 / sqrt((double)g->ny * (double)g->nx);
  a.rot_sgn = (flags & OX_FLAG_IAU) ? 1.0 : -1.0;
  typedef typename Vec2<T>::type T2;
  if (harm) {
    long long n = (long long)g->ny * g->nx;
    dim3 grid((unsigned)((n + SIM_THREADS - 1) / SIM_THREADS), nsim);
    if (mode == OX_NOISE_HOST) sim_harm_kernel<T, NC, OX_NOISE_HOST><<<grid, SIM_THREADS, 0, g_stream>>>(a, (T2 *)out);
    else if (mode == OX_NOISE_PHILOX) sim_harm_kernel<T, NC, OX_NOISE_PHILOX><<<grid, SIM_THREADS, 0, g_stream>>>(a, (T2 *)out);
    else sim_harm_kernel<T, NC, OX_NOISE_PHILOX_HERMITIAN><<<grid, SIM_THREADS, 0, g_stream>>>(a, (T2 *)out);
    OX_KERNEL_CHECK();
    return OX_OK;
  }
  long long nh = (long long)g->ny * g->nxh;
  dim3 grid((unsigned)((nh + SIM_THREADS - 1) / SIM_THREADS), nsim);
  const bool rot = (flags & OX_FLAG_ROT) && NC == 3;
#define OX_LAUNCH(M)                                                                                      \
  do {                                                                                                    \
    if (rot) sim_fill_kernel<T, NC, M, true><<<grid, SIM_THREADS, 0, g_stream>>>(a, (T2 *)out);           \
    else sim_fill_kernel<T, NC, M, false><<<grid, SIM_THREADS, 0, g_stream>>>(a, (T2 *)out);              \
  } while (0)
  if (mode == OX_NOISE_HOST) OX_LAUNCH(OX_NOISE_HOST);
  else if (mode == OX_NOISE_PHILOX) OX_LAUNCH(OX_NOISE_PHILOX);
  else OX_LAUNCH(OX_NOISE_PHILOX_HERMITIAN);
#undef OX_LAUNCH
  OX_KERNEL_CHECK();
  return OX_OK;
}

template <typename T>
int dispatch_fill(ox_simplan *p, int nsim, int mode, const double *noise_dev, int flags, bool harm, void *out) {
  switch (p->ncomp) {
    case 1: return launch_fill<T, 1>(p, nsim, mode, noise_dev, flags, harm, out);
    case 2: return launch_fill<T, 2>(p, nsim, mode, noise_dev, flags, harm, out);
    case 3: return launch_fill<T, 3>(p, nsim, mode, noise_dev, flags, harm, out);
    case 4: return launch_fill<T, 4>(p, nsim, mode, noise_dev, flags, harm, out);
  }
  set_error("MapGen: ncomp must be 1..4 (got %d)", p->ncomp);
  return OX_ERR_UNSUPPORTED;
}

int prepare_inputs(ox_simplan *p, const long long *seeds_host, int nsim, int mode, const double *noise, int noise_where,
                   const double **noise_dev) {
  OX_REQUIRE(p, "null plan");
  OX_REQUIRE(nsim >= 1 && nsim <= p->max_batch, "nsim=%d outside 1..max_batch=%d", nsim, p->max_batch);
  OX_REQUIRE(mode >= OX_NOISE_HOST && mode <= OX_NOISE_PHILOX_HERMITIAN, "unknown noise mode %d", mode);
  *noise_dev = nullptr;
  if (mode == OX_NOISE_HOST) {
    OX_REQUIRE(noise, "OX_NOISE_HOST needs the noise array");
    size_t bytes = sizeof(double) * 2 * p->ncomp * (size_t)p->g->ny * p->g->nx * nsim;
    const void *d;
    OX_TRY(stage_in(noise, noise_where, bytes, p->noise, &d));
    *noise_dev = (const double *)d;
  } else {
    OX_REQUIRE(seeds_host, "Philox modes need seeds");
    OX_TRY(p->seeds.ensure(sizeof(long long) * p->max_batch));
    OX_CUDA(cudaMemcpyAsync(p->seeds.p, seeds_host, sizeof(long long) * nsim, cudaMemcpyHostToDevice, g_stream));
  }
  return OX_OK;
}

}  // namespace

namespace ox {

int sim_stage_inputs(ox_simplan *p, const long long *seeds_host, int nsim, int noise_mode, const double *noise,
                     int noise_where, const double **noise_dev) {
  return prepare_inputs(p, seeds_host, nsim, noise_mode, noise, noise_where, noise_dev);
}

int sim_fill_half(ox_simplan *p, const long long *seeds_host, int nsim, int noise_mode, const double *noise,
                  int noise_where, int flags) {
  const double *noise_dev;
  OX_TRY(prepare_inputs(p, seeds_host, nsim, noise_mode, noise, noise_where, &noise_dev));
  OX_REQUIRE(!(flags & OX_FLAG_ROT) || p->ncomp == 3, "EB->QU rotation needs ncomp == 3 (got %d)", p->ncomp);
  size_t es = elem_size(p->dtype);
  OX_TRY(p->kh.ensure(2 * es * (size_t)p->max_batch * p->ncomp * p->g->ny * p->g->nxh));
  if (p->dtype == OX_F64) return dispatch_fill<double>(p, nsim, noise_mode, noise_dev, flags, false, p->kh.p);
  return dispatch_fill<float>(p, nsim, noise_mode, noise_dev, flags, false, p->kh.p);
}

int sim_to_maps(ox_simplan *p, int nsim) {
  size_t es = elem_size(p->dtype);
  OX_TRY(p->maps.ensure(es * (size_t)p->max_batch * p->ncomp * p->g->ny * p->g->nx));
  return p->fft.exec_c2r(nsim * p->ncomp, p->kh.p, p->maps.p);
}

}  // namespace ox

extern "C" {

int ox_simplan_create(ox_geometry *g, int ncomp, const double *covsqrt, int where, int dtype, int max_batch,
                      ox_simplan **out) {
  OX_REQUIRE(g && covsqrt && out, "ox_simplan_create: null pointer");
  OX_REQUIRE(ncomp >= 1 && ncomp <= 4, "MapGen: ncomp must be 1..4 (got %d)", ncomp);
  OX_REQUIRE(dtype == OX_F64 || dtype == OX_F32, "bad dtype %d", dtype);
  OX_REQUIRE(max_batch >= 1, "max_batch must be >= 1");
  ox_simplan *p = new ox_simplan;
  p->g = g;
  p->ncomp = ncomp;
  p->dtype = dtype;
  p->max_batch = max_batch;
  p->fft.ny = g->ny;
  p->fft.nx = g->nx;
  p->fft.dtype = dtype;
  size_t nel = (size_t)ncomp * ncomp * g->ny * g->nx;
  int st;
  auto fail = [&](int s) { delete p; return s; };
  const void *d;
  ox::DevBuf tmp;
  if ((st = stage_in(covsqrt, where, sizeof(double) * nel, tmp, &d)) != OX_OK) return fail(st);
  if ((st = p->covsqrt.ensure(elem_size(dtype) * nel)) != OX_OK) return fail(st);
  if ((st = cast_from_f64((const double *)d, p->covsqrt.p, (long long)nel, dtype)) != OX_OK) return fail(st);
  if (cudaStreamSynchronize(g_stream) != cudaSuccess) { set_error("ox_simplan_create: sync failed"); return fail(OX_ERR_CUDA); }
  *out = p;
  return OX_OK;
}

int ox_simplan_destroy(ox_simplan *p) {
  delete p;
  return OX_OK;
}

int ox_sim_generate(ox_simplan *p, const long long *seeds, int nsim, int noise_mode, const double *noise, int noise_where,
                    int flags, void *out, int out_where) {
  OX_REQUIRE(p && out, "ox_sim_generate: null pointer");
  size_t es = elem_size(p->dtype);
  ox_geometry *g = p->g;
  if (flags & OX_FLAG_HARM) {
    const double *noise_dev;
    OX_TRY(prepare_inputs(p, seeds, nsim, noise_mode, noise, noise_where, &noise_dev));
    size_t bytes = 2 * es * (size_t)nsim * p->ncomp * g->ny * g->nx;
    void *dst = out;
    if (out_where == OX_HOST) {
      OX_TRY(p->stage.ensure(bytes));
      dst = p->stage.p;
    }
    if (p->dtype == OX_F64) OX_TRY(dispatch_fill<double>(p, nsim, noise_mode, noise_dev, flags, true, dst));
    else OX_TRY(dispatch_fill<float>(p, nsim, noise_mode, noise_dev, flags, true, dst));
    if (out_where == OX_HOST) OX_TRY(stage_out(out, OX_HOST, dst, bytes));
    return OX_OK;
  }
  OX_TRY(sim_fill_half(p, seeds, nsim, noise_mode, noise, noise_where, flags));
  OX_TRY(sim_to_maps(p, nsim));
  return stage_out(out, out_where, p->maps.p, es * (size_t)nsim * p->ncomp * g->ny * g->nx);
}

}  // extern "C"
