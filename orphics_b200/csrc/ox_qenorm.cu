// Set-up of the quadratic estimator on the device: the filters W_XY, W_Y and the normalisation A_L of the
// historical QuadNorm (call site tutorials/tt_verification.ipynb:81; arithmetic SURVEY.md Appendix B) from 2-D tables
// that stay in HBM.  The 13 (TT) / 24 (EB) full-plane transforms and the products between them used to cross PCIe
// twice each as pageable complex128 arrays with numpy doing the products; here they are cuFFT c2c transforms and
// three small kernels.  float64 throughout, full planes [ny][nx] (the tables need not be symmetric).
#include "ox_common.cuh"

namespace {
using namespace ox;

int grid_of(long long n, int block) {
  long long want = (n + block - 1) / block, cap = (long long)sm_count() * 32;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

__device__ __forceinline__ double nan_to_num0(double x) { return isfinite(x) ? x : 0.0; }   // np.nan_to_num(x, posinf=0, neginf=0)

__device__ __forceinline__ double modl_of(const double *ly, const double *lx, int iy, int ix) {
  const double y = ly[iy], x = lx[ix];
  return __dsqrt_rn(__dadd_rn(__dmul_rn(y, y), __dmul_rn(x, x)));
}

// W = nan_to_num(num / (lcl B^2 + N)) B, zeroed where mask < 1e-3, L > cut_gt, L >= cut_ge  (num == null: 1)
__global__ void qe_filter_kernel(const double *__restrict__ num, const double *__restrict__ lcl, const double *__restrict__ noise,
                                 const double *__restrict__ beam, const double *__restrict__ mask, const double *__restrict__ ly,
                                 const double *__restrict__ lx, int ny, int nx, double cut_gt, double cut_ge, double *__restrict__ out) {
  const long long n = (long long)ny * nx, stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int iy = (int)(i / nx), ix = (int)(i - (long long)iy * nx);
    const double B = beam ? beam[i] : 1.0;
    const double tot = lcl[i] * (B * B) + (noise ? noise[i] : 0.0);
    double w = nan_to_num0((num ? num[i] : 1.0) / tot) * B;
    const double L = modl_of(ly, lx, iy, ix);
    if (mask && mask[i] < 1.e-3) w = 0.0;
    if (L > cut_gt || L >= cut_ge) w = 0.0;
    out[i] = w;
  }
}

// the planes whose inverse transforms enter A_L, as complex arrays scaled by 1/Npix
//  TT, term t (e1,e2) = (lx,lx), (ly,ly), (r lx, r ly), r = 2^(1/4):  P0 = e1 e2 Cl W1, P1 = e1 W1, P2 = e2 Cl W2;  G0 = W2 (t < 0)
//  EB, term t ellsq = lx^2, ly^2, sqrt2 lx ly:  P_a = ellsq Cl W1 fF_a;  G_b = W2 fG_b (t < 0)
//      fF = (s2^2, c2^2, i sqrt2 s2 c2), fG = (c2^2, s2^2, i sqrt2 s2 c2), s2 = 2 lxhat lyhat, c2 = lyhat^2 - lxhat^2
template <int EST>
__global__ void qe_norm_fill_kernel(const double *__restrict__ cl, const double *__restrict__ w1, const double *__restrict__ w2,
                                    const double *__restrict__ ly, const double *__restrict__ lx, int ny, int nx, int t, double invn,
                                    double2 *__restrict__ out) {
  const long long n = (long long)ny * nx, stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int iy = (int)(i / nx), ix = (int)(i - (long long)iy * nx);
    const double y = ly[iy], x = lx[ix];
    double2 z;
    if (EST == OX_QE_TT) {
      if (t < 0) {
        z.x = w2[i] * invn;
        z.y = 0.0;
        out[i] = z;
        continue;
      }
      const double r = t == 2 ? 1.1892071150027210667 : 1.0;   // 2^(1/4)
      const double e1 = t == 1 ? y : r * x, e2 = t == 0 ? x : r * y;
      z.y = 0.0;
      z.x = e1 * e2 * cl[i] * w1[i] * invn;
      out[i] = z;
      z.x = e1 * w1[i] * invn;
      out[n + i] = z;
      z.x = e2 * cl[i] * w2[i] * invn;
      out[2 * n + i] = z;
    } else {
      const double L = __dsqrt_rn(__dadd_rn(__dmul_rn(y, y), __dmul_rn(x, x)));
      const double inv = nan_to_num0(1.0 / L);
      const double xh = x * inv, yh = y * inv;
      const double s2 = 2.0 * xh * yh, c2 = yh * yh - xh * xh;
      const double sc = 1.4142135623730950488 * s2 * c2;
      double base, f0, f1;
      if (t < 0) {
        base = w2[i] * invn;
        f0 = c2 * c2;
        f1 = s2 * s2;
      } else {
        const double ellsq = t == 0 ? x * x : (t == 1 ? y * y : 1.4142135623730950488 * x * y);
        base = ellsq * cl[i] * w1[i] * invn;
        f0 = s2 * s2;
        f1 = c2 * c2;
      }
      z.y = 0.0;
      z.x = base * f0;
      out[i] = z;
      z.x = base * f1;
      out[n + i] = z;
      z.x = 0.0;
      z.y = base * sc;
      out[2 * n + i] = z;
    }
  }
}

// Q = P0 G0 + P1 P2 (TT)  /  Q = P0 G0 + P1 G1 + P2 G2 (EB)
template <int EST>
__global__ void qe_norm_product_kernel(const double2 *__restrict__ P, const double2 *__restrict__ G, long long n, double2 *__restrict__ Q) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double2 a0 = P[i], a1 = P[n + i], a2 = P[2 * n + i], g0 = G[i];
    double2 q;
    q.x = a0.x * g0.x - a0.y * g0.y;
    q.y = a0.x * g0.y + a0.y * g0.x;
    if (EST == OX_QE_TT) {
      q.x += a1.x * a2.x - a1.y * a2.y;
      q.y += a1.x * a2.y + a1.y * a2.x;
    } else {
      const double2 g1 = G[n + i], g2 = G[2 * n + i];
      q.x += a1.x * g1.x - a1.y * g1.y;
      q.y += a1.x * g1.y + a1.y * g1.x;
      q.x += a2.x * g2.x - a2.y * g2.y;
      q.y += a2.x * g2.y + a2.y * g2.x;
    }
    Q[i] = q;
  }
}

// acc (+)= e Re F, e = e1 e2 (TT) or ellsq (EB) of term t
template <int EST>
__global__ void qe_norm_accumulate_kernel(const double2 *__restrict__ F, const double *__restrict__ ly, const double *__restrict__ lx,
                                          int ny, int nx, int t, double *__restrict__ acc) {
  const long long n = (long long)ny * nx, stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int iy = (int)(i / nx), ix = (int)(i - (long long)iy * nx);
    const double y = ly[iy], x = lx[ix];
    double e;
    if (EST == OX_QE_TT) {
      const double r = t == 2 ? 1.1892071150027210667 : 1.0;
      e = (t == 1 ? y : r * x) * (t == 0 ? x : r * y);
    } else {
      e = t == 0 ? x * x : (t == 1 ? y * y : 1.4142135623730950488 * x * y);
    }
    const double v = e * F[i].x;
    acc[i] = t == 0 ? v : acc[i] + v;
  }
}

// N_L^{kappa kappa} and the multiplier applied in kappa_from_map from 1/A_L^{-1}
__global__ void qe_norm_finish_kernel(const double *__restrict__ acc, const double *__restrict__ kmask, const double *__restrict__ ly,
                                      const double *__restrict__ lx, int ny, int nx, double bigell, double pix_area,
                                      double *__restrict__ nlkk, double *__restrict__ al) {
  const long long n = (long long)ny * nx, stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int iy = (int)(i / nx), ix = (int)(i - (long long)iy * nx);
    const double L = modl_of(ly, lx, iy, ix);
    double alval = nan_to_num0(1.0 / acc[i]);
    if (kmask && kmask[i] < 1.e-3) alval = 0.0;
    double NL = (L * L) * ((L + 1.0) * (L + 1.0)) * alval / 4.0;
    if (L >= bigell || L < 2.0) NL = 0.0;
    double ret = NL * pix_area;
    if (isnan(ret)) ret = 0.0;                      // np.nan_to_num: nan -> 0, +-inf -> +-max
    else if (isinf(ret)) ret = ret > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308;
    nlkk[i] = ret;
    al[i] = ret * 2.0 * nan_to_num0(1.0 / L / (L + 1.0));
  }
}

// max |a(p) - a(p')| over the plane (p' = index-negated pixel), max |a|, and whether the Nyquist row / column hold a non-zero
__global__ void plane_symmetry_kernel(const double *__restrict__ a, int ny, int nx, double *__restrict__ res) {
  const long long n = (long long)ny * nx, stride = (long long)gridDim.x * blockDim.x;
  double dmax = 0.0, amax = 0.0, nyq = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int iy = (int)(i / nx), ix = (int)(i - (long long)iy * nx);
    const int my = iy ? ny - iy : 0, mx = ix ? nx - ix : 0;
    const double v = a[i], d = fabs(v - a[(long long)my * nx + mx]);
    dmax = fmax(dmax, d);
    amax = fmax(amax, fabs(v));
    if ((2 * iy == ny || 2 * ix == nx) && v != 0.0) nyq = 1.0;
    if (isnan(v)) dmax = INFINITY;
  }
  for (int o = 16; o > 0; o >>= 1) {
    dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    nyq = fmax(nyq, __shfl_xor_sync(0xffffffffu, nyq, o));
  }
  if ((threadIdx.x & 31) == 0) {   // values are non-negative: the integer ordering of their bit patterns is the float ordering
    atomicMax(reinterpret_cast<unsigned long long *>(res), (unsigned long long)__double_as_longlong(dmax));
    atomicMax(reinterpret_cast<unsigned long long *>(res + 1), (unsigned long long)__double_as_longlong(amax));
    atomicMax(reinterpret_cast<unsigned long long *>(res + 2), (unsigned long long)__double_as_longlong(nyq));
  }
}

}  // namespace

extern "C" {

// res_host[3] = { max |a(l) - a(-l)|, max |a|, 1 if the Nyquist row or column holds a non-zero else 0 } of a device plane
int ox_plane_symmetry(ox_geometry *g, const double *a_dev, double *res_host) {
  OX_REQUIRE(g && a_dev && res_host, "ox_plane_symmetry: null pointer");
  ox::DevBuf r;
  OX_TRY(r.ensure(3 * sizeof(double)));
  OX_CUDA(cudaMemsetAsync(r.p, 0, 3 * sizeof(double), g_stream));
  const long long n = (long long)g->ny * g->nx;
  plane_symmetry_kernel<<<grid_of(n, 256), 256, 0, g_stream>>>(a_dev, g->ny, g->nx, r.as<double>());
  OX_KERNEL_CHECK();
  OX_CUDA(cudaMemcpyAsync(res_host, r.p, 3 * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  OX_CUDA(cudaStreamSynchronize(g_stream));
  return OX_OK;
}

int ox_qe_filter(ox_geometry *g, const double *num, const double *lcl, const double *noise, const double *beam, const double *mask,
                 double cut_gt, double cut_ge, double *out) {
  OX_REQUIRE(g && lcl && out, "ox_qe_filter: null pointer");
  const long long n = (long long)g->ny * g->nx;
  qe_filter_kernel<<<grid_of(n, 256), 256, 0, g_stream>>>(num, lcl, noise, beam, mask, g->ly.as<double>(), g->lx.as<double>(), g->ny, g->nx,
                                                          cut_gt, cut_ge, out);
  OX_KERNEL_CHECK();
  return OX_OK;
}

int ox_qe_norm(ox_geometry *g, int est, const double *cl, const double *w1, const double *w2, const double *kmask_k, double bigell,
               double pix_area, double *nlkk_out, double *al_out) {
  OX_REQUIRE(g && cl && w1 && w2 && nlkk_out && al_out, "ox_qe_norm: null pointer");
  OX_REQUIRE(est == OX_QE_TT || est == OX_QE_EB, "unknown estimator %d", est);
  const long long n = (long long)g->ny * g->nx;
  const double invn = 1.0 / (double)n;
  FFTPlans fft;
  fft.ny = g->ny;
  fft.nx = g->nx;
  fft.dtype = OX_F64;
  ox::DevBuf G, P, Q, acc;
  const int ng = est == OX_QE_TT ? 1 : 3;
  OX_TRY(G.ensure(sizeof(double2) * n * ng));
  OX_TRY(P.ensure(sizeof(double2) * n * 3));
  OX_TRY(Q.ensure(sizeof(double2) * n));
  OX_TRY(acc.ensure(sizeof(double) * n));
  const int grid = grid_of(n, 256);
  const double *ly = g->ly.as<double>(), *lx = g->lx.as<double>();
#define OX_NORM(EST)                                                                                                            \
  do {                                                                                                                          \
    qe_norm_fill_kernel<EST><<<grid, 256, 0, g_stream>>>(cl, w1, w2, ly, lx, g->ny, g->nx, -1, invn, G.as<double2>());          \
    OX_KERNEL_CHECK();                                                                                                          \
    OX_TRY(fft.exec_c2c(ng, G.p, G.p, CUFFT_INVERSE));                                                                          \
    for (int t = 0; t < 3; t++) {                                                                                               \
      qe_norm_fill_kernel<EST><<<grid, 256, 0, g_stream>>>(cl, w1, w2, ly, lx, g->ny, g->nx, t, invn, P.as<double2>());         \
      OX_KERNEL_CHECK();                                                                                                        \
      OX_TRY(fft.exec_c2c(3, P.p, P.p, CUFFT_INVERSE));                                                                         \
      qe_norm_product_kernel<EST><<<grid, 256, 0, g_stream>>>(P.as<double2>(), G.as<double2>(), n, Q.as<double2>());            \
      OX_KERNEL_CHECK();                                                                                                        \
      OX_TRY(fft.exec_c2c(1, Q.p, Q.p, CUFFT_FORWARD));                                                                         \
      qe_norm_accumulate_kernel<EST><<<grid, 256, 0, g_stream>>>(Q.as<double2>(), ly, lx, g->ny, g->nx, t, acc.as<double>());   \
      OX_KERNEL_CHECK();                                                                                                        \
    }                                                                                                                           \
  } while (0)
  if (est == OX_QE_TT) OX_NORM(OX_QE_TT);
  else OX_NORM(OX_QE_EB);
#undef OX_NORM
  qe_norm_finish_kernel<<<grid, 256, 0, g_stream>>>(acc.as<double>(), kmask_k, ly, lx, g->ny, g->nx, bigell, pix_area, nlkk_out, al_out);
  OX_KERNEL_CHECK();
  OX_CUDA(cudaStreamSynchronize(g_stream));   // the work buffers are freed on return
  return OX_OK;
}

}  // extern "C"
