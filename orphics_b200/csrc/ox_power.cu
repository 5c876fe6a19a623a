// maps.FourierCalc on the device (maps.py:1594-1677): forward FFTs (cuFFT r2c), the
// QU->EB rotation, conj(k1).k2 auto/cross spectra, and the fused power2d+bin2D path.
#include <math.h>

#include "ox_common.cuh"

using namespace ox;

namespace {

constexpr int PW_THREADS = 256;

int grid_1d(long long n, int block) {
  long long want = (n + block - 1) / block;
  long long cap = (long long)ox::sm_count() * 16;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

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

// r2c half plane -> numpy-layout full plane; optional (Q,U)->(E,B) rotation evaluated at the
// FULL-plane pixel (queb_rotmat(lmap) as FourierCalc.__init__ builds it, maps.py:1607)
template <typename T2, int NC, bool ROT>
__global__ void __launch_bounds__(PW_THREADS)
expand_half_kernel(const T2 *__restrict__ kh, T2 *__restrict__ full, const double *__restrict__ ly,
                   const double *__restrict__ lx, int ny, int nx, int nxh, double scale, double rot_sgn) {
  const long long n = (long long)ny * nx, nh = (long long)ny * nxh;
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= n) return;
  const long long m = blockIdx.y;
  const int iy = (int)(p / nx), ix = (int)(p - (long long)iy * nx);
  const bool mirror = ix >= nxh;
  const int sy = mirror ? (iy ? ny - iy : 0) : iy, sx = mirror ? nx - ix : ix;
  const long long src = (long long)sy * nxh + sx;
  double re[NC], im[NC];
#pragma unroll
  for (int c = 0; c < NC; c++) {
    T2 z = kh[(m * NC + c) * nh + src];
    re[c] = (double)z.x * scale;
    im[c] = (mirror ? -(double)z.y : (double)z.y) * scale;
  }
  if (ROT && NC >= 2) {   // the LAST two components, as maps.py:1615 rotates emap[...,-2:,:,:] (ncomp 2: Q,U; 3: I,Q,U)
    constexpr int A = NC >= 2 ? NC - 2 : 0, B = NC >= 2 ? NC - 1 : 0;
    double c, s;
    rot_cs(ly[iy], lx[ix], rot_sgn, c, s);
    double er = c * re[A] - s * re[B], ei = c * im[A] - s * im[B];
    double br = s * re[A] + c * re[B], bi = s * im[A] + c * im[B];
    re[A] = er; im[A] = ei; re[B] = br; im[B] = bi;
  }
#pragma unroll
  for (int c = 0; c < NC; c++) {
    T2 z;
    z.x = re[c];
    z.y = im[c];
    full[(m * NC + c) * n + p] = z;
  }
}

template <typename T>
__global__ void window_kernel(T *__restrict__ maps, const T *__restrict__ window, long long npix, long long nplanes) {
  typedef typename Vec2<T>::type T2;
  const long long half = npix >> 1;  // npix even is checked by the caller for the vector path
  T2 *m2 = reinterpret_cast<T2 *>(maps);
  const T2 *w2 = reinterpret_cast<const T2 *>(window);
  const long long total = half * nplanes;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long j = i % half;
    T2 v = m2[i], w = w2[j];
    v.x *= w.x;
    v.y *= w.y;
    m2[i] = v;
  }
}

template <typename T>
__global__ void window_scalar_kernel(T *__restrict__ maps, const T *__restrict__ window, long long npix, long long nplanes) {
  const long long total = npix * nplanes;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    maps[i] *= window[i % npix];
}

template <typename T2, typename T>
__global__ void f2power_kernel(const T2 *__restrict__ a, const T2 *__restrict__ b, long long n, double norm, T *__restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    T2 x = a[i], y = b[i];
    out[i] = (T)(((double)x.x * (double)y.x + (double)x.y * (double)y.y) * norm);
  }
}

// p2d[m][i][j] = Re(conj(k1_i) k2_j)*norm for i<=j, mirrored (maps.py:1662-1670)
template <typename T2, typename T, int NC>
__global__ void __launch_bounds__(PW_THREADS)
power2d_kernel(const T2 *__restrict__ k1, const T2 *__restrict__ k2, long long n, double norm, int skip_cross,
               T *__restrict__ p2d) {
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= n) return;
  const long long m = blockIdx.y;
  double ar[NC], ai[NC], br[NC], bi[NC];
#pragma unroll
  for (int c = 0; c < NC; c++) {
    T2 x = k1[(m * NC + c) * n + p], y = k2[(m * NC + c) * n + p];
    ar[c] = x.x; ai[c] = x.y; br[c] = y.x; bi[c] = y.y;
  }
  T *o = p2d + m * NC * NC * n + p;
#pragma unroll
  for (int i = 0; i < NC; i++)
#pragma unroll
    for (int j = i; j < NC; j++) {
      double v = (i == j || !skip_cross) ? (ar[i] * br[j] + ai[i] * bi[j]) * norm : 0.0;
      o[(long long)(i * NC + j) * n] = (T)v;
      if (i != j) o[(long long)(j * NC + i) * n] = (T)v;
    }
}

template <typename T2>
__global__ void scale_complex_kernel(T2 *__restrict__ a, long long n, double s) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    T2 z = a[i];
    z.x = z.x * s;
    z.y = z.y * s;
    a[i] = z;
  }
}

template <typename T>
__global__ void cast_in_kernel(const double *__restrict__ in, T *__restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = (T)in[i];
}

template <typename T2, int NC>
int launch_expand(const void *kh, void *full, ox_geometry *g, int nbatch, double scale, int flags) {
  long long n = (long long)g->ny * g->nx;
  dim3 grid((unsigned)((n + PW_THREADS - 1) / PW_THREADS), nbatch);
  double sgn = (flags & OX_FLAG_IAU) ? 1.0 : -1.0;
  if ((flags & OX_FLAG_ROT) && NC >= 2)
    expand_half_kernel<T2, NC, true><<<grid, PW_THREADS, 0, g_stream>>>((const T2 *)kh, (T2 *)full, g->ly.as<double>(),
                                                                       g->lx.as<double>(), g->ny, g->nx, g->nxh, scale, sgn);
  else
    expand_half_kernel<T2, NC, false><<<grid, PW_THREADS, 0, g_stream>>>((const T2 *)kh, (T2 *)full, g->ly.as<double>(),
                                                                        g->lx.as<double>(), g->ny, g->nx, g->nxh, scale, sgn);
  OX_KERNEL_CHECK();
  return OX_OK;
}

int expand(ox_powerplan *p, const void *kh, void *full, int nbatch, double scale, int flags) {
  ox_geometry *g = p->g;
  if (p->dtype == OX_F64) {
    switch (p->ncomp) {
      case 1: return launch_expand<double2, 1>(kh, full, g, nbatch, scale, flags);
      case 2: return launch_expand<double2, 2>(kh, full, g, nbatch, scale, flags);
      case 3: return launch_expand<double2, 3>(kh, full, g, nbatch, scale, flags);
    }
  } else {
    switch (p->ncomp) {
      case 1: return launch_expand<float2, 1>(kh, full, g, nbatch, scale, flags);
      case 2: return launch_expand<float2, 2>(kh, full, g, nbatch, scale, flags);
      case 3: return launch_expand<float2, 3>(kh, full, g, nbatch, scale, flags);
    }
  }
  set_error("FourierCalc: ncomp must be 1..3 (got %d)", p->ncomp);
  return OX_ERR_UNSUPPORTED;
}

// real maps (host or device) -> half-plane Fourier array in `kh`; the maps are copied into
// `in` first because cuFFT may not be handed caller memory it must not overwrite, and so that
// a window can be applied in place
int forward_half(ox_powerplan *p, const void *maps, int where, int nbatch, const void *window_dev, ox::DevBuf &in,
                 ox::DevBuf &kh) {
  ox_geometry *g = p->g;
  size_t es = elem_size(p->dtype);
  long long npix = (long long)g->ny * g->nx;
  size_t nreal = (size_t)nbatch * p->ncomp * npix;
  OX_TRY(in.ensure(es * (size_t)p->max_batch * p->ncomp * npix));
  OX_TRY(kh.ensure(2 * es * (size_t)p->max_batch * p->ncomp * g->ny * g->nxh));
  OX_CUDA(cudaMemcpyAsync(in.p, maps, es * nreal, where == OX_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice,
                          g_stream));
  if (window_dev) OX_TRY(apply_window(p->dtype, in.p, window_dev, npix, (long long)nbatch * p->ncomp));
  return p->fft.exec_r2c(nbatch * p->ncomp, in.p, kh.p);
}

// maps.filter_map (maps.py:1922-1923): Re(ifft(fft(m) * kfilter)) / Npix with a real full-plane kfilter.
// For a real map, taking the real part keeps the Hermitian part of k*f, i.e. k(p) * 1/2 [f(p) + f(p')]:
// one fused multiply between the r2c and c2r transforms, any real filter (beam, l-mask, Wiener).
template <typename T2>
__global__ void filter_half_kernel(T2 *__restrict__ kh, const double *__restrict__ f, int ny, int nx, int nxh, double invn) {
  const long long nh = (long long)ny * nxh;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= nh) return;
  const long long m = blockIdx.y;
  const int iy = (int)(i / nxh), ix = (int)(i - (long long)iy * nxh);
  const int my = iy ? ny - iy : 0, mx = ix ? nx - ix : 0;
  const double w = 0.5 * (f[(long long)iy * nx + ix] + f[(long long)my * nx + mx]) * invn;
  T2 z = kh[m * nh + i];
  z.x = (double)z.x * w;
  z.y = (double)z.y * w;
  kh[m * nh + i] = z;
}

// the same with a complex filter (maps.py:1923 accepts any array): Hermitian part of k f = k(p) 1/2 [f(p) + conj f(p')]
template <typename T2>
__global__ void filter_half_complex_kernel(T2 *__restrict__ kh, const double2 *__restrict__ f, int ny, int nx, int nxh, double invn) {
  const long long nh = (long long)ny * nxh;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= nh) return;
  const long long m = blockIdx.y;
  const int iy = (int)(i / nxh), ix = (int)(i - (long long)iy * nxh);
  const int my = iy ? ny - iy : 0, mx = ix ? nx - ix : 0;
  const double2 a = f[(long long)iy * nx + ix], b = f[(long long)my * nx + mx];
  const double wr = 0.5 * (a.x + b.x) * invn, wi = 0.5 * (a.y - b.y) * invn;
  T2 z = kh[m * nh + i];
  const double zr = (double)z.x, zi = (double)z.y;
  z.x = zr * wr - zi * wi;
  z.y = zr * wi + zi * wr;
  kh[m * nh + i] = z;
}

}  // namespace

namespace ox {

int apply_window(int dtype, void *maps, const void *window, long long npix, long long nplanes) {
  int grid = grid_1d(npix * nplanes / 2 + 1, PW_THREADS);
  if (npix % 2 == 0) {
    if (dtype == OX_F64) window_kernel<double><<<grid, PW_THREADS, 0, g_stream>>>((double *)maps, (const double *)window, npix, nplanes);
    else window_kernel<float><<<grid, PW_THREADS, 0, g_stream>>>((float *)maps, (const float *)window, npix, nplanes);
  } else {
    if (dtype == OX_F64) window_scalar_kernel<double><<<grid, PW_THREADS, 0, g_stream>>>((double *)maps, (const double *)window, npix, nplanes);
    else window_scalar_kernel<float><<<grid, PW_THREADS, 0, g_stream>>>((float *)maps, (const float *)window, npix, nplanes);
  }
  OX_KERNEL_CHECK();
  return OX_OK;
}

}  // namespace ox

extern "C" {

int ox_powerplan_create(ox_geometry *g, int ncomp, int dtype, int max_batch, ox_powerplan **out) {
  OX_REQUIRE(g && out, "ox_powerplan_create: null pointer");
  OX_REQUIRE(ncomp >= 1 && ncomp <= 3, "FourierCalc: ncomp must be 1..3 (got %d)", ncomp);
  OX_REQUIRE(dtype == OX_F64 || dtype == OX_F32, "bad dtype %d", dtype);
  OX_REQUIRE(max_batch >= 1, "max_batch must be >= 1");
  ox_powerplan *p = new ox_powerplan;
  p->g = g;
  p->ncomp = ncomp;
  p->dtype = dtype;
  p->max_batch = max_batch;
  double npix = (double)g->ny * (double)g->nx;
  p->normfact = g->area / (npix * npix);  // maps.py:1605
  p->fft.ny = g->ny;
  p->fft.nx = g->nx;
  p->fft.dtype = dtype;
  *out = p;
  return OX_OK;
}

int ox_powerplan_destroy(ox_powerplan *p) {
  delete p;
  return OX_OK;
}

int ox_power_fft(ox_powerplan *p, const void *maps, int where, int nbatch, int flags, void *kmap_out, int out_where) {
  OX_REQUIRE(p && maps && kmap_out, "ox_power_fft: null pointer");
  OX_REQUIRE(nbatch >= 1 && nbatch <= p->max_batch, "nbatch=%d outside 1..max_batch=%d", nbatch, p->max_batch);
  ox_geometry *g = p->g;
  size_t es = elem_size(p->dtype);
  OX_TRY(forward_half(p, maps, where, nbatch, nullptr, p->in1, p->kh1));
  size_t bytes = 2 * es * (size_t)nbatch * p->ncomp * g->ny * g->nx;
  void *dst = kmap_out;
  if (out_where == OX_HOST) {
    OX_TRY(p->full1.ensure(bytes));
    dst = p->full1.p;
  }
  double scale = (flags & OX_FLAG_UNITARY) ? 1.0 / sqrt((double)g->ny * (double)g->nx) : 1.0;
  OX_TRY(expand(p, p->kh1.p, dst, nbatch, scale, flags));
  if (out_where == OX_HOST) OX_TRY(stage_out(kmap_out, OX_HOST, dst, bytes));
  return OX_OK;
}

int ox_power_ifft(ox_powerplan *p, const void *kmap, int where, int nbatch, void *out, int out_where) {
  OX_REQUIRE(p && kmap && out, "ox_power_ifft: null pointer");
  OX_REQUIRE(nbatch >= 1 && nbatch <= p->max_batch, "nbatch=%d outside 1..max_batch=%d", nbatch, p->max_batch);
  ox_geometry *g = p->g;
  size_t es = elem_size(p->dtype);
  long long n = (long long)nbatch * p->ncomp * g->ny * g->nx;
  size_t bytes = 2 * es * (size_t)n;
  OX_TRY(p->full1.ensure(bytes));
  OX_CUDA(cudaMemcpyAsync(p->full1.p, kmap, bytes, where == OX_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, g_stream));
  OX_TRY(p->fft.exec_c2c(nbatch * p->ncomp, p->full1.p, p->full1.p, CUFFT_INVERSE));
  double s = 1.0 / ((double)g->ny * (double)g->nx);
  if (p->dtype == OX_F64) scale_complex_kernel<double2><<<grid_1d(n, 256), 256, 0, g_stream>>>(p->full1.as<double2>(), n, s);
  else scale_complex_kernel<float2><<<grid_1d(n, 256), 256, 0, g_stream>>>(p->full1.as<float2>(), n, s);
  OX_KERNEL_CHECK();
  return stage_out(out, out_where, p->full1.p, bytes);
}

int ox_power_filter(ox_powerplan *p, const void *maps, int where, int nbatch, const double *kfilter, int kwhere, void *out,
                    int out_where) {
  OX_REQUIRE(p && maps && kfilter && out, "ox_power_filter: null pointer");
  OX_REQUIRE(nbatch >= 1 && nbatch <= p->max_batch, "nbatch=%d outside 1..max_batch=%d", nbatch, p->max_batch);
  ox_geometry *g = p->g;
  size_t es = elem_size(p->dtype);
  long long npix = (long long)g->ny * g->nx, nh = (long long)g->ny * g->nxh;
  const void *fdev;
  OX_TRY(stage_in(kfilter, kwhere, sizeof(double) * npix, p->window, &fdev));
  OX_TRY(forward_half(p, maps, where, nbatch, nullptr, p->in1, p->kh1));
  int planes = nbatch * p->ncomp;
  dim3 grid((unsigned)((nh + PW_THREADS - 1) / PW_THREADS), planes);
  double invn = 1.0 / ((double)g->ny * (double)g->nx);
  if (p->dtype == OX_F64)
    filter_half_kernel<double2><<<grid, PW_THREADS, 0, g_stream>>>(p->kh1.as<double2>(), (const double *)fdev, g->ny, g->nx, g->nxh, invn);
  else
    filter_half_kernel<float2><<<grid, PW_THREADS, 0, g_stream>>>(p->kh1.as<float2>(), (const double *)fdev, g->ny, g->nx, g->nxh, invn);
  OX_KERNEL_CHECK();
  OX_TRY(p->fft.exec_c2r(planes, p->kh1.p, p->in1.p));
  return stage_out(out, out_where, p->in1.p, es * (size_t)planes * npix);
}

int ox_power_filter_complex(ox_powerplan *p, const void *maps, int where, int nbatch, const void *kfilter, int kwhere, void *out,
                            int out_where) {
  OX_REQUIRE(p && maps && kfilter && out, "ox_power_filter_complex: null pointer");
  OX_REQUIRE(nbatch >= 1 && nbatch <= p->max_batch, "nbatch=%d outside 1..max_batch=%d", nbatch, p->max_batch);
  ox_geometry *g = p->g;
  size_t es = elem_size(p->dtype);
  long long npix = (long long)g->ny * g->nx, nh = (long long)g->ny * g->nxh;
  const void *fdev;
  OX_TRY(stage_in(kfilter, kwhere, sizeof(double2) * npix, p->window, &fdev));
  OX_TRY(forward_half(p, maps, where, nbatch, nullptr, p->in1, p->kh1));
  int planes = nbatch * p->ncomp;
  dim3 grid((unsigned)((nh + PW_THREADS - 1) / PW_THREADS), planes);
  double invn = 1.0 / ((double)g->ny * (double)g->nx);
  if (p->dtype == OX_F64)
    filter_half_complex_kernel<double2><<<grid, PW_THREADS, 0, g_stream>>>(p->kh1.as<double2>(), (const double2 *)fdev, g->ny, g->nx, g->nxh, invn);
  else
    filter_half_complex_kernel<float2><<<grid, PW_THREADS, 0, g_stream>>>(p->kh1.as<float2>(), (const double2 *)fdev, g->ny, g->nx, g->nxh, invn);
  OX_KERNEL_CHECK();
  OX_TRY(p->fft.exec_c2r(planes, p->kh1.p, p->in1.p));
  return stage_out(out, out_where, p->in1.p, es * (size_t)planes * npix);
}

int ox_fft_c2c(ox_powerplan *p, const void *in, int where, int nplanes, int direction, double scale, void *out, int out_where) {
  OX_REQUIRE(p && in && out && nplanes >= 1, "ox_fft_c2c: bad arguments");
  OX_REQUIRE(direction == 1 || direction == -1, "direction must be -1 (forward) or +1 (backward)");
  ox_geometry *g = p->g;
  size_t es = elem_size(p->dtype);
  long long n = (long long)nplanes * g->ny * g->nx;
  size_t bytes = 2 * es * (size_t)n;
  OX_TRY(p->full1.ensure(bytes));
  OX_CUDA(cudaMemcpyAsync(p->full1.p, in, bytes, where == OX_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, g_stream));
  OX_TRY(p->fft.exec_c2c(nplanes, p->full1.p, p->full1.p, direction > 0 ? CUFFT_INVERSE : CUFFT_FORWARD));
  if (scale != 1.0) {
    if (p->dtype == OX_F64) scale_complex_kernel<double2><<<grid_1d(n, 256), 256, 0, g_stream>>>(p->full1.as<double2>(), n, scale);
    else scale_complex_kernel<float2><<<grid_1d(n, 256), 256, 0, g_stream>>>(p->full1.as<float2>(), n, scale);
    OX_KERNEL_CHECK();
  }
  return stage_out(out, out_where, p->full1.p, bytes);
}

int ox_power_f2power(ox_powerplan *p, const void *k1, const void *k2, int where, long long n, int flags, void *out, int out_where) {
  OX_REQUIRE(p && k1 && k2 && out && n > 0, "ox_power_f2power: bad arguments");
  size_t es = elem_size(p->dtype);
  const void *d1, *d2;
  OX_TRY(stage_in(k1, where, 2 * es * n, p->full1, &d1));
  if (k2 == k1) d2 = d1;
  else OX_TRY(stage_in(k2, where, 2 * es * n, p->full2, &d2));
  void *dst = out;
  if (out_where == OX_HOST) {
    OX_TRY(p->p2d.ensure(es * n));
    dst = p->p2d.p;
  }
  double norm = (flags & OX_FLAG_PIXEL_UNITS) ? 1.0 : p->normfact;
  if (p->dtype == OX_F64)
    f2power_kernel<double2, double><<<grid_1d(n, 256), 256, 0, g_stream>>>((const double2 *)d1, (const double2 *)d2, n, norm, (double *)dst);
  else
    f2power_kernel<float2, float><<<grid_1d(n, 256), 256, 0, g_stream>>>((const float2 *)d1, (const float2 *)d2, n, norm, (float *)dst);
  OX_KERNEL_CHECK();
  if (out_where == OX_HOST) return stage_out(out, OX_HOST, dst, es * n);
  return OX_OK;
}

int ox_power2d(ox_powerplan *p, const void *maps1, const void *maps2, int where, int nbatch, int flags, void *p2d_out,
               void *kmap1_out, void *kmap2_out, int out_where) {
  OX_REQUIRE(p && maps1 && p2d_out, "ox_power2d: null pointer");
  OX_REQUIRE(nbatch >= 1 && nbatch <= p->max_batch, "nbatch=%d outside 1..max_batch=%d", nbatch, p->max_batch);
  ox_geometry *g = p->g;
  size_t es = elem_size(p->dtype);
  long long n = (long long)g->ny * g->nx;
  size_t kbytes = 2 * es * (size_t)nbatch * p->ncomp * n;
  size_t pbytes = es * (size_t)nbatch * p->ncomp * p->ncomp * n;
  OX_TRY(forward_half(p, maps1, where, nbatch, nullptr, p->in1, p->kh1));
  // device-resident outputs: the full-plane k-maps are expanded straight into the caller's buffers (no copy afterwards)
  void *f1 = (out_where == OX_DEVICE && kmap1_out) ? kmap1_out : nullptr;
  if (!f1) {
    OX_TRY(p->full1.ensure(kbytes));
    f1 = p->full1.p;
  }
  OX_TRY(expand(p, p->kh1.p, f1, nbatch, 1.0, flags));
  const void *f2 = f1;
  if (maps2 && maps2 != maps1) {
    OX_TRY(forward_half(p, maps2, where, nbatch, nullptr, p->in2, p->kh2));
    void *g2 = (out_where == OX_DEVICE && kmap2_out) ? kmap2_out : nullptr;
    if (!g2) {
      OX_TRY(p->full2.ensure(kbytes));
      g2 = p->full2.p;
    }
    OX_TRY(expand(p, p->kh2.p, g2, nbatch, 1.0, flags));
    f2 = g2;
  }
  void *dst = p2d_out;
  if (out_where == OX_HOST) {
    OX_TRY(p->p2d.ensure(pbytes));
    dst = p->p2d.p;
  }
  double norm = (flags & OX_FLAG_PIXEL_UNITS) ? 1.0 : p->normfact;
  int skip = (flags & OX_FLAG_SKIP_CROSS) ? 1 : 0;
  dim3 grid((unsigned)((n + PW_THREADS - 1) / PW_THREADS), nbatch);
#define OX_LAUNCH(T2, T, NC)                                                                                   \
  power2d_kernel<T2, T, NC><<<grid, PW_THREADS, 0, g_stream>>>((const T2 *)f1, (const T2 *)f2, n, norm, skip, (T *)dst)
  if (p->dtype == OX_F64) {
    if (p->ncomp == 1) OX_LAUNCH(double2, double, 1);
    else if (p->ncomp == 2) OX_LAUNCH(double2, double, 2);
    else OX_LAUNCH(double2, double, 3);
  } else {
    if (p->ncomp == 1) OX_LAUNCH(float2, float, 1);
    else if (p->ncomp == 2) OX_LAUNCH(float2, float, 2);
    else OX_LAUNCH(float2, float, 3);
  }
#undef OX_LAUNCH
  OX_KERNEL_CHECK();
  if (out_where == OX_HOST) OX_TRY(stage_out(p2d_out, OX_HOST, dst, pbytes));
  if (kmap1_out) OX_TRY(stage_out(kmap1_out, out_where, f1, kbytes));      // (no-op when f1 already is the caller's buffer)
  if (kmap2_out) OX_TRY(stage_out(kmap2_out, out_where, f2, kbytes));
  return OX_OK;
}

int ox_power_bin(ox_powerplan *p, ox_binner *b, const void *maps1, const void *maps2, int where, int nbatch, int flags,
                 const void *window, int window_where, double *bandpowers, int out_where) {
  OX_REQUIRE(p && b && maps1 && bandpowers, "ox_power_bin: null pointer");
  OX_REQUIRE(nbatch >= 1 && nbatch <= p->max_batch, "nbatch=%d outside 1..max_batch=%d", nbatch, p->max_batch);
  ox_geometry *g = p->g;
  size_t es = elem_size(p->dtype);
  const void *wdev = nullptr;
  if (window) OX_TRY(stage_in(window, window_where, es * (size_t)g->ny * g->nx, p->window, &wdev));
  OX_TRY(forward_half(p, maps1, where, nbatch, wdev, p->in1, p->kh1));
  const void *k2 = nullptr;
  if (maps2 && maps2 != maps1) {
    OX_TRY(forward_half(p, maps2, where, nbatch, wdev, p->in2, p->kh2));
    k2 = p->kh2.p;
  }
  const bool skip = (flags & OX_FLAG_SKIP_CROSS) && p->ncomp > 1;
  int ns = skip ? p->ncomp : p->ncomp * (p->ncomp + 1) / 2;
  int nbins = b->nslots - 2;
  size_t obytes = sizeof(double) * (size_t)nbatch * ns * nbins;
  double *dst = bandpowers;
  if (out_where == OX_HOST) {
    OX_TRY(p->bp.ensure(obytes));
    dst = p->bp.as<double>();
  }
  OX_TRY(power_bin_half(g, b, p->dtype, p->ncomp, p->kh1.p, k2, nbatch, flags, p->normfact, p->partial, dst));
  if (out_where == OX_HOST) OX_TRY(stage_out(bandpowers, OX_HOST, dst, obytes));
  return OX_OK;
}

}  // extern "C"
