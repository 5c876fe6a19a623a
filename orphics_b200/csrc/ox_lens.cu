// Lensing of flat-sky maps on the device: kappa -> phi (lensing.py:651-665), the Taylens remap + Taylor series
// (lensing.py:395-440) and a bicubic displacement standing in for pixell.lensing.displace_map (lensing.py:512), so
// that FlatLensingSims.get_sim (lensing.py:499-521) has no host arithmetic between its seeds and `observed`.
//
// Every Fourier array here is the transform of a real map times a factor, and the reference takes np.real of a
// full-plane c2c inverse: that keeps the Hermitian part 1/2 [F(p) + conj F(p')] (p' = index-negated pixel).  For
// F(p) = i^n c lx^a ly^b k(p) with Hermitian k the Hermitian part is F itself, except on the Nyquist column
// (row) where lx (ly) does not change sign under p -> p': there it is F for even a (b) and 0 for odd a (b), and
// on the Nyquist corner F for even n, 0 for odd n.  With that factor the half-plane c2r transform is exactly
// the reference's Re ifft, at half the bytes.  The 'phys' normalisations of enmap.fft / ifft cancel to 1/Npix.
#include "ox_common.cuh"

struct ox_lensplan {
  ox_geometry *g = nullptr;
  double py = 0, px = 0;
  int max_planes = 0;
  FFTPlans plans;
  ox::DevBuf in;      // staged real input [max_planes][ny][nx]
  ox::DevBuf kh;      // half-plane transform of the input(s) [max_planes][ny][nxh]
  ox::DevBuf dh;      // half-plane derivative planes [max_planes][ny][nxh]
  ox::DevBuf dr;      // real derivative planes [max_planes][ny][nx]
  ox::DevBuf alpha;   // [2][ny][nx]: alphaY, alphaX (radians)
  ox::DevBuf src;     // int32[ny*nx]: flat index of the pixel the nearest-pixel remap reads
  ox::DevBuf dxy;     // double2[ny*nx]: sub-pixel remainder (dX, dY) in radians
  ox::DevBuf out;     // staged result [max_planes][ny][nx]
  bool have_phi = false;
};

namespace {
using namespace ox;

int grid_of(long long n, int block) {
  long long want = (n + block - 1) / block, cap = (long long)sm_count() * 32;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

// weight of the Hermitian part of i^n lx^a ly^b k(p) relative to the term itself (see the header): 1 or 0
__device__ __forceinline__ double herm_weight(int iy, int ix, int ny, int nx, int n, int a, int b) {
  const bool nyq_y = (2 * iy == ny), nyq_x = (2 * ix == nx);
  if (nyq_x && nyq_y) return (n & 1) ? 0.0 : 1.0;
  if (nyq_x && (a & 1)) return 0.0;
  if (nyq_y && (b & 1)) return 0.0;
  return 1.0;
}

__device__ __forceinline__ double ipow(double x, int e) {
  double r = 1.0;
  for (int i = 0; i < e; i++) r *= x;
  return r;
}

// out[k][p] = i^n binom(n,k)/n! lx^(n-k) ly^k kmap[p] / Npix (Hermitian part), k = 0..n, half plane
__global__ void deriv_fill_kernel(const double2 *__restrict__ kh, const double *__restrict__ ly, const double *__restrict__ lx,
                                  int ny, int nx, int nxh, int n, double invn, double2 *__restrict__ out) {
  const long long nh = (long long)ny * nxh;
  const long long stride = (long long)gridDim.x * blockDim.x;
  double fact = 1.0;
  for (int i = 2; i <= n; i++) fact *= i;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nh; i += stride) {
    const int iy = (int)(i / nxh), ix = (int)(i - (long long)iy * nxh);
    const double2 z = kh[i];
    const double y = ly[iy], x = lx[ix];
    double binom = 1.0;
    for (int k = 0; k <= n; k++) {
      if (k > 0) binom = binom * (double)(n - (k - 1)) / (double)k;   // exact for the small n used
      const double f = binom * ipow(x, n - k) * ipow(y, k) / fact * invn * herm_weight(iy, ix, ny, nx, n, n - k, k);
      // i^n (zr + i zi): n mod 4 = 0: z, 1: i z, 2: -z, 3: -i z
      double2 r;
      switch (n & 3) {
        case 0: r.x = z.x; r.y = z.y; break;
        case 1: r.x = -z.y; r.y = z.x; break;
        case 2: r.x = -z.x; r.y = -z.y; break;
        default: r.x = z.y; r.y = -z.x; break;
      }
      r.x *= f;
      r.y *= f;
      out[(long long)k * nh + i] = r;
    }
  }
}

// phi(l) = 2 kappa(l) / L / (L + 1) / Npix, zero for L < 2 (lensing.py:662-665), half plane, in place
__global__ void kappa_to_phi_kernel(double2 *__restrict__ kh, const double *__restrict__ ly, const double *__restrict__ lx, int ny,
                                    int nxh, double invn) {
  const long long nh = (long long)ny * nxh;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nh; i += stride) {
    const int iy = (int)(i / nxh), ix = (int)(i - (long long)iy * nxh);
    const double y = ly[iy], x = lx[ix];
    const double L = __dsqrt_rn(__dadd_rn(__dmul_rn(y, y), __dmul_rn(x, x)));   // = modlmap, bit for bit
    double2 z = kh[i];
    if (L < 2.0) {
      z.x = 0.0;
      z.y = 0.0;
    } else {
      z.x = 2.0 * z.x / L / (L + 1.0) * invn;
      z.y = 2.0 * z.y / L / (L + 1.0) * invn;
    }
    kh[i] = z;
  }
}

// deflection split of Taylens (lensing.py:421-427): nearest pixel + remainder
__global__ void split_alpha_kernel(const double *__restrict__ alphaY, const double *__restrict__ alphaX, int ny, int nx, double py,
                                   double px, int *__restrict__ src, double2 *__restrict__ dxy) {
  const long long n = (long long)ny * nx;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int iy = (int)(i / nx), ix = (int)(i - (long long)iy * nx);
    const double ax = alphaX[i], ay = alphaY[i];
    const double rx = rint(ax / px), ry = rint(ay / py);   // np.round: half to even
    double2 d;
    d.x = ax - rx * px;
    d.y = ay - ry * py;
    long long sy = ((long long)iy + (long long)ry) % ny, sx = ((long long)ix + (long long)rx) % nx;
    if (sy < 0) sy += ny;
    if (sx < 0) sx += nx;
    src[i] = (int)(sy * nx + sx);
    dxy[i] = d;
  }
}

// lensed = imap[src]  (order 0)  /  lensed += sum_k deriv[k][src] dX^(n-k) dY^k  (order n >= 1)
__global__ void taylens_gather_kernel(const double *__restrict__ planes, const int *__restrict__ src,
                                      const double2 *__restrict__ dxy, long long npix, int n, double *__restrict__ lensed) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += stride) {
    const int s = src[i];
    if (n == 0) {
      lensed[i] = planes[s];
      continue;
    }
    const double2 d = dxy[i];
    double acc = lensed[i];
    for (int k = 0; k <= n; k++) acc += planes[(long long)k * npix + s] * ipow(d.x, n - k) * ipow(d.y, k);   // the reference's order of terms
    lensed[i] = acc;
  }
}

// Keys cubic convolution weights (a = -1/2) for the fractional offset t in [0,1): taps at -1, 0, 1, 2
__device__ __forceinline__ void keys_weights(double t, double *w) {
  const double t2 = t * t, t3 = t2 * t;
  w[0] = -0.5 * t3 + t2 - 0.5 * t;
  w[1] = 1.5 * t3 - 2.5 * t2 + 1.0;
  w[2] = -1.5 * t3 + 2.0 * t2 + 0.5 * t;
  w[3] = 0.5 * t3 - 0.5 * t2;
}

// out[p] = imap interpolated at pixel coordinates (iy + alphaY/py, ix + alphaX/px), periodic, bicubic
__global__ void displace_bicubic_kernel(const double *__restrict__ imap, const double *__restrict__ alphaY,
                                        const double *__restrict__ alphaX, int ny, int nx, double py, double px,
                                        double *__restrict__ out) {
  const long long n = (long long)ny * nx;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long m = blockIdx.y;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int iy = (int)(i / nx), ix = (int)(i - (long long)iy * nx);
    const double fy = (double)iy + alphaY[i] / py, fx = (double)ix + alphaX[i] / px;
    const double by = floor(fy), bx = floor(fx);
    double wy[4], wx[4];
    keys_weights(fy - by, wy);
    keys_weights(fx - bx, wx);
    long long y0 = ((long long)by - 1) % ny, x0 = ((long long)bx - 1) % nx;
    if (y0 < 0) y0 += ny;
    if (x0 < 0) x0 += nx;
    double acc = 0.0;
    for (int jy = 0; jy < 4; jy++) {
      long long yy = y0 + jy;
      if (yy >= ny) yy -= ny;
      const double *rowp = imap + m * n + yy * nx;
      double r = 0.0;
      for (int jx = 0; jx < 4; jx++) {
        long long xx = x0 + jx;
        if (xx >= nx) xx -= nx;
        r += wx[jx] * rowp[xx];
      }
      acc += wy[jy] * r;
    }
    out[m * n + i] = acc;
  }
}

}  // namespace

extern "C" {

int ox_lensplan_create(ox_geometry *g, double py, double px, int max_planes, ox_lensplan **out) {
  OX_REQUIRE(g && out, "ox_lensplan_create: null pointer");
  OX_REQUIRE(py != 0.0 && px != 0.0 && max_planes >= 2, "ox_lensplan_create: pixel shape (%g, %g), max_planes %d", py, px, max_planes);
  OX_REQUIRE((long long)g->ny * g->nx < (1LL << 31), "lensing: the map is too large for 32-bit source indices");
  ox_lensplan *p = new ox_lensplan;
  p->g = g;
  p->py = py;
  p->px = px;
  p->max_planes = max_planes;
  p->plans.ny = g->ny;
  p->plans.nx = g->nx;
  p->plans.dtype = OX_F64;
  *out = p;
  return OX_OK;
}

int ox_lensplan_destroy(ox_lensplan *p) {
  delete p;
  return OX_OK;
}

// kappa -> phi (lensing.py:651-657): phi = Re ifft( 2 fft(kappa) / (L (L+1)) ), zero for L < 2
int ox_lens_kappa_to_phi(ox_lensplan *p, const double *kappa, int where, double *phi_out, int out_where) {
  OX_REQUIRE(p && kappa && phi_out, "ox_lens_kappa_to_phi: null pointer");
  ox_geometry *g = p->g;
  const long long npix = (long long)g->ny * g->nx, nh = (long long)g->ny * g->nxh;
  OX_TRY(p->in.ensure(sizeof(double) * npix));
  OX_TRY(p->kh.ensure(sizeof(double2) * nh));
  OX_CUDA(cudaMemcpyAsync(p->in.p, kappa, sizeof(double) * npix, where == OX_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, g_stream));
  OX_TRY(p->plans.exec_r2c(1, p->in.p, p->kh.p));
  kappa_to_phi_kernel<<<grid_of(nh, 256), 256, 0, g_stream>>>(p->kh.as<double2>(), g->ly.as<double>(), g->lx.as<double>(), g->ny, g->nxh,
                                                              1.0 / (double)npix);
  OX_KERNEL_CHECK();
  OX_TRY(p->plans.exec_c2r(1, p->kh.p, p->in.p));
  return stage_out(phi_out, out_where, p->in.p, sizeof(double) * npix);
}

// the deflection field of a lensing potential, alpha = grad phi by FFT (lensing.py:414-419), split into the
// nearest-pixel remap and the sub-pixel remainder (lensing.py:421-427); kept in the plan for the calls below
int ox_lens_set_phi(ox_lensplan *p, const double *phi, int where) {
  OX_REQUIRE(p && phi, "ox_lens_set_phi: null pointer");
  ox_geometry *g = p->g;
  const long long npix = (long long)g->ny * g->nx, nh = (long long)g->ny * g->nxh;
  OX_TRY(p->in.ensure(sizeof(double) * npix * p->max_planes));
  OX_TRY(p->kh.ensure(sizeof(double2) * nh * p->max_planes));
  OX_TRY(p->dh.ensure(sizeof(double2) * nh * p->max_planes));
  OX_TRY(p->alpha.ensure(sizeof(double) * npix * 2));
  OX_TRY(p->src.ensure(sizeof(int) * npix));
  OX_TRY(p->dxy.ensure(sizeof(double2) * npix));
  OX_CUDA(cudaMemcpyAsync(p->in.p, phi, sizeof(double) * npix, where == OX_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, g_stream));
  OX_TRY(p->plans.exec_r2c(1, p->in.p, p->kh.p));
  // order-1 derivative planes: k = 0 -> i lx kphi (alphaX), k = 1 -> i ly kphi (alphaY)
  deriv_fill_kernel<<<grid_of(nh, 256), 256, 0, g_stream>>>(p->kh.as<double2>(), g->ly.as<double>(), g->lx.as<double>(), g->ny, g->nx,
                                                            g->nxh, 1, 1.0 / (double)npix, p->dh.as<double2>());
  OX_KERNEL_CHECK();
  // alpha buffer is [alphaX, alphaY] in the order the planes come out
  OX_TRY(p->plans.exec_c2r(2, p->dh.p, p->alpha.p));
  split_alpha_kernel<<<grid_of(npix, 256), 256, 0, g_stream>>>(p->alpha.as<double>() + npix, p->alpha.as<double>(), g->ny, g->nx, p->py,
                                                               p->px, p->src.as<int>(), p->dxy.as<double2>());
  OX_KERNEL_CHECK();
  p->have_phi = true;
  return OX_OK;
}

// alpha_out: [2][ny][nx] = (alphaY, alphaX) in radians, the deflection of the current potential
int ox_lens_alpha(ox_lensplan *p, double *alpha_out, int out_where) {
  OX_REQUIRE(p && alpha_out && p->have_phi, "ox_lens_alpha: no potential set");
  const long long npix = (long long)p->g->ny * p->g->nx;
  OX_TRY(stage_out(alpha_out, out_where, p->alpha.as<double>() + npix, sizeof(double) * npix));
  return stage_out(alpha_out + npix, out_where, p->alpha.p, sizeof(double) * npix);
}

// Taylens (lensing.py:395-440) of nmaps maps [nmaps][ny][nx] by the current potential
int ox_lens_taylens(ox_lensplan *p, const double *imap, int where, int nmaps, int taylor_order, double *out, int out_where) {
  OX_REQUIRE(p && imap && out && p->have_phi, "ox_lens_taylens: null pointer or no potential set");
  OX_REQUIRE(nmaps >= 1 && taylor_order >= 1 && taylor_order <= p->max_planes, "ox_lens_taylens: nmaps=%d, taylor_order=%d (plan allows %d)",
             nmaps, taylor_order, p->max_planes);
  ox_geometry *g = p->g;
  const long long npix = (long long)g->ny * g->nx, nh = (long long)g->ny * g->nxh;
  OX_TRY(p->dr.ensure(sizeof(double) * npix * p->max_planes));
  OX_TRY(p->out.ensure(sizeof(double) * npix));
  OX_TRY(p->in.ensure(sizeof(double) * npix * p->max_planes));
  ox::DevBuf stage;
  const void *src_dev = nullptr;
  OX_TRY(stage_in(imap, where, sizeof(double) * npix * nmaps, stage, &src_dev));
  for (int m = 0; m < nmaps; m++) {
    const double *mp = (const double *)src_dev + (long long)m * npix;
    double *dst = out_where == OX_DEVICE ? out + (long long)m * npix : p->out.as<double>();
    taylens_gather_kernel<<<grid_of(npix, 256), 256, 0, g_stream>>>(mp, p->src.as<int>(), p->dxy.as<double2>(), npix, 0, dst);
    OX_KERNEL_CHECK();
    if (taylor_order > 1) {
      OX_CUDA(cudaMemcpyAsync(p->in.p, mp, sizeof(double) * npix, cudaMemcpyDeviceToDevice, g_stream));   // (cuFFT may overwrite its input)
      OX_TRY(p->plans.exec_r2c(1, p->in.p, p->kh.p));
    }
    for (int n = 1; n < taylor_order; n++) {
      deriv_fill_kernel<<<grid_of(nh, 256), 256, 0, g_stream>>>(p->kh.as<double2>(), g->ly.as<double>(), g->lx.as<double>(), g->ny,
                                                                g->nx, g->nxh, n, 1.0 / (double)npix, p->dh.as<double2>());
      OX_KERNEL_CHECK();
      OX_TRY(p->plans.exec_c2r(n + 1, p->dh.p, p->dr.p));
      taylens_gather_kernel<<<grid_of(npix, 256), 256, 0, g_stream>>>(p->dr.as<double>(), p->src.as<int>(), p->dxy.as<double2>(), npix, n,
                                                                      dst);
      OX_KERNEL_CHECK();
    }
    if (out_where == OX_HOST) OX_TRY(stage_out(out + (long long)m * npix, OX_HOST, dst, sizeof(double) * npix));
  }
  return OX_OK;
}

// bicubic (Keys, a = -1/2) interpolation of nmaps maps at the displaced positions, periodic boundaries: the
// stand-in for pixell.lensing.displace_map(imap, alpha_pix, order) at FlatLensingSims.get_sim (lensing.py:512)
int ox_lens_displace(ox_lensplan *p, const double *imap, int where, int nmaps, double *out, int out_where) {
  OX_REQUIRE(p && imap && out && p->have_phi, "ox_lens_displace: null pointer or no potential set");
  OX_REQUIRE(nmaps >= 1, "ox_lens_displace: nmaps=%d", nmaps);
  ox_geometry *g = p->g;
  const long long npix = (long long)g->ny * g->nx;
  ox::DevBuf stage, ostage;
  const void *src_dev = nullptr;
  OX_TRY(stage_in(imap, where, sizeof(double) * npix * nmaps, stage, &src_dev));
  double *dst = out;
  if (out_where == OX_HOST) {
    OX_TRY(ostage.ensure(sizeof(double) * npix * nmaps));
    dst = ostage.as<double>();
  }
  dim3 grid(grid_of(npix, 256), nmaps);
  displace_bicubic_kernel<<<grid, 256, 0, g_stream>>>((const double *)src_dev, p->alpha.as<double>() + npix, p->alpha.as<double>(), g->ny,
                                                      g->nx, p->py, p->px, dst);
  OX_KERNEL_CHECK();
  if (out_where == OX_HOST) {
    OX_TRY(stage_out(out, OX_HOST, dst, sizeof(double) * npix * nmaps));
    OX_CUDA(cudaStreamSynchronize(g_stream));
  }
  return OX_OK;
}

}  // extern "C"
