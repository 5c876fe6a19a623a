// lensing.qest.kappa_from_map on the device (call sites tutorials/tt_verification.ipynb:608,610;
// lensing.py:973-976; arithmetic restated in oracle/qe_np.py from the historical estimator):
//
//   kappa_hat(l) = -A_L(l) M_K(l) FFT[ Re IFFT( i lx P_x + i ly P_y ) ],
//   P_{x,y} = FFT[ IFFT(i l_{x,y} k_X W_XY phase) . conj IFFT(k_Y W_Y phase phaseB) ]
//
// TT with filters that vanish on the Nyquist row/column (any lmax below Nyquist) involves only
// Hermitian Fourier arrays and real fields: it runs on half planes with cuFFT Z2D/D2Z -- half the
// bytes of the reference's c2c transforms -- and the reference's ifft -> .real -> fft round trip is
// the identity.  EB (spin-2 phases -> complex real-space fields) and TT with unsymmetric filters
// run the general full-plane c2c chain; the round trip becomes a Hermitian projection
// 1/2 [F(p) + conj F(p')] in the final kernel.  Elementwise work is fused into three kernels:
// legs (1 read, 3 writes), real-space products (3 reads, 2 writes), divergence x normalisation
// (2 reads, 1 write, + the mean-field accumulator).
#include <math.h>
#include <stdlib.h>

#include "ox_common.cuh"
#include "ox_fused_kernels.cuh"

using namespace ox;

struct ox_qeplan {
  ox_geometry *g = nullptr;
  int est = OX_QE_TT, dtype = OX_F64, max_batch = 1;
  bool real_path = false;  // the filters are symmetric and vanish at Nyquist: Hermitian inputs give real fields
  FFTPlans fft;
  DevBuf wxyF, wyF, normF; // T [ny][nx] full-plane tables (general c2c chain; always kept as the fall-back)
  DevBuf wxy, wy, norm;    // T [ny][nxh] half-plane tables (real path)
  DevBuf herm;             // double[2]: scratch of the Hermitian-input check
  DevBuf kx, ky;           // staged inputs (half or full plane complex)
  DevBuf in_real;          // staged real maps
  DevBuf legs, fields, prod, pk, khat, full, full2, out_real;
  // mean-field accumulator packed for ONE all-reduce (Statistics.add_stack, stats.py:1227-1228):
  // float64 [2 * ny * nxh (the complex128 sum of kappa_hat(l) on the half plane) | count | pad]
  DevBuf mf;
  double *mf_count() const { return mf.as<double>() + 2 * (size_t)g->ny * g->nxh; }
  // hand-written FFT path (TT on half planes, power-of-two maps): tables and intermediates in the
  // transposed half-plane layout [plane][ix][iy] of ox_fused_kernels.cuh
  bool fused = false;
  int tw_len = 0;
  DevBuf tw, wxyT, wyT, normT, legT;
  int nlegs = 3;           // 3 (TT) or 6 (EB: real and imaginary parts of the spin-2 fields)
  DevBuf Hx, Hy, Kx, Ky, Lt, Pt, khT;
  // host <-> device pipeline of ox_qe_reconstruct (host buffers, nbatch > 1): two staging slots, copy streams, events
  DevBuf pipe_x[2], pipe_y[2], pipe_out[2];
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_c[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
  ~ox_qeplan() {
    for (int b = 0; b < 2; b++) {
      if (ev_in[b]) cudaEventDestroy(ev_in[b]);
      if (ev_c[b]) cudaEventDestroy(ev_c[b]);
      if (ev_out[b]) cudaEventDestroy(ev_out[b]);
    }
    if (s_h2d) cudaStreamDestroy(s_h2d);
    if (s_d2h) cudaStreamDestroy(s_d2h);
  }
};

namespace {

constexpr int QT = 256;

template <typename T2>
__device__ __forceinline__ T2 mk(double x, double y) {
  T2 r;
  r.x = x;
  r.y = y;
  return r;
}

// e^{2 i theta}, theta = atan2(ly, lx); 1 at l = 0
__device__ __forceinline__ void phase2(double y, double x, double &c, double &s) {
  double l2 = y * y + x * x;
  c = 1.0;
  s = 0.0;
  if (l2 > 0.0) {
    double inv = 1.0 / l2;
    c = (x * x - y * y) * inv;
    s = 2.0 * x * y * inv;
  }
}

// ---- TT on half planes -------------------------------------------------------------------
// legs[m][0] = i lx k W_XY / N, legs[m][1] = i ly k W_XY / N, legs[m][2] = k W_Y / N
template <typename T, typename T2>
__global__ void qe_tt_legs_kernel(const T2 *__restrict__ kx, const T2 *__restrict__ ky, const T *__restrict__ wxy,
                                  const T *__restrict__ wy, const double *__restrict__ ly, const double *__restrict__ lx,
                                  int ny, int nxh, double invn, T2 *__restrict__ legs) {
  const long long nh = (long long)ny * nxh;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= nh) return;
  const long long m = blockIdx.y;
  const int iy = (int)(i / nxh), ix = (int)(i - (long long)iy * nxh);
  T2 a = kx[m * nh + i], b = ky[m * nh + i];
  double wg = (double)wxy[i] * invn, wh = (double)wy[i] * invn;
  double gx = lx[ix] * wg, gy = ly[iy] * wg;
  T2 *o = legs + m * 3 * nh + i;
  o[0] = mk<T2>(-(double)a.y * gx, (double)a.x * gx);  // i * a * gx
  o[nh] = mk<T2>(-(double)a.y * gy, (double)a.x * gy);
  o[2 * nh] = mk<T2>((double)b.x * wh, (double)b.y * wh);
}

// prod[m][0] = gx*h, prod[m][1] = gy*h  (real fields [m][3][n])
template <typename T>
__global__ void qe_tt_prod_kernel(const T *__restrict__ f, long long n, T *__restrict__ prod) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long m = blockIdx.y;
  const T *p = f + m * 3 * n + i;
  T h = p[2 * n];
  prod[m * 2 * n + i] = p[0] * h;
  prod[(m * 2 + 1) * n + i] = p[n] * h;
}

// khat[m] = -norm (i lx Px + i ly Py)   (half plane); optional mean-field accumulation over m
template <typename T, typename T2>
__global__ void qe_tt_div_kernel(const T2 *__restrict__ pk, const T *__restrict__ norm, const double *__restrict__ ly,
                                 const double *__restrict__ lx, int ny, int nx, int nxh, int nb, T2 *__restrict__ khat,
                                 double2 *__restrict__ mf) {
  const long long nh = (long long)ny * nxh;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= nh) return;
  const int iy = (int)(i / nxh), ix = (int)(i - (long long)iy * nxh);
  // on the Nyquist column/row the reference's ifft -> .real -> fft projects the odd l_x / l_y term away
  const double x = (2 * ix == nx) ? 0.0 : lx[ix], y = (2 * iy == ny) ? 0.0 : ly[iy], a = -(double)norm[i];
  double2 acc = mf ? mf[i] : make_double2(0.0, 0.0);
  for (int m = 0; m < nb; m++) {
    T2 px = pk[(long long)m * 2 * nh + i], py = pk[((long long)m * 2 + 1) * nh + i];
    double fr = -(x * (double)px.y + y * (double)py.y), fi = x * (double)px.x + y * (double)py.x;  // i (x px + y py)
    double kr = a * fr, ki = a * fi;
    khat[(long long)m * nh + i] = mk<T2>(kr, ki);
    acc.x += kr;
    acc.y += ki;
  }
  if (mf) mf[i] = acc;
}

// ---- general full-plane chain (EB, or TT with unsymmetric filters) ---------------------------
template <typename T, typename T2, bool SPIN2>
__global__ void qe_gen_legs_kernel(const T2 *__restrict__ kx, const T2 *__restrict__ ky, const T *__restrict__ wxy,
                                   const T *__restrict__ wy, const double *__restrict__ ly, const double *__restrict__ lx,
                                   int ny, int nx, double invn, T2 *__restrict__ legs) {
  const long long n = (long long)ny * nx;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long m = blockIdx.y;
  const int iy = (int)(i / nx), ix = (int)(i - (long long)iy * nx);
  const double x = lx[ix], y = ly[iy];
  T2 a = kx[m * n + i], b = ky[m * n + i];
  double ar = a.x, ai = a.y, br = b.x, bi = b.y;
  if (SPIN2) {
    double c, s;
    phase2(y, x, c, s);
    double tr = ar * c - ai * s, ti = ar * s + ai * c;  // k_E e^{2 i theta}
    ar = tr; ai = ti;
    tr = br * c - bi * s; ti = br * s + bi * c;          // k_B e^{2 i theta}
    br = -ti; bi = tr;                                   // x phaseB = i
  }
  double wg = (double)wxy[i] * invn, wh = (double)wy[i] * invn;
  T2 *o = legs + m * 3 * n + i;
  o[0] = mk<T2>(-ai * x * wg, ar * x * wg);
  o[n] = mk<T2>(-ai * y * wg, ar * y * wg);
  o[2 * n] = mk<T2>(br * wh, bi * wh);
}

// prod[m][0] = gx conj(h), prod[m][1] = gy conj(h)
template <typename T2>
__global__ void qe_gen_prod_kernel(const T2 *__restrict__ f, long long n, T2 *__restrict__ prod) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long m = blockIdx.y;
  const T2 *p = f + m * 3 * n + i;
  T2 gx = p[0], gy = p[n], h = p[2 * n];
  prod[m * 2 * n + i] = mk<T2>((double)gx.x * h.x + (double)gx.y * h.y, (double)gx.y * h.x - (double)gx.x * h.y);
  prod[(m * 2 + 1) * n + i] = mk<T2>((double)gy.x * h.x + (double)gy.y * h.y, (double)gy.y * h.x - (double)gy.x * h.y);
}

// F = i lx Px + i ly Py; the reference's ifft -> .real -> fft is the Hermitian projection
// Fp(p) = 1/2 [F(p) + conj F(p')].  Outputs for half-plane pixels p (p' = mirrored pixel):
//   full[m](p) = -norm(p) Fp(p), full[m](p') = -norm(p') conj Fp(p)      (returnFt, what the reference returns)
//   khat[m](p) = -1/2 (norm(p) + norm(p')) Fp(p)   = Hermitian part of the above (kappa map, mean field)
template <typename T, typename T2>
__global__ void qe_gen_div_kernel(const T2 *__restrict__ pk, const T *__restrict__ norm, const double *__restrict__ ly,
                                  const double *__restrict__ lx, int ny, int nx, int nxh, int nb, T2 *__restrict__ khat,
                                  T2 *__restrict__ full, double2 *__restrict__ mf) {
  const long long nh = (long long)ny * nxh, n = (long long)ny * nx;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= nh) return;
  const int iy = (int)(i / nxh), ix = (int)(i - (long long)iy * nxh);
  const int my = iy ? ny - iy : 0, mx = ix ? nx - ix : 0;
  const long long p = (long long)iy * nx + ix, q = (long long)my * nx + mx;
  const double x = lx[ix], y = ly[iy], xq = lx[mx], yq = ly[my];
  const double np_ = (double)norm[p], nq = (double)norm[q], a = -0.5 * (np_ + nq);
  double2 acc = mf ? mf[i] : make_double2(0.0, 0.0);
  for (int m = 0; m < nb; m++) {
    const T2 *base = pk + (long long)m * 2 * n;
    T2 px = base[p], py = base[n + p], qx = base[q], qy = base[n + q];
    double fr = -(x * (double)px.y + y * (double)py.y), fi = x * (double)px.x + y * (double)py.x;
    double gr = -(xq * (double)qx.y + yq * (double)qy.y), gi = xq * (double)qx.x + yq * (double)qy.x;
    double pr = 0.5 * (fr + gr), pi = 0.5 * (fi - gi);  // Fp(p)
    double kr = a * pr, ki = a * pi;
    khat[(long long)m * nh + i] = mk<T2>(kr, ki);
    if (full) {
      full[(long long)m * n + p] = mk<T2>(-np_ * pr, -np_ * pi);
      full[(long long)m * n + q] = mk<T2>(-nq * pr, nq * pi);
    }
    acc.x += kr;
    acc.y += ki;
  }
  if (mf) mf[i] = acc;
}

// ---- layout helpers ---------------------------------------------------------------------------
template <typename T2>
__global__ void full_to_half_kernel(const T2 *__restrict__ full, int ny, int nx, int nxh, T2 *__restrict__ half) {
  const long long nh = (long long)ny * nxh;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= nh) return;
  const long long m = blockIdx.y;
  const int iy = (int)(i / nxh), ix = (int)(i - (long long)iy * nxh);
  half[m * nh + i] = full[m * (long long)ny * nx + (long long)iy * nx + ix];
}

template <typename T2>
__global__ void half_to_full_kernel(const T2 *__restrict__ half, int ny, int nx, int nxh, double scale, T2 *__restrict__ full) {
  const long long n = (long long)ny * nx, nh = (long long)ny * nxh;
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= n) return;
  const long long m = blockIdx.y;
  const int iy = (int)(p / nx), ix = (int)(p - (long long)iy * nx);
  const bool mirror = ix >= nxh;
  const int sy = mirror ? (iy ? ny - iy : 0) : iy, sx = mirror ? nx - ix : ix;
  T2 z = half[m * nh + (long long)sy * nxh + sx];
  full[m * n + p] = mk<T2>((double)z.x * scale, (mirror ? -(double)z.y : (double)z.y) * scale);
}

template <typename T2>
__global__ void scale_half_kernel(T2 *__restrict__ a, long long n, double s) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    T2 z = a[i];
    a[i] = mk<T2>((double)z.x * s, (double)z.y * s);
  }
}

// take the half-plane columns of a full-plane real table
template <typename T>
__global__ void table_half_kernel(const T *__restrict__ full, int ny, int nx, int nxh, T *__restrict__ half) {
  const long long nh = (long long)ny * nxh;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= nh) return;
  const int iy = (int)(i / nxh), ix = (int)(i - (long long)iy * nxh);
  half[i] = full[(long long)iy * nx + ix];
}

// ---- TT on the hand-written FFT engine -------------------------------------------------------
// Same arithmetic as the three kernels above, attached to the 1-D FFT passes as load/store functors:
//   Q1 rows   : r2c of the input maps                      -> transposed half plane      (fused_row_kernel)
//   Q2a cols  : forward FFT along y                         -> k(l)                       (ColPlainOps)
//   Q2b cols  : k x {i lx W_XY, i ly W_XY, W_Y}/N -> inverse FFT along y, three legs      (LegsOps)
//   Q3a rows  : c2r of the three legs                       -> real fields
//   Q3b rows  : (gradient leg x third field) -> r2c, two products (the "window" of the row kernel)
//   Q4  cols  : forward FFT along y of P_x, then of P_y; the store functors form
//               kappa_hat = -A_L (i lx P_x + i ly P_y) in the transposed layout           (DivOps)
//   finish    : tiled transposition to the natural layout (full plane for returnFt) + mean field
template <typename T>
struct ColPlainOps {
  typedef typename oxfft::V2<T>::type T2;
  const T2 *src;
  T2 *dst;
  int ny, nxh;
  double scale;
  struct Load {
    const T2 *p;
    double sc;
    __device__ __forceinline__ T2 operator()(int e, int) const {
      T2 z = p[e];
      return mk<T2>((double)z.x * sc, (double)z.y * sc);
    }
  };
  typedef oxk::GlobalStore<T2> Store;
  __device__ __forceinline__ Load load(int ix, int plane) const { return Load{src + ((size_t)plane * nxh + ix) * ny, scale}; }
  __device__ __forceinline__ Store store(int ix, int plane) const { return Store{dst + ((size_t)plane * nxh + ix) * ny}; }
};

template <typename T>
struct LegsOps {
  typedef typename oxfft::V2<T>::type T2;
  const T2 *kx, *ky;   // [nb][nxh][ny]: X leg (T or E) and Y leg (T or B)
  const T *legT;       // [nlegs][nxh][ny] precombined tables (one table value per element keeps all the
                       // loads of a thread in flight): TT {W_XY, ly W_XY, W_Y}/N; EB see eb_leg_tables_kernel
  const double *lx;
  T2 *Lt;              // [nb][nlegs][nxh][ny]
  int ny, nxh, nb, nlegs;
  // leg factor (alpha + i beta) x table: kind 0 = i lx (x-gradient), 1 = i (y-gradient, ly is in the table),
  // 2 = 1 (the Y leg); no branch on the leg inside the element loop
  struct Load {
    const T2 *k;
    const T *w;
    double alpha, beta;
    __device__ __forceinline__ T2 operator()(int e, int) const {
      const T2 z = k[e];
      const double wg = (double)__ldg(w + e);
      return mk<T2>(wg * (alpha * (double)z.x - beta * (double)z.y), wg * (alpha * (double)z.y + beta * (double)z.x));
    }
  };
  typedef oxk::GlobalStore<T2> Store;
  // plane = leg * nb + m: the realisations of one leg share its table column, the legs of one
  // realisation share its k column
  __device__ __forceinline__ Load load(int ix, int plane) const {
    const int leg = plane / nb, m = plane - leg * nb;
    const size_t col = ((size_t)m * nxh + ix) * ny;
    // TT: legs (x-grad, y-grad, Y);  EB: (x-grad re, x-grad im, y-grad re, y-grad im, Y re, Y im)
    const int kind = nlegs == 3 ? leg : leg >> 1;
    return Load{(kind == 2 ? ky : kx) + col, legT + ((size_t)leg * nxh + ix) * ny, kind == 2 ? 1.0 : 0.0,
                kind == 0 ? lx[ix] : (kind == 1 ? 1.0 : 0.0)};
  }
  __device__ __forceinline__ Store store(int ix, int plane) const {
    const int leg = plane / nb, m = plane - leg * nb;
    return Store{Lt + (((size_t)m * nlegs + leg) * nxh + ix) * ny};
  }
};

// EB on real fields.  With e^{2 i theta} = c + i s (theta = atan2(ly, lx)) the complex spin-2 fields of the
// estimator split into real ones, each the inverse transform of a Hermitian array:
//   g_x = IFFT(i lx E W_XY (c + i s)) = IFFT(i lx E W_XY c) + i IFFT(i lx E W_XY s),  same for g_y with ly,
//   h   = IFFT(i B W_Y (c + i s))     = IFFT(-B W_Y s)      + i IFFT(B W_Y c),
// and only Re(g conj h) = g_r h_r + g_i h_i survives the reference's ifft -> .real -> fft round trip
// (the imaginary part of the product transforms to an anti-Hermitian term).  Tables, transposed layout:
// {W_XY c, W_XY s, ly W_XY c, ly W_XY s, -W_Y s, W_Y c} / N
template <typename T>
__global__ void eb_leg_tables_kernel(const T *__restrict__ wxyT, const T *__restrict__ wyT, const double *__restrict__ ly,
                                     const double *__restrict__ lx, int ny, int nxh, double invn, T *__restrict__ legT) {
  const long long nh = (long long)ny * nxh;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nh; i += (long long)gridDim.x * blockDim.x) {
    const int ix = (int)(i / ny), iy = (int)(i - (long long)ix * ny);
    double c, s;
    phase2(ly[iy], lx[ix], c, s);
    const double wg = (double)wxyT[i] * invn, wh = (double)wyT[i] * invn;
    legT[i] = (T)(wg * c);
    legT[nh + i] = (T)(wg * s);
    legT[2 * nh + i] = (T)(ly[iy] * wg * c);
    legT[3 * nh + i] = (T)(ly[iy] * wg * s);
    legT[4 * nh + i] = (T)(-wh * s);
    legT[5 * nh + i] = (T)(wh * c);
  }
}

// EXHAUSTIVE Hermitian test: sum over every pair {p, p'} (rows 0..ny/2 against their mirror rows, so each element
// of the plane is read once: 2 s N bytes against the >= 28 s N of the chain) of |k(p) - conj k(p')|^2 and
// |k(p)|^2 + |k(p')|^2: acc[0], acc[1].  The caller takes the half-plane path only when the anti-Hermitian part
// carries less than 1e-20 (fp64) of the power, i.e. when dropping it changes the result by < 1e-10 in norm.
template <typename T2>
__global__ void herm_check_kernel(const T2 *__restrict__ k, int ny, int nx, double *__restrict__ acc) {
  const long long plane = blockIdx.y;
  const T2 *kp = k + plane * (long long)ny * nx;
  double d2 = 0.0, n2 = 0.0;
  const int nrows = ny / 2 + 1;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < (long long)nrows * nx; t += (long long)gridDim.x * blockDim.x) {
    const int iy = (int)(t / nx), ix = (int)(t % nx);
    const int my = iy ? ny - iy : 0, mx = ix ? nx - ix : 0;
    const T2 a = kp[(long long)iy * nx + ix], b = kp[(long long)my * nx + mx];
    const double dx = (double)a.x - (double)b.x, dy = (double)a.y + (double)b.y;
    d2 += dx * dx + dy * dy;
    n2 += (double)a.x * a.x + (double)a.y * a.y + (double)b.x * b.x + (double)b.y * b.y;
  }
  for (int o = 16; o > 0; o >>= 1) {
    d2 += __shfl_xor_sync(0xffffffffu, d2, o);
    n2 += __shfl_xor_sync(0xffffffffu, n2, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(acc, d2);
    atomicAdd(acc + 1, n2);
  }
}

// legT[0] = W_XY/N, legT[1] = ly W_XY/N, legT[2] = W_Y/N in the transposed layout
template <typename T>
__global__ void leg_tables_kernel(const T *__restrict__ wxyT, const T *__restrict__ wyT, const double *__restrict__ ly, int ny,
                                  int nxh, double invn, T *__restrict__ legT) {
  const long long nh = (long long)ny * nxh;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nh; i += (long long)gridDim.x * blockDim.x) {
    const int iy = (int)(i % ny);
    const double wg = (double)wxyT[i] * invn;
    legT[i] = (T)wg;
    legT[nh + i] = (T)(ly[iy] * wg);
    legT[2 * nh + i] = (T)((double)wyT[i] * invn);
  }
}

// kappa_hat in the transposed layout (input of the inverse transform to the kappa map):
// khT[m][ix][iy] = -A (i lx P_x + i ly P_y), same arithmetic as qe_tt_div_kernel
template <typename T, typename T2>
__global__ void qe_divT_kernel(const T2 *__restrict__ Pk, const T *__restrict__ normT, const double *__restrict__ ly,
                               const double *__restrict__ lx, int ny, int nx, int nxh, double scale, T2 *__restrict__ khT) {
  const long long nh = (long long)ny * nxh;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= nh) return;
  const long long m = blockIdx.y;
  const int ix = (int)(i / ny), iy = (int)(i - (long long)ix * ny);
  const double x = (2 * ix == nx) ? 0.0 : lx[ix], y = (2 * iy == ny) ? 0.0 : ly[iy], a = -(double)normT[i];
  const T2 px = Pk[m * 2 * nh + i], py = Pk[(m * 2 + 1) * nh + i];
  const double fr = -(x * (double)px.y + y * (double)py.y), fi = x * (double)px.x + y * (double)py.x;
  khT[m * nh + i] = mk<T2>(a * fr * scale, a * fi * scale);
}

// out[ix][iy] = in[iy][ix] for the nxh half-plane columns of a [ny][ldin] array (tables, alreadyFTed k-maps)
template <typename V>
__global__ void to_halfT_kernel(const V *__restrict__ in, int ny, int ldin, int nxh, long long in_plane, V *__restrict__ out) {
  __shared__ V tile[32][33];
  const long long m = blockIdx.z;
  const int ix0 = blockIdx.x * 32, iy0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    int iy = iy0 + j, ix = ix0 + threadIdx.x;
    if (iy < ny && ix < nxh) tile[j][threadIdx.x] = in[m * in_plane + (long long)iy * ldin + ix];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    int ix = ix0 + j, iy = iy0 + threadIdx.x;
    if (iy < ny && ix < nxh) out[(m * nxh + ix) * ny + iy] = tile[threadIdx.x][j];
  }
}

// kappa_hat(l) = -A (i lx P_x + i ly P_y) (arithmetic of qe_tt_div_kernel) from the transposed products
// Pk[m][2][ix][iy], written in the natural layout: full[m] (full plane, Hermitian extension; returnFt)
// and the mean-field accumulator mf[iy][ix] += sum_m kappa_hat (fixed m order)
template <typename T, typename T2>
__global__ void qe_finish_kernel(const T2 *__restrict__ Pk, const T *__restrict__ norm, const double *__restrict__ ly,
                                 const double *__restrict__ lx, int ny, int nx, int nxh, int nb, T2 *__restrict__ full,
                                 double2 *__restrict__ mf) {
  __shared__ T2 tile[2][32][33];
  const int ix0 = blockIdx.x * 32, iy0 = blockIdx.y * 32;
  const long long n = (long long)ny * nx, nh = (long long)ny * nxh;
  const int ixo = ix0 + threadIdx.x;   // this thread's output column
  const double x = (ixo < nxh && 2 * ixo != nx) ? lx[ixo] : 0.0;
  double2 acc[4];
  double a[4], y[4];
#pragma unroll
  for (int jj = 0; jj < 4; jj++) {
    const int iy = iy0 + threadIdx.y + 8 * jj;
    acc[jj] = make_double2(0.0, 0.0);
    a[jj] = ixo < nxh ? -(double)norm[(long long)iy * nxh + ixo] : 0.0;
    y[jj] = (2 * iy == ny) ? 0.0 : ly[iy];
  }
  for (int m = 0; m < nb; m++) {
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
      const int ix = ix0 + j, iy = iy0 + threadIdx.x;
      if (ix < nxh) {
        tile[0][j][threadIdx.x] = Pk[((long long)m * 2 * nxh + ix) * ny + iy];
        tile[1][j][threadIdx.x] = Pk[(((long long)m * 2 + 1) * nxh + ix) * ny + iy];
      }
    }
    __syncthreads();
    if (ixo >= nxh) continue;
#pragma unroll
    for (int jj = 0; jj < 4; jj++) {
      const int j = threadIdx.y + 8 * jj, iy = iy0 + j;
      const T2 px = tile[0][threadIdx.x][j], py = tile[1][threadIdx.x][j];
      const double fr = -(x * (double)px.y + y[jj] * (double)py.y), fi = x * (double)px.x + y[jj] * (double)py.x;
      const double kr = a[jj] * fr, ki = a[jj] * fi;
      acc[jj].x += kr;
      acc[jj].y += ki;
      if (full) {
        full[m * n + (long long)iy * nx + ixo] = mk<T2>(kr, ki);
        if (ixo > 0 && 2 * ixo < nx) {
          const int my = iy ? ny - iy : 0;
          full[m * n + (long long)my * nx + (nx - ixo)] = mk<T2>(kr, -ki);
        }
      }
    }
  }
  if (mf && ixo < nxh) {
#pragma unroll
    for (int jj = 0; jj < 4; jj++) {
      const long long o = (long long)(iy0 + threadIdx.y + 8 * jj) * nxh + ixo;
      double2 v = mf[o];
      v.x += acc[jj].x;
      v.y += acc[jj].y;
      mf[o] = v;
    }
  }
}

// row passes of the estimator: r2c of maps, c2r to maps, product with a second map -> r2c
typedef oxk::RowModes<oxk::ROW_OUT_H, oxk::ROW_IN_H | oxk::ROW_OUT_MAP, oxk::ROW_WIN | oxk::ROW_OUT_H,
                      oxk::ROW_WIN | oxk::ROW_WIN2 | oxk::ROW_OUT_H> QeRowModes;

bool qe_fused_supported(int ny, int nx) {
  auto p2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
  return p2(ny) && p2(nx) && ny >= 512 && ny <= 8192 && nx >= 256 && nx <= 8192;
}

template <typename T, typename T2>
int reconstruct_fused_T(ox_qeplan *q, const void *x, const void *y, int where, int nb, int already_ft, int return_ft,
                        int accumulate, void *out, int out_where) {
  ox_geometry *g = q->g;
  const int ny = g->ny, nx = g->nx, nxh = g->nxh;
  const long long n = (long long)ny * nx, nh = (long long)ny * nxh;
  const double invn = 1.0 / ((double)ny * (double)nx);
  const double *ly = g->ly.as<double>(), *lx = g->lx.as<double>();
  const bool two = (y != nullptr && y != x);
  const size_t hb = sizeof(T2) * (size_t)q->max_batch * nh;
  OX_TRY(q->Kx.ensure(hb));
  if (two) OX_TRY(q->Ky.ensure(hb));
  const dim3 tb(32, 8);
  int st;
  // ---- inputs -> k(l) in the transposed half-plane layout
  for (int leg = 0; leg < (two ? 2 : 1); leg++) {
    const void *src = leg ? y : x;
    T2 *K = leg ? q->Ky.as<T2>() : q->Kx.as<T2>();
    if (already_ft) {
      const void *dsrc;
      OX_TRY(stage_in(src, where, sizeof(T2) * (size_t)nb * n, q->full, &dsrc));
      dim3 grid((nxh + 31) / 32, (ny + 31) / 32, nb);
      to_halfT_kernel<T2><<<grid, tb, 0, g_stream>>>((const T2 *)dsrc, ny, nx, nxh, n, K);
      OX_KERNEL_CHECK();
    } else {
      const void *dsrc = src;
      if (where == OX_HOST) {
        OX_TRY(q->in_real.ensure(sizeof(T) * (size_t)q->max_batch * n));
        OX_CUDA(cudaMemcpyAsync(q->in_real.p, src, sizeof(T) * (size_t)nb * n, cudaMemcpyHostToDevice, g_stream));
        dsrc = q->in_real.p;
      }
      OX_TRY(q->Hx.ensure(hb));
      oxk::RowArgs<T> ra;
      ra.Hin = nullptr;
      ra.map_in = (const T *)dsrc;
      ra.Hout = q->Hx.as<T2>();
      ra.map_out = nullptr;
      ra.window = nullptr;
      ra.tw = q->tw.as<T2>();
      ra.tw_len = q->tw_len;
      ra.ny = ny; ra.nx = nx; ra.mx = nx / 2;
      ra.map_in_group_stride = n;
      OX_TRY(oxk::launch_row_any<T>(ra, nb, QeRowModes()));                                            // Q1
      stage_mark("Q1 rows r2c");
      ColPlainOps<T> op{q->Hx.as<T2>(), K, ny, nxh, 1.0};
      OX_COL_DISPATCH(T, -1, op, q->tw.p, q->tw_len, nxh, nb, ny, st);                   // Q2a
      OX_TRY(st);
      stage_mark("Q2a cols fwd");
    }
  }
  // ---- legs -> real fields -> products
  const int nl = q->nlegs;   // 3 (TT) or 6 (EB)
  OX_TRY(q->Lt.ensure((size_t)nl * hb));
  OX_TRY(q->fields.ensure(sizeof(T) * (size_t)q->max_batch * nl * n));
  OX_TRY(q->Pt.ensure(2 * hb));
  {
    LegsOps<T> op{q->Kx.as<T2>(), two ? q->Ky.as<T2>() : q->Kx.as<T2>(), q->legT.as<T>(), lx, q->Lt.as<T2>(), ny, nxh, nb, nl};
    OX_COL_DISPATCH(T, +1, op, q->tw.p, q->tw_len, nxh, (long long)nl * nb, ny, st);     // Q2b
    OX_TRY(st);
    stage_mark("Q2b legs cols inv");
  }
  {
    oxk::RowArgs<T> ra;
    ra.Hin = q->Lt.as<T2>();
    ra.map_in = nullptr;
    ra.Hout = nullptr;
    ra.map_out = q->fields.as<T>();
    ra.window = nullptr;
    ra.tw = q->tw.as<T2>();
    ra.tw_len = q->tw_len;
    ra.ny = ny; ra.nx = nx; ra.mx = nx / 2;
    OX_TRY(oxk::launch_row_any<T>(ra, (long long)nl * nb, QeRowModes()));                              // Q3a
    stage_mark("Q3a rows c2r");
    ra.Hin = nullptr;
    ra.map_out = nullptr;
    ra.Hout = q->Pt.as<T2>();
    ra.group = 2;
    ra.map_in_group_stride = nl * n;
    ra.win_group_stride = nl * n;
    if (nl == 3) {
      ra.map_in = q->fields.as<T>();           // g_x, g_y
      ra.window = q->fields.as<T>() + 2 * n;   // x the third field of each realisation
    } else {
      // Re(g conj h) = g_r h_r + g_i h_i: fields (gx_r, gx_i, gy_r, gy_i, h_r, h_i)
      ra.map_in = q->fields.as<T>();
      ra.map_in2 = q->fields.as<T>() + n;
      ra.map_in_sub_stride = 2 * n;
      ra.window = q->fields.as<T>() + 4 * n;
      ra.window2 = q->fields.as<T>() + 5 * n;
    }
    OX_TRY(oxk::launch_row_any<T>(ra, 2LL * nb, QeRowModes()));                                        // Q3b
    stage_mark("Q3b product rows r2c");
  }
  {
    ColPlainOps<T> op{q->Pt.as<T2>(), q->Pt.as<T2>(), ny, nxh, 1.0};   // in place: a CTA owns its column
    OX_COL_DISPATCH(T, -1, op, q->tw.p, q->tw_len, nxh, 2LL * nb, ny, st);               // Q4
    OX_TRY(st);
    stage_mark("Q4 cols fwd");
  }
  // ---- outputs
  double2 *mf = accumulate ? q->mf.as<double2>() : nullptr;
  const dim3 fgrid((nxh + 31) / 32, ny / 32);
  if (return_ft || mf) {
    T2 *fullp = nullptr;
    if (return_ft) {
      fullp = (T2 *)out;
      if (out_where == OX_HOST) {
        OX_TRY(q->full.ensure(sizeof(T2) * (size_t)nb * n));
        fullp = q->full.as<T2>();
      }
    }
    qe_finish_kernel<T, T2><<<fgrid, tb, 0, g_stream>>>(q->Pt.as<T2>(), q->norm.as<T>(), ly, lx, ny, nx, nxh, nb, fullp, mf);
    OX_KERNEL_CHECK();
    stage_mark("finish div+meanfield");
    if (return_ft) {
      if (out_where == OX_HOST) OX_TRY(stage_out(out, OX_HOST, fullp, sizeof(T2) * (size_t)nb * n));
      return OX_OK;
    }
  }
  // real-space kappa = IFFT(kappa_hat)/N: kappa_hat/N in the transposed layout, inverse FFT along y, c2r rows
  OX_TRY(q->khT.ensure(hb));
  OX_TRY(q->out_real.ensure(sizeof(T) * (size_t)q->max_batch * n));
  {
    dim3 gh((unsigned)((nh + QT - 1) / QT), nb);
    qe_divT_kernel<T, T2><<<gh, QT, 0, g_stream>>>(q->Pt.as<T2>(), q->normT.as<T>(), ly, lx, ny, nx, nxh, invn, q->khT.as<T2>());
    OX_KERNEL_CHECK();
    ColPlainOps<T> op{q->khT.as<T2>(), q->khT.as<T2>(), ny, nxh, 1.0};
    OX_COL_DISPATCH(T, +1, op, q->tw.p, q->tw_len, nxh, nb, ny, st);
    OX_TRY(st);
    oxk::RowArgs<T> ra;
    ra.Hin = q->khT.as<T2>();
    ra.map_in = nullptr;
    ra.Hout = nullptr;
    ra.map_out = q->out_real.as<T>();
    ra.window = nullptr;
    ra.tw = q->tw.as<T2>();
    ra.tw_len = q->tw_len;
    ra.ny = ny; ra.nx = nx; ra.mx = nx / 2;
    OX_TRY(oxk::launch_row_any<T>(ra, nb, QeRowModes()));
    stage_mark("kappa map (div, cols inv, rows c2r)");
  }
  return stage_out(out, out_where, q->out_real.p, sizeof(T) * (size_t)nb * n);
}

template <typename T, typename T2>
int reconstruct_T(ox_qeplan *q, const void *x, const void *y, int where, int nb, int already_ft, int return_ft,
                  int accumulate, void *out, int out_where) {
  ox_geometry *g = q->g;
  const int ny = g->ny, nx = g->nx, nxh = g->nxh;
  const long long n = (long long)ny * nx, nh = (long long)ny * nxh;
  const double invn = 1.0 / ((double)ny * (double)nx);
  const double *ly = g->ly.as<double>(), *lx = g->lx.as<double>();
  const bool two = (y != nullptr && y != x);
  // the half-plane paths need Hermitian Fourier inputs (transforms of real maps): verified on EVERY pixel pair
  // when the caller hands in k-maps; anything else goes through the reference's c2c chain
  bool real = q->real_path && (q->est == OX_QE_TT || q->fused);
  if (real && already_ft) {
    OX_TRY(q->herm.ensure(2 * sizeof(double)));
    OX_CUDA(cudaMemsetAsync(q->herm.p, 0, 2 * sizeof(double), g_stream));
    const void *staged[2] = {nullptr, nullptr};
    for (int leg = 0; leg < (two ? 2 : 1); leg++) {
      OX_TRY(stage_in(leg ? y : x, where, sizeof(T2) * (size_t)nb * n, leg ? q->full2 : q->full, &staged[leg]));
      herm_check_kernel<T2><<<dim3(sm_count() * 4, nb), QT, 0, g_stream>>>((const T2 *)staged[leg], ny, nx, q->herm.as<double>());
      OX_KERNEL_CHECK();
    }
    // host inputs are on the device now: hand the staged copies on instead of uploading them again
    x = staged[0];
    y = two ? staged[1] : (y ? staged[0] : nullptr);
    where = OX_DEVICE;
    double h[2];
    OX_CUDA(cudaMemcpyAsync(h, q->herm.p, sizeof(h), cudaMemcpyDeviceToHost, g_stream));
    OX_CUDA(cudaStreamSynchronize(g_stream));
    const double tol = q->dtype == OX_F64 ? 1e-20 : 1e-9;   // squared relative asymmetry
    if (!(h[0] <= tol * h[1])) real = false;
  }
  if (real && q->fused) return reconstruct_fused_T<T, T2>(q, x, y, where, nb, already_ft, return_ft, accumulate, out, out_where);
  const long long plane = real ? nh : n;  // complex elements per staged k-map
  OX_TRY(q->kx.ensure(sizeof(T2) * (size_t)q->max_batch * plane));
  if (two) OX_TRY(q->ky.ensure(sizeof(T2) * (size_t)q->max_batch * plane));
  // ---- inputs -> k-maps in the layout of the chosen path
  for (int leg = 0; leg < (two ? 2 : 1); leg++) {
    const void *src = leg ? y : x;
    T2 *dstk = leg ? q->ky.as<T2>() : q->kx.as<T2>();
    if (already_ft) {
      const void *dsrc;
      OX_TRY(stage_in(src, where, sizeof(T2) * (size_t)nb * n, q->full, &dsrc));
      if (real) {
        dim3 grid((unsigned)((nh + QT - 1) / QT), nb);
        full_to_half_kernel<T2><<<grid, QT, 0, g_stream>>>((const T2 *)dsrc, ny, nx, nxh, dstk);
        OX_KERNEL_CHECK();
      } else {
        OX_CUDA(cudaMemcpyAsync(dstk, dsrc, sizeof(T2) * (size_t)nb * n, cudaMemcpyDeviceToDevice, g_stream));
      }
    } else {
      OX_TRY(q->in_real.ensure(sizeof(T) * (size_t)q->max_batch * n));
      OX_CUDA(cudaMemcpyAsync(q->in_real.p, src, sizeof(T) * (size_t)nb * n,
                              where == OX_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, g_stream));
      if (real) {
        OX_TRY(q->fft.exec_r2c(nb, q->in_real.p, dstk));
      } else {
        OX_TRY(q->khat.ensure(sizeof(T2) * (size_t)q->max_batch * nh));
        OX_TRY(q->fft.exec_r2c(nb, q->in_real.p, q->khat.p));
        dim3 grid((unsigned)((n + QT - 1) / QT), nb);
        half_to_full_kernel<T2><<<grid, QT, 0, g_stream>>>(q->khat.as<T2>(), ny, nx, nxh, 1.0, dstk);
        OX_KERNEL_CHECK();
      }
    }
  }
  const T2 *kx = q->kx.as<T2>(), *ky = two ? q->ky.as<T2>() : kx;
  OX_TRY(q->khat.ensure(sizeof(T2) * (size_t)q->max_batch * nh));
  double2 *mf = accumulate ? q->mf.as<double2>() : nullptr;
  if (real) {
    OX_TRY(q->legs.ensure(sizeof(T2) * (size_t)q->max_batch * 3 * nh));
    OX_TRY(q->fields.ensure(sizeof(T) * (size_t)q->max_batch * 3 * n));
    OX_TRY(q->prod.ensure(sizeof(T) * (size_t)q->max_batch * 2 * n));
    OX_TRY(q->pk.ensure(sizeof(T2) * (size_t)q->max_batch * 2 * nh));
    dim3 gh((unsigned)((nh + QT - 1) / QT), nb), gf((unsigned)((n + QT - 1) / QT), nb);
    qe_tt_legs_kernel<T, T2><<<gh, QT, 0, g_stream>>>(kx, ky, q->wxy.as<T>(), q->wy.as<T>(), ly, lx, ny, nxh, invn, q->legs.as<T2>());
    OX_KERNEL_CHECK();
    OX_TRY(q->fft.exec_c2r(3 * nb, q->legs.p, q->fields.p));
    qe_tt_prod_kernel<T><<<gf, QT, 0, g_stream>>>(q->fields.as<T>(), n, q->prod.as<T>());
    OX_KERNEL_CHECK();
    OX_TRY(q->fft.exec_r2c(2 * nb, q->prod.p, q->pk.p));
    qe_tt_div_kernel<T, T2><<<(unsigned)((nh + QT - 1) / QT), QT, 0, g_stream>>>(q->pk.as<T2>(), q->norm.as<T>(), ly, lx, ny, nx, nxh, nb,
                                                                              q->khat.as<T2>(), mf);
    OX_KERNEL_CHECK();
  } else {
    OX_TRY(q->legs.ensure(sizeof(T2) * (size_t)q->max_batch * 3 * n));
    OX_TRY(q->pk.ensure(sizeof(T2) * (size_t)q->max_batch * 2 * n));
    dim3 gf((unsigned)((n + QT - 1) / QT), nb);
    if (q->est == OX_QE_EB)
      qe_gen_legs_kernel<T, T2, true><<<gf, QT, 0, g_stream>>>(kx, ky, q->wxyF.as<T>(), q->wyF.as<T>(), ly, lx, ny, nx, invn, q->legs.as<T2>());
    else
      qe_gen_legs_kernel<T, T2, false><<<gf, QT, 0, g_stream>>>(kx, ky, q->wxyF.as<T>(), q->wyF.as<T>(), ly, lx, ny, nx, invn, q->legs.as<T2>());
    OX_KERNEL_CHECK();
    OX_TRY(q->fft.exec_c2c(3 * nb, q->legs.p, q->legs.p, CUFFT_INVERSE));
    qe_gen_prod_kernel<T2><<<gf, QT, 0, g_stream>>>(q->legs.as<T2>(), n, q->pk.as<T2>());
    OX_KERNEL_CHECK();
    OX_TRY(q->fft.exec_c2c(2 * nb, q->pk.p, q->pk.p, CUFFT_FORWARD));
    T2 *fullp = nullptr;
    if (return_ft) {
      fullp = (T2 *)out;
      if (out_where == OX_HOST) {
        OX_TRY(q->full.ensure(sizeof(T2) * (size_t)nb * n));
        fullp = q->full.as<T2>();
      }
    }
    qe_gen_div_kernel<T, T2><<<(unsigned)((nh + QT - 1) / QT), QT, 0, g_stream>>>(q->pk.as<T2>(), q->normF.as<T>(), ly, lx, ny, nx, nxh,
                                                                               nb, q->khat.as<T2>(), fullp, mf);
    OX_KERNEL_CHECK();
    if (return_ft) {
      if (out_where == OX_HOST) OX_TRY(stage_out(out, OX_HOST, fullp, sizeof(T2) * (size_t)nb * n));
      return OX_OK;
    }
  }
  // ---- outputs
  if (return_ft) {
    size_t bytes = sizeof(T2) * (size_t)nb * n;
    void *dst = out;
    if (out_where == OX_HOST) {
      OX_TRY(q->full.ensure(bytes));
      dst = q->full.p;
    }
    dim3 gf((unsigned)((n + QT - 1) / QT), nb);
    half_to_full_kernel<T2><<<gf, QT, 0, g_stream>>>(q->khat.as<T2>(), ny, nx, nxh, 1.0, (T2 *)dst);
    OX_KERNEL_CHECK();
    if (out_where == OX_HOST) OX_TRY(stage_out(out, OX_HOST, dst, bytes));
    return OX_OK;
  }
  // real-space kappa = IFFT(kappa_hat)/N: scale a copy so the mean field keeps the unscaled values
  size_t bytes = sizeof(T) * (size_t)nb * n;
  OX_TRY(q->out_real.ensure(sizeof(T) * (size_t)q->max_batch * n));
  scale_half_kernel<T2><<<sm_count() * 8, QT, 0, g_stream>>>(q->khat.as<T2>(), (long long)nb * nh, invn);
  OX_KERNEL_CHECK();
  OX_TRY(q->fft.exec_c2r(nb, q->khat.p, q->out_real.p));
  return stage_out(out, out_where, q->out_real.p, bytes);
}

__global__ void add_count_kernel(double *c, double v) { *c += v; }

template <typename T>
int upload_tables(ox_qeplan *q, const double *wxy, const double *wy, const double *norm, int where) {
  ox_geometry *g = q->g;
  const long long n = (long long)g->ny * g->nx, nh = (long long)g->ny * g->nxh;
  const double *src[3] = {wxy, wy, norm};
  DevBuf *dst[3] = {&q->wxy, &q->wy, &q->norm};
  DevBuf stage, fullT;
  for (int t = 0; t < 3; t++) {
    const void *d;
    OX_TRY(stage_in(src[t], where, sizeof(double) * n, stage, &d));
    OX_TRY(fullT.ensure(sizeof(T) * n));
    OX_TRY(cast_from_f64((const double *)d, fullT.p, n, q->dtype));
    if (q->fused) {
      DevBuf *dstT[3] = {&q->wxyT, &q->wyT, &q->normT};
      OX_TRY(dstT[t]->ensure(sizeof(T) * nh));
      dim3 grid((g->nxh + 31) / 32, (g->ny + 31) / 32, 1);
      to_halfT_kernel<T><<<grid, dim3(32, 8), 0, g_stream>>>(fullT.as<T>(), g->ny, g->nx, g->nxh, n, dstT[t]->as<T>());
      OX_KERNEL_CHECK();
    }
    if (q->real_path) {
      OX_TRY(dst[t]->ensure(sizeof(T) * nh));
      table_half_kernel<T><<<(unsigned)((nh + QT - 1) / QT), QT, 0, g_stream>>>(fullT.as<T>(), g->ny, g->nx, g->nxh, dst[t]->as<T>());
      OX_KERNEL_CHECK();
    }
    DevBuf *dstF[3] = {&q->wxyF, &q->wyF, &q->normF};
    OX_TRY(dstF[t]->ensure(sizeof(T) * n));
    OX_CUDA(cudaMemcpyAsync(dstF[t]->p, fullT.p, sizeof(T) * n, cudaMemcpyDeviceToDevice, g_stream));
    OX_CUDA(cudaStreamSynchronize(g_stream));
  }
  return OX_OK;
}

template <typename T>
int make_leg_tables(ox_qeplan *q) {
  ox_geometry *g = q->g;
  const long long nh = (long long)g->ny * g->nxh;
  const double invn = 1.0 / ((double)g->ny * (double)g->nx);
  q->nlegs = q->est == OX_QE_EB ? 6 : 3;
  OX_TRY(q->legT.ensure(sizeof(T) * (size_t)q->nlegs * nh));
  if (q->est == OX_QE_EB)
    eb_leg_tables_kernel<T><<<sm_count() * 8, QT, 0, g_stream>>>(q->wxyT.as<T>(), q->wyT.as<T>(), g->ly.as<double>(),
                                                                 g->lx.as<double>(), g->ny, g->nxh, invn, q->legT.as<T>());
  else
    leg_tables_kernel<T><<<sm_count() * 8, QT, 0, g_stream>>>(q->wxyT.as<T>(), q->wyT.as<T>(), g->ly.as<double>(), g->ny, g->nxh,
                                                              invn, q->legT.as<T>());
  OX_KERNEL_CHECK();
  OX_CUDA(cudaStreamSynchronize(g_stream));
  q->wxyT.release();
  q->wyT.release();
  return OX_OK;
}

}  // namespace

extern "C" {

int ox_qeplan_create(ox_geometry *g, int est, const double *wxy, const double *wy, const double *norm, int where, int dtype,
                     int max_batch, int real_path, ox_qeplan **out) {
  OX_REQUIRE(g && wxy && wy && norm && out, "ox_qeplan_create: null pointer");
  OX_REQUIRE(est == OX_QE_TT || est == OX_QE_EB, "unknown estimator %d", est);
  OX_REQUIRE(dtype == OX_F64 || dtype == OX_F32, "bad dtype %d", dtype);
  OX_REQUIRE(max_batch >= 1, "max_batch must be >= 1");
  ox_qeplan *q = new ox_qeplan;
  q->g = g;
  q->est = est;
  q->dtype = dtype;
  q->max_batch = max_batch;
  q->real_path = real_path != 0;
  q->fft.ny = g->ny;
  q->fft.nx = g->nx;
  q->fft.dtype = dtype;
  // hand-written FFT path for TT on half planes (ORPHX_QE=cufft keeps the cuFFT chain)
  const char *env = getenv("ORPHX_QE");
  q->fused = q->real_path && qe_fused_supported(g->ny, g->nx) && !(env && !strcmp(env, "cufft"));
  if (q->fused) {
    q->tw_len = g->ny > g->nx ? g->ny : g->nx;
    int ts = fused_make_twiddles(q->tw_len, dtype, q->tw);
    if (ts != OX_OK) {
      delete q;
      return ts;
    }
  }
  int st = dtype == OX_F64 ? upload_tables<double>(q, wxy, wy, norm, where) : upload_tables<float>(q, wxy, wy, norm, where);
  if (st == OX_OK && q->fused) st = dtype == OX_F64 ? make_leg_tables<double>(q) : make_leg_tables<float>(q);
  if (st == OX_OK) st = q->mf.ensure(sizeof(double2) * ((size_t)g->ny * g->nxh + 1));
  if (st != OX_OK) {
    delete q;
    return st;
  }
  *out = q;
  return ox_qe_meanfield_reset(q);
}

int ox_qeplan_destroy(ox_qeplan *q) {
  delete q;
  return OX_OK;
}

int ox_qe_meanfield_reset(ox_qeplan *q) {
  OX_REQUIRE(q, "null plan");
  OX_CUDA(cudaMemsetAsync(q->mf.p, 0, sizeof(double2) * ((size_t)q->g->ny * q->g->nxh + 1), g_stream));
  return OX_OK;
}

int ox_qe_path(ox_qeplan *q) {
  if (!q) return -1;
  if (q->fused) return q->est == OX_QE_EB ? 3 : 2;
  return (q->real_path && q->est == OX_QE_TT) ? 1 : 0;
}

int ox_qe_meanfield(ox_qeplan *q, double **packed_dev, long long *nelem) {
  OX_REQUIRE(q, "null plan");
  if (packed_dev) *packed_dev = q->mf.as<double>();
  if (nelem) *nelem = (long long)q->g->ny * q->g->nxh;
  return OX_OK;
}

int ox_qe_meanfield_allreduce(ox_comm *c, ox_qeplan *q) {
  OX_REQUIRE(c && q, "null pointer");
  return comm_allreduce_f64(c, q->mf.as<double>(), 2 * (long long)q->g->ny * q->g->nxh + 1);   // [stack | count]
}

int ox_qe_reconstruct(ox_qeplan *q, const void *x, const void *y, int where, int nbatch, int already_ft, int return_ft,
                      int accumulate_meanfield, void *kappa_out, int out_where) {
  OX_REQUIRE(q && x && kappa_out, "ox_qe_reconstruct: null pointer");
  OX_REQUIRE(nbatch >= 1 && nbatch <= q->max_batch, "nbatch=%d outside 1..max_batch=%d", nbatch, q->max_batch);
  OX_REQUIRE(q->est != OX_QE_EB || (y && y != x), "EB needs the B leg");
  {
    // Host buffers, several realisations: one realisation at a time through two device staging slots, the upload of
    // realisation r+1 and the download of realisation r-1 on their own (non-blocking) streams while realisation r is
    // reconstructed on the library stream -- the call is PCIe-bound (2 x 134 MB per realisation at 4096^2 against
    // 0.8 ms of kernels) and the two directions of the link run at the same time.  ORPHX_QE_PIPELINE=0: one batch, serial.
    const char *env = getenv("ORPHX_QE_PIPELINE");
    if (where == OX_HOST && out_where == OX_HOST && nbatch > 1 && !(env && env[0] == '0')) {
      const size_t es = q->dtype == OX_F64 ? 8 : 4;
      const size_t npix = (size_t)q->g->ny * q->g->nx;
      const size_t in_bytes = (already_ft ? 2 : 1) * es * npix, out_bytes = (return_ft ? 2 : 1) * es * npix;
      if (!q->s_h2d) {
        OX_CUDA(cudaStreamCreateWithFlags(&q->s_h2d, cudaStreamNonBlocking));
        OX_CUDA(cudaStreamCreateWithFlags(&q->s_d2h, cudaStreamNonBlocking));
        for (int b = 0; b < 2; b++) {
          OX_CUDA(cudaEventCreateWithFlags(&q->ev_in[b], cudaEventDisableTiming));
          OX_CUDA(cudaEventCreateWithFlags(&q->ev_c[b], cudaEventDisableTiming));
          OX_CUDA(cudaEventCreateWithFlags(&q->ev_out[b], cudaEventDisableTiming));
        }
      }
      for (int b = 0; b < 2; b++) {
        OX_TRY(q->pipe_x[b].ensure(in_bytes));
        if (y) OX_TRY(q->pipe_y[b].ensure(in_bytes));
        OX_TRY(q->pipe_out[b].ensure(out_bytes));
      }
      OX_CUDA(cudaStreamSynchronize(g_stream));   // earlier work of the library stream may still use the buffers' memory
      int st = OX_OK;
      for (int r = 0; r < nbatch && st == OX_OK; r++) {
        const int b = r & 1;
        if (r >= 2) OX_CUDA(cudaStreamWaitEvent(q->s_h2d, q->ev_c[b], 0));      // slot b's previous realisation has been consumed
        OX_CUDA(cudaMemcpyAsync(q->pipe_x[b].p, (const char *)x + (size_t)r * in_bytes, in_bytes, cudaMemcpyHostToDevice, q->s_h2d));
        if (y) OX_CUDA(cudaMemcpyAsync(q->pipe_y[b].p, (const char *)y + (size_t)r * in_bytes, in_bytes, cudaMemcpyHostToDevice, q->s_h2d));
        OX_CUDA(cudaEventRecord(q->ev_in[b], q->s_h2d));
        OX_CUDA(cudaStreamWaitEvent(g_stream, q->ev_in[b], 0));
        if (r >= 2) OX_CUDA(cudaStreamWaitEvent(g_stream, q->ev_out[b], 0));    // slot b's previous result has left
        st = q->dtype == OX_F64
                 ? reconstruct_T<double, double2>(q, q->pipe_x[b].p, y ? q->pipe_y[b].p : nullptr, OX_DEVICE, 1, already_ft, return_ft,
                                                  accumulate_meanfield, q->pipe_out[b].p, OX_DEVICE)
                 : reconstruct_T<float, float2>(q, q->pipe_x[b].p, y ? q->pipe_y[b].p : nullptr, OX_DEVICE, 1, already_ft, return_ft,
                                                accumulate_meanfield, q->pipe_out[b].p, OX_DEVICE);
        if (st != OX_OK) break;
        OX_CUDA(cudaEventRecord(q->ev_c[b], g_stream));
        OX_CUDA(cudaStreamWaitEvent(q->s_d2h, q->ev_c[b], 0));
        OX_CUDA(cudaMemcpyAsync((char *)kappa_out + (size_t)r * out_bytes, q->pipe_out[b].p, out_bytes, cudaMemcpyDeviceToHost, q->s_d2h));
        OX_CUDA(cudaEventRecord(q->ev_out[b], q->s_d2h));
      }
      cudaStreamSynchronize(q->s_h2d);
      cudaStreamSynchronize(q->s_d2h);
      OX_CUDA(cudaStreamSynchronize(g_stream));
      if (st == OX_OK && accumulate_meanfield) {
        add_count_kernel<<<1, 1, 0, g_stream>>>(q->mf_count(), (double)nbatch);
        OX_KERNEL_CHECK();
      }
      return st;
    }
  }
  int st = q->dtype == OX_F64
               ? reconstruct_T<double, double2>(q, x, y, where, nbatch, already_ft, return_ft, accumulate_meanfield, kappa_out, out_where)
               : reconstruct_T<float, float2>(q, x, y, where, nbatch, already_ft, return_ft, accumulate_meanfield, kappa_out, out_where);
  if (st == OX_OK && accumulate_meanfield) {
    add_count_kernel<<<1, 1, 0, g_stream>>>(q->mf_count(), (double)nbatch);
    OX_KERNEL_CHECK();
  }
  return st;
}

}  // extern "C"
