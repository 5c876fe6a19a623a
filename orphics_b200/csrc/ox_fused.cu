// Fused hand-written path for sim -> FFT -> power2d -> bin2D on power-of-two maps.
//
// cuFFT's 2-D real transforms run at ~50% of their own two-pass HBM bound on B200
// (40 us per 2048^2 fp64 transform; tools/ubench/fft_layouts.cu), which alone caps the
// pipeline at ~40% of its roofline.  Here the four 1-D FFT passes are hand-written
// (ox_fft.cuh: register-resident radix-16 Stockham stages exchanging through shared memory)
// and fused with their producers and consumers, so a map costs three kernels and ~4.s.N
// bytes of HBM traffic instead of ten passes:
//
//   K_A fused_sim_col : Philox/Box-Muller noise x covsqrt [x EB->QU] -> Hermitian k_h written
//                        to shared memory -> inverse FFT along y -> last stage stores straight
//                        to HBM, TRANSPOSED: Ht[ix][y] (contiguous, coalesced)
//   K_B fused_row     : tile of R rows: Hermitian packing fused into the first stage of the
//                        half-length inverse FFT along x -> last stage multiplies by the taper
//                        (and optionally stores the real map) -> forward FFT -> unpacking fused
//                        into the transposed store Ht'[ix][y]
//   K_C fused_col_bin : first stage loads a column from HBM -> forward FFT along y -> last stage
//                        feeds |k|^2 straight into the deterministic annular binning
//                        (T-only) or via shared memory for the six TEB spectra
//
// The intermediate layout is the half plane transposed, [plane][ix = 0..Nx/2][iy = 0..Ny),
// so K_A writes and K_C reads whole columns contiguously and only K_B touches R x 16 B
// segments (R = 4 rows -> 64 B, sector aligned).
#include <math.h>

#include "ox_common.cuh"
#include "ox_fft.cuh"
#include "ox_fused_kernels.cuh"
#include "ox_rng.cuh"

using namespace ox;
using namespace oxfft;
using namespace oxk;

// compile-time tuning knobs (defaults = the measured best on B200, profiles/r01_variants.txt)
#ifndef OX_KA_REGS
#define OX_KA_REGS 100 // register cap per thread asked of K_A (T-only), 0 = none: 100 -> five CTAs of 128 threads per SM
#endif
#ifndef OX_KA3_SEQ
#define OX_KA3_SEQ 0    // three-component K_A: 1 = one CTA of LY/16 threads transforms the components one after the other
#endif
#ifndef OX_KA3_UNROLL
#define OX_KA3_UNROLL 2  // unroll factor of the pixel loop of the three-component K_A
#endif
#ifndef OX_KA3_MINB
#define OX_KA3_MINB 2   // resident CTAs per SM asked of the three-component K_A (384 threads): 2 -> 80 registers, the
                        // 136 B of spills cost less than a lone CTA whose noise and FFT phases cannot overlap
#endif
#ifndef OX_KA_ROLLED
#define OX_KA_ROLLED 4   // unroll factor of the noise loop of K_A (0 = 16 pixels straight-line into registers)
#endif
#ifndef OX_KC_REGS
#define OX_KC_REGS 128   // register cap per thread asked of K_C (T-only), 0 = none
#endif
#ifndef OX_KC_COLS
#define OX_KC_COLS 8    // columns per CTA of K_C
#endif

namespace {

// min resident CTAs per SM that caps the registers per thread at `regs` (0 = no cap)
__host__ __device__ constexpr int ka_threads(int nc, int ly) { return (nc > 1 && OX_KA3_SEQ) ? ly / 16 : nc * (ly / 16); }
constexpr int minb_for(int threads, int regs) { return regs > 0 && 65536 / (threads * regs) > 1 ? 65536 / (threads * regs) : 1; }

// ---- shared helpers -------------------------------------------------------------------
// 1/x for normal positive x to ~1 ulp without the IEEE division's slow-path branch (which would split
// the straight-line pixel code): MUFU seed (2^-20) + two Newton steps
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(fma(-x, r, 1.0), r, r);
  r = fma(fma(-x, r, 1.0), r, r);
  return r;
}

// cos/sin of the Q/U <-> E/B rotation angle 2 phi_l (enmap.queb_rotmat; healpix sign unless iau)
__device__ __forceinline__ void rot_cs(double y, double x, double sgn, double &c, double &s) {
  const double l2 = y * y + x * x;
  const bool nz = l2 > 0.0;
  const double inv = fast_rcp(nz ? l2 : 1.0);
  c = nz ? (y * y - x * x) * inv : 1.0;
  s = sgn * (-2.0 * x * y) * inv;
}

// deterministic in-warp reduce-by-key: afterwards the lowest lane of every group of equal keys
// holds the group's sums (fixed shuffle tree, see ox_binner.cu)
template <int NV>
__device__ __forceinline__ void reduce_peers(unsigned peers, double (&v)[NV]) {
  const int lane = threadIdx.x & 31;
  unsigned rel = __popc(peers & ((1u << lane) - 1u));
  peers &= (lane == 31) ? 0u : (0xfffffffeu << lane);
  while (__any_sync(0xffffffffu, peers)) {
    int next = __ffs(peers);
    int src = next ? next - 1 : lane;
#pragma unroll
    for (int k = 0; k < NV; k++) {
      double t = __shfl_sync(0xffffffffu, v[k], src);
      if (next) v[k] += t;
    }
    unsigned done = rel & 1u;
    peers &= __ballot_sync(0xffffffffu, !done);
    rel >>= 1;
  }
}

// ---- K_A -------------------------------------------------------------------------------
template <typename T>
struct SimColArgs {
  const T *covT;           // [NC][NC][nx][ny] covsqrt transposed (column ix contiguous in iy)
  const double *noise;     // host-noise mode: [nsim][2][NC][ny][nx] natural layout
  const long long *seeds;  // [nsim]
  const double *ly, *lx;
  const typename V2<T>::type *tw;  // exp(-2 pi i j / tw_len)
  int tw_len;
  int ny, nx, mx;  // mx = nx/2
  int rot, cov_symmetric;
  double scale, rot_sgn;
};

// Hermitian part of the simulated Fourier field at pixel (iy, ix) of the half plane:
// z[c] = 1/2 [k_c(p) + conj k_c(p')] / sqrt(N), k = covsqrt . r [rotated EB -> QU]
// (MapGen.get_map, maps.py:1576-1587, followed by enmap.ifft(...).real)
// PATH >= 1 (interior): the column is neither ix = 0 nor the Nyquist column, so no pixel is its own
// mirror image and the canonical member of each pair {p, p'} is decided by the row alone.
// PATH == 2 (interior, covsqrt(p') = covsqrt(p), Hermitian noise): k(p') = conj k(p) before the
// rotation and the rotation at p' has the same cosine and the same sine -- except on the Nyquist row,
// where the sine flips -- so 1/2 [k(p) + conj k(p')] = rotation of k(p) with the sine zeroed on that
// row: the mirrored pixel is never touched (half the mixing, one rotation instead of two)
template <typename T, int NC, int MODE, int PATH>
__device__ __forceinline__ void sim_pixel(const SimColArgs<T> &a, const oxrng::PhiloxKeys &keys, const double2 *logtab,
                                          const double *noise_sim, int ix, int mxp, int iy, double h,
                                          typename V2<T>::type (&z)[NC]) {
  const int my = iy ? a.ny - iy : 0;
  const unsigned n = (unsigned)a.ny * (unsigned)a.nx;
  const unsigned p = (unsigned)iy * a.nx + ix, q = (unsigned)my * a.nx + mxp;
  const unsigned tp = (unsigned)ix * a.ny + iy, tq = (unsigned)mxp * a.ny + my;
  constexpr bool INTERIOR = PATH >= 1;
  constexpr bool ONESIDED = PATH == 2 && MODE == OX_NOISE_PHILOX_HERMITIAN;
  T covp[NC * NC], covq[NC * NC];
#pragma unroll
  for (int e = 0; e < NC * NC; e++) {
    covp[e] = a.covT[(size_t)e * n + tp];
    covq[e] = (ONESIDED || a.cov_symmetric) ? covp[e] : a.covT[(size_t)e * n + tq];
  }
  double pr[NC], pi[NC], qr[NC], qi[NC];
  if (MODE == OX_NOISE_HOST) {
#pragma unroll
    for (int c = 0; c < NC; c++) {
      pr[c] = noise_sim[(size_t)c * n + p];
      pi[c] = noise_sim[(size_t)(NC + c) * n + p];
      qr[c] = noise_sim[(size_t)c * n + q];
      qi[c] = noise_sim[(size_t)(NC + c) * n + q];
    }
  } else if (MODE == OX_NOISE_PHILOX) {
#pragma unroll
    for (int c = 0; c < NC; c++) {
      oxrng::box_muller<T>(oxrng::philox4x32_10(make_uint4(p, 0u, c, 0u), keys), logtab, pr[c], pi[c]);
      // (the self-conjugate pixels draw twice: same counter, same value)
      oxrng::box_muller<T>(oxrng::philox4x32_10(make_uint4(q, 0u, c, 0u), keys), logtab, qr[c], qi[c]);
    }
  } else {
    const bool conj_me = INTERIOR ? iy > (a.ny >> 1) : q < p, selfc = INTERIOR ? false : q == p;
    const unsigned canon = conj_me ? q : p;
#pragma unroll
    for (int c = 0; c < NC; c++) {
      double n1, n2;
      oxrng::box_muller<T>(oxrng::philox4x32_10(make_uint4(canon, 0u, c, 1u), keys), logtab, n1, n2);
      pr[c] = selfc ? n1 : n1 * 0.70710678118654752440;
      pi[c] = selfc ? 0.0 : (conj_me ? -n2 : n2) * 0.70710678118654752440;
      qr[c] = pr[c];
      qi[c] = -pi[c];
    }
  }
  if constexpr (ONESIDED) {
    double kr[NC], ki[NC];
#pragma unroll
    for (int i = 0; i < NC; i++) {
      double sr = 0, si = 0;
#pragma unroll
      for (int j = 0; j < NC; j++) {
        const double cp = (double)covp[i * NC + j];
        sr += cp * pr[j]; si += cp * pi[j];
      }
      kr[i] = sr; ki[i] = si;
    }
    if (NC == 3 && a.rot) {
      double c, sn;
      rot_cs(a.ly[iy], a.lx[ix], a.rot_sgn, c, sn);
      if (2 * iy == a.ny) sn = 0.0;
      const double t1 = c * kr[1] + sn * kr[2], t2 = c * ki[1] + sn * ki[2];
      const double t3 = -sn * kr[1] + c * kr[2], t4 = -sn * ki[1] + c * ki[2];
      kr[1] = t1; ki[1] = t2; kr[2] = t3; ki[2] = t4;
    }
    const double h2 = 2.0 * h;
#pragma unroll
    for (int c = 0; c < NC; c++) {
      z[c].x = (T)(h2 * kr[c]);
      z[c].y = (T)(h2 * ki[c]);
    }
  } else {
  // k(p) = covsqrt(p) r(p), k(p') = covsqrt(p') r(p')
  double kpr[NC], kpi[NC], kqr[NC], kqi[NC];
#pragma unroll
  for (int i = 0; i < NC; i++) {
    double sr = 0, si = 0, ur = 0, ui = 0;
#pragma unroll
    for (int j = 0; j < NC; j++) {
      double cp = (double)covp[i * NC + j], cq = (double)covq[i * NC + j];
      sr += cp * pr[j]; si += cp * pi[j];
      ur += cq * qr[j]; ui += cq * qi[j];
    }
    kpr[i] = sr; kpi[i] = si; kqr[i] = ur; kqi[i] = ui;
  }
  if (NC == 3 && a.rot) {
    // inverse queb_rotmat: [Q;U] = [[c, s],[-s, c]] [E;B], each pixel with its own angle
    double c, sn;
    rot_cs(a.ly[iy], a.lx[ix], a.rot_sgn, c, sn);
    double t1 = c * kpr[1] + sn * kpr[2], t2 = c * kpi[1] + sn * kpi[2];
    double t3 = -sn * kpr[1] + c * kpr[2], t4 = -sn * kpi[1] + c * kpi[2];
    kpr[1] = t1; kpi[1] = t2; kpr[2] = t3; kpi[2] = t4;
    rot_cs(a.ly[my], a.lx[mxp], a.rot_sgn, c, sn);
    t1 = c * kqr[1] + sn * kqr[2]; t2 = c * kqi[1] + sn * kqi[2];
    t3 = -sn * kqr[1] + c * kqr[2]; t4 = -sn * kqi[1] + c * kqi[2];
    kqr[1] = t1; kqi[1] = t2; kqr[2] = t3; kqi[2] = t4;
  }
#pragma unroll
  for (int c = 0; c < NC; c++) {
    z[c].x = (T)(h * (kpr[c] + kqr[c]));
    z[c].y = (T)(h * (kpi[c] - kqi[c]));
  }
  }
}

template <typename T, int LY, int NC, int MODE>
__global__ void __launch_bounds__(ka_threads(NC, LY), (NC == 1 ? (LY == 2048 ? minb_for(LY / 16, OX_KA_REGS) : 1) : (LY == 2048 ? OX_KA3_MINB : 1)))  // (the caps spill at other lengths)
fused_sim_col_kernel(SimColArgs<T> a, typename V2<T>::type *__restrict__ Ht /*[nsim][NC][mx+1][ny]*/) {
  typedef typename V2<T>::type T2;
  typedef BlockFFT<T, LY> FFT;
  constexpr int NT = FFT::NT, NTHREADS = ka_threads(NC, LY), PS = padded_size(LY);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T2 *s = reinterpret_cast<T2 *>(smem_raw);                                        // [NC][PS]
  double2 *logtab = reinterpret_cast<double2 *>(smem_raw + sizeof(T2) * NC * PS);  // [129]
  // NC == 1: columns vary fastest (blockIdx.x = ix).  NC > 1: realisations vary fastest (blockIdx.x = sim), so that the CTAs
  // resident together work on the SAME column of different realisations: the 3x3 covsqrt of a 2048^2 map is 302 MB, more
  // than L2, and with columns fastest it streamed from DRAM again for every realisation (ncu: 2.7 GB read per 16
  // realisations next to 1.6 GB written)
  const int ix = NC == 1 ? blockIdx.x : blockIdx.y, sim = NC == 1 ? blockIdx.y : blockIdx.x, tid = threadIdx.x;
  const int mxp = ix ? a.nx - ix : 0;  // mirrored column
  const double h = 0.5 * a.scale;
  oxrng::PhiloxKeys keys;
  const double *noise_sim = nullptr;
  if (MODE == OX_NOISE_HOST) {
    noise_sim = a.noise + (size_t)sim * 2 * NC * a.ny * a.nx;
  } else {
    keys.init((unsigned long long)a.seeds[sim]);
    oxrng::load_log_table(logtab, tid, NTHREADS);
    __syncthreads();
  }
  const int f = tid / NT, u = tid - f * NT;
  typename FFT::Twiddles tws;
  tws.init(a.tw, a.tw_len / LY, u);
  GlobalStore<T2> st{Ht + (((size_t)sim * NC + f) * (a.mx + 1) + ix) * a.ny};
  if (NC == 1) {
    // T-only: a thread generates exactly the 16 inputs u + m*NT of its first radix-16 butterfly:
    // straight-line code (16 independent Philox/Box-Muller chains for the scheduler to interleave),
    // no staging in shared memory
    T2 v[16];
    const bool interior = ix != 0 && ix != a.mx;
#define OX_GEN16(PATH)                                                                        \
  _Pragma("unroll") for (int m = 0; m < 16; m++) {                                            \
    T2 z[NC];                                                                                 \
    sim_pixel<T, NC, MODE, PATH>(a, keys, logtab, noise_sim, ix, mxp, u + m * NT, h, z);      \
    v[m] = z[0];                                                                              \
  }
#if OX_KA_ROLLED
    // The 16 values are generated four at a time in a rolled loop and parked in the thread's OWN shared-memory
    // slots (u + m*NT are exactly the inputs of its first butterfly, so no barrier is needed before reading
    // them back).  Generating all 16 straight-line into registers (OX_KA_ROLLED=0) costs 186 registers and
    // 54 KB of instructions per warp with no reuse: ncu showed 0.3 warps per issue stalled on instruction
    // fetch; the rolled loop is 12% faster (profiles/r01_variants.txt).
    {
      constexpr int UR = OX_KA_ROLLED;
#define OX_ROLL16(PATH)                                                                       \
  _Pragma("unroll UR") for (int m = 0; m < 16; m++) {                                         \
    T2 z[NC];                                                                                 \
    sim_pixel<T, NC, MODE, PATH>(a, keys, logtab, noise_sim, ix, mxp, u + m * NT, h, z);      \
    s[pad(u + m * NT)] = z[0];                                                                \
  }
      if (interior && a.cov_symmetric && MODE == OX_NOISE_PHILOX_HERMITIAN) {
        OX_ROLL16(2)
      } else if (interior) {
        OX_ROLL16(1)
      } else {
        OX_ROLL16(0)
      }
#undef OX_ROLL16
      SmemLoad<T2> lds{s};
      FFT::template run<+1, true, false>(s, tws, u, 0, lds, st);
      return;
    }
#endif
    if (interior && a.cov_symmetric && MODE == OX_NOISE_PHILOX_HERMITIAN) {
      OX_GEN16(2)
    } else if (interior) {
      OX_GEN16(1)
    } else {
      OX_GEN16(0)
    }
#undef OX_GEN16
    RegLoad<T2> ld{v};
    FFT::template run<+1, false, false>(s, tws, u, 0, ld, st);
  } else if (OX_KA3_SEQ) {
    // experiment: LY/16 threads generate all components of their 16 pixels into their own slots of the NC
    // buffers, then transform the components one after the other
    const bool fast = ix != 0 && ix != a.mx && a.cov_symmetric && MODE == OX_NOISE_PHILOX_HERMITIAN;
#pragma unroll 2
    for (int m = 0; m < 16; m++) {
      T2 z[NC];
      if (fast) sim_pixel<T, NC, MODE, 2>(a, keys, logtab, noise_sim, ix, mxp, tid + m * NT, h, z);
      else sim_pixel<T, NC, MODE, 0>(a, keys, logtab, noise_sim, ix, mxp, tid + m * NT, h, z);
#pragma unroll
      for (int c = 0; c < NC; c++) s[c * PS + pad(tid + m * NT)] = z[c];
    }
    typename FFT::Twiddles tw2;
    tw2.init(a.tw, a.tw_len / LY, tid);
    for (int c = 0; c < NC; c++) {
      SmemLoad<T2> ld{s + c * PS};
      GlobalStore<T2> stc{Ht + (((size_t)sim * NC + c) * (a.mx + 1) + ix) * a.ny};
      FFT::template run<+1, true, false>(s + c * PS, tw2, tid, 0, ld, stc);
    }
  } else {
    constexpr int UR3 = OX_KA3_UNROLL;
    if (ix != 0 && ix != a.mx && a.cov_symmetric && MODE == OX_NOISE_PHILOX_HERMITIAN) {
#pragma unroll UR3
      for (int iy = tid; iy < LY; iy += NTHREADS) {
        T2 z[NC];
        sim_pixel<T, NC, MODE, 2>(a, keys, logtab, noise_sim, ix, mxp, iy, h, z);
#pragma unroll
        for (int c = 0; c < NC; c++) s[c * PS + pad(iy)] = z[c];
      }
    } else {
#pragma unroll UR3
      for (int iy = tid; iy < LY; iy += NTHREADS) {
        T2 z[NC];
        sim_pixel<T, NC, MODE, 0>(a, keys, logtab, noise_sim, ix, mxp, iy, h, z);
#pragma unroll
        for (int c = 0; c < NC; c++) s[c * PS + pad(iy)] = z[c];
      }
    }
    __syncthreads();
    const int bar = (NT % 32 == 0) ? 1 + f : 0;
    SmemLoad<T2> ld{s + f * PS};
    FFT::template run<+1, true, false>(s + f * PS, tws, u, bar, ld, st);
  }
}

// ---- K_C -------------------------------------------------------------------------------
template <typename T>
struct ColBinArgs {
  const typename V2<T>::type *H;  // [nbatch][NC][mx+1][ny]
  const uint16_t *idxT;           // [mx+1][ny]: slot | 0x8000 if Hermitian weight 2
  const double *ly, *lx;
  const typename V2<T>::type *tw;
  int tw_len;
  int ny, mx, nslots, cols_per_block, rot;
  double rot_sgn;
};

// last stage of the T-only column transform: |k|^2 x Hermitian weight straight into the
// warp-private slot sums.  Slots 0 (at/below the first edge) and nslots-1 (above the last edge)
// are dropped by bin2D's [1:-1] (stats.py:796-797): nothing is accumulated for them.
template <typename T2>
struct BinStore {
  const unsigned short (&raws)[16];
  double *mine;
  unsigned trash, last;
  int lane;
  __device__ __forceinline__ BinStore(const unsigned short (&r)[16], double *m, unsigned t, unsigned l, int ln)
      : raws(r), mine(m), trash(t), last(l), lane(ln) {}
  __device__ __forceinline__ void operator()(int, T2 z, int m) const {
    const unsigned raw = raws[m];
    unsigned key = raw & 0x7fffu;
    if (key == 0u || key == last) key = trash;
    if (__all_sync(0xffffffffu, key == trash)) return;
    double v[1];
    v[0] = key != trash ? ((double)z.x * (double)z.x + (double)z.y * (double)z.y) * ((raw & 0x8000u) ? 2.0 : 1.0) : 0.0;
    unsigned peers = __match_any_sync(0xffffffffu, key);
    bool leader = (__ffs(peers) - 1) == lane;
    reduce_peers<1>(peers, v);
    if (leader) mine[key] += v[0];
    __syncwarp();
  }
};

template <typename T, int LY, int NC>
__global__ void __launch_bounds__(NC *(LY / 16), (NC == 1 ? minb_for(LY / 16, OX_KC_REGS) : 1))
fused_col_bin_kernel(ColBinArgs<T> a, double *__restrict__ partial /*[nbatch][gridDim.x][NS][nslots]*/) {
  typedef typename V2<T>::type T2;
  typedef BlockFFT<T, LY> FFT;
  constexpr int NT = FFT::NT, NTHREADS = NC * NT, PS = padded_size(LY), NS = NC * (NC + 1) / 2, NWARPS = NTHREADS / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T2 *s = reinterpret_cast<T2 *>(smem_raw);                                           // [NC][PS]
  double *bins = reinterpret_cast<double *>(smem_raw + sizeof(T2) * (size_t)NC * PS);  // [NWARPS][NS][nslots+1]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int stride = a.nslots + 1;
  double *mine = bins + (size_t)warp * NS * stride;
  for (int i = lane; i < NS * stride; i += 32) mine[i] = 0.0;
  const long long b = blockIdx.y;
  const int f = tid / NT, u = tid - f * NT;
  const int ix0 = blockIdx.x * a.cols_per_block;
  const int ix1 = min(ix0 + a.cols_per_block, a.mx + 1);
  typename FFT::Twiddles tws;
  tws.init(a.tw, a.tw_len / LY, u);
  const unsigned trash = a.nslots, last = a.nslots - 1;
  const int bar = (NC > 1 && NT % 32 == 0) ? 1 + f : 0;
  for (int ix = ix0; ix < ix1; ix++) {
    const uint16_t *idx = a.idxT + (long long)ix * a.ny;
    GlobalLoad<T2> ld{a.H + ((b * NC + f) * (a.mx + 1) + ix) * (long long)a.ny};
    __syncthreads();  // the previous column's readers are done with s
    if (NC == 1) {
      // slot indices of this thread's 16 outputs (u + m*NT), fetched before the transform
      unsigned short raws[16];
#pragma unroll
      for (int m = 0; m < 16; m++) raws[m] = idx[u + m * NT];
      BinStore<T2> st(raws, mine, trash, last, lane);
      FFT::template run<-1, false, false>(s, tws, u, 0, ld, st);
    } else {
      SmemStore<T2> st{s + f * PS};
      FFT::template run<-1, false, true>(s + f * PS, tws, u, bar, ld, st);
      __syncthreads();
      const double x = a.lx[ix];
      for (int base = 0; base < LY; base += NTHREADS) {
        const int iy = base + tid;
        unsigned raw = iy < LY ? idx[iy] : 0;
        unsigned key = raw & 0x7fffu;
        if (iy >= LY || key == 0u || key == last) key = trash;
        if (__all_sync(0xffffffffu, key == trash)) continue;
        double v[NS];
#pragma unroll
        for (int q = 0; q < NS; q++) v[q] = 0.0;
        if (key != trash) {
          const double w = (raw & 0x8000u) ? 2.0 : 1.0;
          double re[NC], im[NC];
#pragma unroll
          for (int c = 0; c < NC; c++) {
            T2 z = s[c * PS + pad(iy)];
            re[c] = (double)z.x;
            im[c] = (double)z.y;
          }
          double c = 1.0, sn = 0.0;
          // Nyquist row of an interior column: the mirrored pixel has the opposite lx, hence the opposite
          // sine of the rotation angle -- its E/B power is accumulated with -sn instead of doubling p's
          bool nyq_pair = false;
          if (NC == 3 && a.rot) {
            rot_cs(a.ly[iy], x, a.rot_sgn, c, sn);
            nyq_pair = (2 * iy == LY) && (raw & 0x8000u);
          }
          const int nterm = nyq_pair ? 2 : 1;
          const double wt = nyq_pair ? 1.0 : w;
          for (int term = 0; term < nterm; term++) {
            double pr[NC], pi[NC];
#pragma unroll
            for (int c2 = 0; c2 < NC; c2++) { pr[c2] = re[c2]; pi[c2] = im[c2]; }
            if (NC == 3 && a.rot) {
              const double st = term ? -sn : sn;
              pr[1] = c * re[1] - st * re[2]; pi[1] = c * im[1] - st * im[2];
              pr[2] = st * re[1] + c * re[2]; pi[2] = st * im[1] + c * im[2];
            }
            int q = 0;
#pragma unroll
            for (int i2 = 0; i2 < NC; i2++)
#pragma unroll
              for (int j = i2; j < NC; j++) v[q++] += (pr[i2] * pr[j] + pi[i2] * pi[j]) * wt;
          }
        }
        unsigned peers = __match_any_sync(0xffffffffu, key);
        bool leader = (__ffs(peers) - 1) == lane;
        reduce_peers<NS>(peers, v);
        if (leader) {
#pragma unroll
          for (int q = 0; q < NS; q++) mine[q * stride + key] += v[q];
        }
        __syncwarp();
      }
    }
  }
  __syncthreads();
  double *out = partial + ((size_t)b * gridDim.x + blockIdx.x) * NS * a.nslots;
  for (int i = tid; i < NS * a.nslots; i += NTHREADS) {
    int sp = i / a.nslots, slot = i - sp * a.nslots;
    double acc = 0.0;
    for (int w = 0; w < NWARPS; w++) acc += bins[(size_t)w * NS * stride + sp * stride + slot];
    out[i] = acc;
  }
}

// ---- set-up kernels ----------------------------------------------------------------------
template <typename T>
__global__ void transpose_cov_kernel(const T *__restrict__ in, T *__restrict__ out, int ny, int nx, int nmat) {
  // out[m][ix][iy] = in[m][iy][ix]
  __shared__ T tile[32][33];
  const int m = blockIdx.z;
  int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 32 + threadIdx.y;
  const T *src = in + (long long)m * ny * nx;
  T *dst = out + (long long)m * ny * nx;
  for (int j = 0; j < 32; j += 8)
    if (x < nx && y + j < ny) tile[threadIdx.y + j][threadIdx.x] = src[(long long)(y + j) * nx + x];
  __syncthreads();
  x = blockIdx.y * 32 + threadIdx.x;  // iy
  y = blockIdx.x * 32 + threadIdx.y;  // ix
  for (int j = 0; j < 32; j += 8)
    if (x < ny && y + j < nx) dst[(long long)(y + j) * ny + x] = tile[threadIdx.x][threadIdx.y + j];
}

template <typename T>
__global__ void cov_symmetry_kernel(const T *__restrict__ cov, int ny, int nx, int nmat, int *__restrict__ mismatch) {
  long long n = (long long)ny * nx;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n * nmat; i += (long long)gridDim.x * blockDim.x) {
    long long m = i / n, p = i - m * n;
    int iy = (int)(p / nx), ix = (int)(p - (long long)iy * nx);
    int my = iy ? ny - iy : 0, mx = ix ? nx - ix : 0;
    if (cov[i] != cov[m * n + (long long)my * nx + mx]) atomicAdd(mismatch, 1);
  }
}

__global__ void transpose_idxh_kernel(const uint16_t *__restrict__ idxh, int ny, int nxh, uint16_t *__restrict__ out) {
  long long n = (long long)ny * nxh;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int ix = (int)(i / ny), iy = (int)(i - (long long)ix * ny);
    out[i] = idxh[(long long)iy * nxh + ix];
  }
}

// one warp per (map, spectrum, bin): lanes stride over the blocks, fixed shuffle tree
__global__ void bandpower_finalize2_kernel(const double *__restrict__ partial, int nblk, int ns, int nslots,
                                           const double *__restrict__ count, double normfact, double *__restrict__ bp) {
  const int nbins = nslots - 2;
  int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  long long m = blockIdx.y;
  if (gw >= ns * nbins) return;
  int sp = gw / nbins, bin = gw - sp * nbins;
  const double *p = partial + (size_t)m * nblk * ns * nslots + (size_t)sp * nslots + bin + 1;
  double acc = 0.0;
  for (int b = lane; b < nblk; b += 32) acc += p[(size_t)b * ns * nslots];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) bp[(size_t)m * ns * nbins + gw] = (acc * normfact) / count[bin + 1];
}

template <typename T, int LY, int NC, int MODE>
int launch_sim_col_mode(SimColArgs<T> &a, void *Ht, int nsim) {
  typedef typename V2<T>::type T2;
  size_t smem = sizeof(T2) * NC * padded_size(LY) + sizeof(double2) * oxrng::LOG_TABLE_ENTRIES;
  OX_REQUIRE(smem <= SMEM_MAX, "fused sim: column of %d x %d comps needs %zu B of shared memory", LY, NC, smem);
  auto k = fused_sim_col_kernel<T, LY, NC, MODE>;
  OX_TRY(set_smem(k, smem, NC == 1 || OX_KA3_MINB > 1));
  dim3 grid(a.mx + 1, nsim);
  if (NC > 1) grid = dim3(nsim, a.mx + 1);
  k<<<grid, ka_threads(NC, LY), smem, g_stream>>>(a, (T2 *)Ht);
  OX_KERNEL_CHECK();
  return OX_OK;
}

template <typename T, int LY, int NC>
int launch_sim_col(SimColArgs<T> &a, void *Ht, int nsim, int mode) {
  switch (mode) {
    case OX_NOISE_HOST: return launch_sim_col_mode<T, LY, NC, OX_NOISE_HOST>(a, Ht, nsim);
    case OX_NOISE_PHILOX: return launch_sim_col_mode<T, LY, NC, OX_NOISE_PHILOX>(a, Ht, nsim);
    case OX_NOISE_PHILOX_HERMITIAN: return launch_sim_col_mode<T, LY, NC, OX_NOISE_PHILOX_HERMITIAN>(a, Ht, nsim);
  }
  set_error("fused sim: unknown noise mode %d", mode);
  return OX_ERR_INVALID;
}

template <typename T, int LY, int NC>
int launch_col_bin(ColBinArgs<T> &a, double *partial, int nbatch, int nblk) {
  typedef typename V2<T>::type T2;
  constexpr int NTHREADS = NC * (LY / 16), NS = NC * (NC + 1) / 2;
  size_t smem = sizeof(T2) * NC * padded_size(LY) + sizeof(double) * ((NTHREADS + 31) / 32) * NS * (a.nslots + 1);
  OX_REQUIRE(smem <= SMEM_MAX, "fused bin: %d slots x %d spectra need %zu B of shared memory", a.nslots, NS, smem);
  auto k = fused_col_bin_kernel<T, LY, NC>;
  OX_TRY(set_smem(k, smem));
  dim3 grid(nblk, nbatch);
  k<<<grid, NTHREADS, smem, g_stream>>>(a, partial);
  OX_KERNEL_CHECK();
  return OX_OK;
}

bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

}  // namespace

// =========================================================================================
namespace ox {

bool fused_supported(int ny, int nx, int ncomp, int dtype) {
  if (!pow2(ny) || !pow2(nx)) return false;
  if (ny < 512 || ny > 8192 || nx < 256 || nx > 8192) return false;  // K_C needs whole warps: ny/16 >= 32
  if (ny == 8192 && ncomp != 1) return false;                        // (three 8192-element columns: cuFFT path)
  if (ncomp != 1 && ncomp != 3) return false;
  size_t es = dtype == OX_F32 ? 8 : 16;
  // K_A / K_C keep ncomp columns in shared memory (plus the slot sums)
  if (es * ncomp * padded_size(ny) + 48 * 1024 > SMEM_MAX) return false;
  return true;
}

int fused_make_twiddles(int len, int dtype, DevBuf &buf) {
  size_t es = dtype == OX_F32 ? 8 : 16;
  OX_TRY(buf.ensure(es * len));
  std::vector<double> h(2 * (size_t)len);
  const long double tau = 6.283185307179586476925286766559005768394L;
  if (len % 8 == 0) {
    // first octant from the library, the rest by the exact symmetries of the circle, so that kernels may rebuild
    // w_{len/4 - j} = (-Im w_j, -Re w_j) etc. from a part of the table and get the same bits (ox_row_tma.cuh)
    const int o = len / 8;
    std::vector<double> c(o + 1), sn(o + 1);
    for (int j = 0; j <= o; j++) {
      long double ang = tau * (long double)j / (long double)len;
      c[j] = (double)cosl(ang);
      sn[j] = (double)sinl(ang);
    }
    sn[o] = c[o];  // cos(pi/4) = sin(pi/4)
    for (int j = 0; j < len; j++) {
      const int q = j / (2 * o), r = j % (2 * o);   // quadrant, position inside it (angle = q pi/2 + 2 pi r/len)
      const double cr = r <= o ? c[r] : sn[2 * o - r], sr = r <= o ? sn[r] : c[2 * o - r];
      double cj, sj;   // cos / sin of 2 pi j / len
      switch (q) {
        case 0: cj = cr; sj = sr; break;
        case 1: cj = -sr; sj = cr; break;
        case 2: cj = -cr; sj = -sr; break;
        default: cj = sr; sj = -cr; break;
      }
      h[2 * j] = cj;
      h[2 * j + 1] = -sj;
    }
    for (auto &v : h)
      if (v == 0.0) v = 0.0;    // (no negative zeros)
  } else {
    for (int j = 0; j < len; j++) {
      long double ang = -tau * (long double)j / (long double)len;
      h[2 * j] = (double)cosl(ang);
      h[2 * j + 1] = (double)sinl(ang);
    }
  }
  if (dtype == OX_F64) {
    OX_CUDA(cudaMemcpyAsync(buf.p, h.data(), es * len, cudaMemcpyHostToDevice, g_stream));
    OX_CUDA(cudaStreamSynchronize(g_stream));
  } else {
    std::vector<float> hf(2 * (size_t)len);
    for (size_t i = 0; i < hf.size(); i++) hf[i] = (float)h[i];
    OX_CUDA(cudaMemcpyAsync(buf.p, hf.data(), es * len, cudaMemcpyHostToDevice, g_stream));
    OX_CUDA(cudaStreamSynchronize(g_stream));
  }
  return OX_OK;
}

template <typename T>
static int prepare_cov_T(ox_simplan *p, FusedState &fs) {
  ox_geometry *g = p->g;
  int nmat = p->ncomp * p->ncomp;
  size_t nel = (size_t)nmat * g->ny * g->nx;
  OX_TRY(fs.covT.ensure(sizeof(T) * nel));
  dim3 grid((g->nx + 31) / 32, (g->ny + 31) / 32, nmat), block(32, 8);
  transpose_cov_kernel<T><<<grid, block, 0, g_stream>>>(p->covsqrt.as<T>(), fs.covT.as<T>(), g->ny, g->nx, nmat);
  OX_KERNEL_CHECK();
  DevBuf mism;
  OX_TRY(mism.ensure(sizeof(int)));
  OX_CUDA(cudaMemsetAsync(mism.p, 0, sizeof(int), g_stream));
  cov_symmetry_kernel<T><<<sm_count() * 8, 256, 0, g_stream>>>(p->covsqrt.as<T>(), g->ny, g->nx, nmat, mism.as<int>());
  OX_KERNEL_CHECK();
  int h = 0;
  OX_CUDA(cudaMemcpyAsync(&h, mism.p, sizeof(int), cudaMemcpyDeviceToHost, g_stream));
  OX_CUDA(cudaStreamSynchronize(g_stream));
  fs.cov_symmetric = (h == 0);
  return OX_OK;
}

int fused_prepare(ox_simplan *s, ox_binner *b, FusedState &fs) {
  ox_geometry *g = s->g;
  fs.tw_len = g->ny > g->nx ? g->ny : g->nx;
  OX_TRY(fused_make_twiddles(fs.tw_len, s->dtype, fs.tw));
  if (s->dtype == OX_F64) OX_TRY(prepare_cov_T<double>(s, fs));
  else OX_TRY(prepare_cov_T<float>(s, fs));
  long long nh = (long long)g->ny * g->nxh;
  OX_TRY(fs.idxT.ensure(sizeof(uint16_t) * nh));
  transpose_idxh_kernel<<<sm_count() * 8, 256, 0, g_stream>>>(b->idxh.as<uint16_t>(), g->ny, g->nxh, fs.idxT.as<uint16_t>());
  OX_KERNEL_CHECK();
  size_t es = elem_size(s->dtype);
  size_t hbytes = 2 * es * (size_t)s->max_batch * s->ncomp * g->nxh * g->ny;
  OX_TRY(fs.Ha.ensure(hbytes));
  OX_TRY(fs.Hb.ensure(hbytes));
  fs.ready = true;
  return OX_OK;
}

template <typename T>
static int fused_run_T(ox_pipeline *pl, int nsim, int noise_mode, const double *noise_dev, int flags, bool keep_maps,
                       cudaEvent_t *ev) {
  typedef typename V2<T>::type T2;
  ox_simplan *s = pl->s;
  ox_geometry *g = s->g;
  FusedState &fs = pl->fused;
  const int nc = s->ncomp;
#define OX_MARK(i) do { if (ev) OX_CUDA(cudaEventRecord(ev[i], g_stream)); } while (0)
  OX_MARK(0);
  SimColArgs<T> sa;
  sa.covT = fs.covT.as<T>();
  sa.noise = noise_dev;
  sa.seeds = s->seeds.as<long long>();
  sa.ly = g->ly.as<double>();
  sa.lx = g->lx.as<double>();
  sa.tw = fs.tw.as<T2>();
  sa.tw_len = fs.tw_len;
  sa.ny = g->ny; sa.nx = g->nx; sa.mx = g->nx / 2;
  sa.rot = (flags & OX_FLAG_ROT) ? 1 : 0;
  sa.cov_symmetric = fs.cov_symmetric ? 1 : 0;
  sa.scale = 1.0 / sqrt((double)g->ny * (double)g->nx);
  sa.rot_sgn = (flags & OX_FLAG_IAU) ? 1.0 : -1.0;
  int st = OX_ERR_UNSUPPORTED;
#define OX_SIMCOL(LY)                                                                     \
  case LY:                                                                                \
    st = nc == 1 ? launch_sim_col<T, LY, 1>(sa, fs.Ha.p, nsim, noise_mode) : launch_sim_col<T, LY, 3>(sa, fs.Ha.p, nsim, noise_mode); \
    break;
  switch (g->ny) {
    OX_SIMCOL(256) OX_SIMCOL(512) OX_SIMCOL(1024) OX_SIMCOL(2048) OX_SIMCOL(4096)
    case 8192:   // (one component only: three 8192-element columns do not fit an SM's shared memory)
      if (nc == 1) st = launch_sim_col<T, 8192, 1>(sa, fs.Ha.p, nsim, noise_mode);
      else set_error("fused path: ny=8192 supports one component");
      break;
    default: set_error("fused path: unsupported ny=%d", g->ny);
  }
#undef OX_SIMCOL
  OX_TRY(st);
  OX_MARK(1);
  stage_mark("K_A sim+col_ifft");
  RowArgs<T> ra;
  ra.Hin = fs.Ha.as<T2>();
  ra.map_in = nullptr;
  ra.Hout = fs.Hb.as<T2>();
  ra.map_out = nullptr;
  if (keep_maps) {
    OX_TRY(s->maps.ensure(elem_size(s->dtype) * (size_t)s->max_batch * nc * g->ny * g->nx));
    ra.map_out = s->maps.as<T>();
  }
  ra.window = pl->has_window ? pl->window.as<T>() : nullptr;
  if (pl->has_window && pl->win_separable) {
    ra.win_x = pl->win_x.as<double>();
    ra.win_y = pl->win_y.as<double>();
  }
  ra.tw = fs.tw.as<T2>();
  ra.tw_len = fs.tw_len;
  ra.ny = g->ny; ra.nx = g->nx; ra.mx = g->nx / 2;
  ra.map_in_group_stride = (long long)g->ny * g->nx;
  typedef RowModes<ROW_IN_H | ROW_OUT_H> PipelineRowModes;
  st = launch_row_any<T>(ra, (long long)nsim * nc, PipelineRowModes());
  OX_TRY(st);
  OX_MARK(2);
  stage_mark("K_B row_c2r+taper+r2c");
  OX_MARK(3);
  OX_MARK(4);
  ColBinArgs<T> ca;
  ca.H = fs.Hb.as<T2>();
  ca.idxT = fs.idxT.as<uint16_t>();
  ca.ly = g->ly.as<double>();
  ca.lx = g->lx.as<double>();
  ca.tw = fs.tw.as<T2>();
  ca.tw_len = fs.tw_len;
  ca.ny = g->ny; ca.mx = g->nx / 2;
  ca.nslots = pl->b->nslots;
  ca.cols_per_block = OX_KC_COLS;
  ca.rot = (flags & OX_FLAG_ROT) ? 1 : 0;
  ca.rot_sgn = sa.rot_sgn;
  int nblk = (g->nxh + ca.cols_per_block - 1) / ca.cols_per_block;
  int ns = nc * (nc + 1) / 2;
  OX_TRY(pl->partial.ensure(sizeof(double) * (size_t)nsim * nblk * ns * pl->b->nslots));
  st = OX_ERR_UNSUPPORTED;
#define OX_COLBIN(LY)                                                                                              \
  case LY:                                                                                                         \
    st = nc == 1 ? launch_col_bin<T, LY, 1>(ca, pl->partial.as<double>(), nsim, nblk)                              \
                 : launch_col_bin<T, LY, 3>(ca, pl->partial.as<double>(), nsim, nblk);                             \
    break;
  switch (g->ny) {
    OX_COLBIN(256) OX_COLBIN(512) OX_COLBIN(1024) OX_COLBIN(2048) OX_COLBIN(4096)
    case 8192:
      if (nc == 1) st = launch_col_bin<T, 8192, 1>(ca, pl->partial.as<double>(), nsim, nblk);
      else set_error("fused path: ny=8192 supports one component");
      break;
    default: set_error("fused path: unsupported ny=%d", g->ny);
  }
#undef OX_COLBIN
  OX_TRY(st);
  int nbins = pl->b->nslots - 2;
  dim3 grid((ns * nbins * 32 + 255) / 256, nsim);
  bandpower_finalize2_kernel<<<grid, 256, 0, g_stream>>>(pl->partial.as<double>(), nblk, ns, pl->b->nslots,
                                                         pl->b->countf.as<double>(), pl->p->normfact, pl->bp.as<double>());
  OX_KERNEL_CHECK();
  OX_MARK(5);
  stage_mark("K_C col_fft+power+bin");
#undef OX_MARK
  return OX_OK;
}

int fused_run(ox_pipeline *pl, int nsim, int noise_mode, const double *noise_dev, int flags, bool keep_maps,
              cudaEvent_t *ev) {
  if (pl->s->dtype == OX_F64) return fused_run_T<double>(pl, nsim, noise_mode, noise_dev, flags, keep_maps, ev);
  return fused_run_T<float>(pl, nsim, noise_mode, noise_dev, flags, keep_maps, ev);
}

}  // namespace ox
