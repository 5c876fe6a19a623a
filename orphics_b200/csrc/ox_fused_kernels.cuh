// Kernels and functors of the hand-written FFT path shared by the sim -> power -> bin pipeline
// (ox_fused.cu) and the quadratic estimator (ox_qe.cu): the row kernel (c2r / taper or product /
// r2c on tiles of R rows of a transposed half plane), a generic column kernel (load functor ->
// FFT along y -> store functor) and their launchers.  See ox_fft.cuh for the FFT engine.
#pragma once
#include <stdlib.h>

#include "ox_common.cuh"
#include "ox_fft.cuh"

#ifndef OX_KB_ROWS64
#define OX_KB_ROWS64 4  // rows per CTA of the row kernel for 16-byte elements (64 B segments)
#endif
#ifndef OX_KB_MINSEG
#define OX_KB_MINSEG 32 // smallest contiguous segment (bytes) of the transposed layout a row tile may touch
#endif
#ifndef OX_KB_WINPF
#define OX_KB_WINPF 0    // full row pass: prefetch the window operands one butterfly ahead (experiment)
#endif
#ifndef OX_KB_MINB_C2R
#define OX_KB_MINB_C2R 2 // the same for the c2r-only row pass (no forward transform: fewer live registers)
#endif
#ifndef OX_KB_MINB
#define OX_KB_MINB 2    // resident CTAs per SM asked of the row kernel (CTAs of <= 256 threads)
#endif

namespace oxk {
using namespace ox;
using namespace oxfft;

// asynchronous global -> shared copy of one element (LDGSTS): no register staging
template <int BYTES>
__device__ __forceinline__ void cp_async(void *smem, const void *gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  if (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// load through the read-only (non-coherent) path: the compiler may reorder it across stores
__device__ __forceinline__ double2 ldg2(const double2 *p) { return __ldg(p); }
__device__ __forceinline__ float2 ldg2(const float2 *p) { return __ldg(p); }

#ifndef OX_STREAM_ST
#define OX_STREAM_ST 0   // 1: planes written once and read by the next kernel are stored with the streaming hint
#endif
#ifndef OX_STREAM_LD
#define OX_STREAM_LD 0   // 1: planes read once are loaded with the streaming hint (no L1 allocation)
#endif
__device__ __forceinline__ void st_once(double2 *p, double2 v) {
  if (OX_STREAM_ST) __stcs(p, v);
  else *p = v;
}
__device__ __forceinline__ void st_once(float2 *p, float2 v) {
  if (OX_STREAM_ST) __stcs(p, v);
  else *p = v;
}
__device__ __forceinline__ double2 ld_once(const double2 *p) { return OX_STREAM_LD ? __ldcs(p) : *p; }
__device__ __forceinline__ float2 ld_once(const float2 *p) { return OX_STREAM_LD ? __ldcs(p) : *p; }

template <typename T2>
struct GlobalStore {
  T2 *dst;
  __device__ __forceinline__ void operator()(int f, T2 v, int) const { st_once(dst + f, v); }
};

// first-stage input held in registers (m is a compile-time constant after unrolling)
template <typename T2>
struct RegLoad {
  const T2 *v;
  __device__ __forceinline__ T2 operator()(int, int m) const { return v[m]; }
};

template <typename T2>
struct GlobalLoad {
  const T2 *src;
  __device__ __forceinline__ T2 operator()(int e, int) const { return ld_once(src + e); }
};

// max_carveout: ask for the largest shared-memory carve-out so that occupancy is set by the registers.  Only
// for kernels that do not lean on L1 for their global loads: K_A gains 6% with five 37 KB CTAs per SM, while the
// row kernel and K_C (window / slot-index / twiddle loads through L1) lose 3% and 30% with it
// (profiles/r01_variants.txt)
template <typename F>
int set_smem(F kernel, size_t bytes, bool max_carveout = false) {
  if (bytes > 48 * 1024) OX_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  if (max_carveout)
    OX_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
  return OX_OK;
}

constexpr size_t SMEM_MAX = 227 * 1024;

// bytes of L2 set aside for persisting accesses (0 = feature off: ORPHX_L2_PERSIST=0, or unsupported);
// set once per process to half of what the device allows
inline size_t l2_persist_budget() {
  static long long budget = -1;
  if (budget < 0) {
    budget = 0;
    const char *env = getenv("ORPHX_L2_PERSIST");
    int dev = 0, maxp = 0;
    if (!(env && env[0] == '0') && cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, dev) == cudaSuccess && maxp > 0) {
      size_t want = (size_t)maxp / 2;
      if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) budget = (long long)want;
    }
    cudaGetLastError();
  }
  return (size_t)budget;
}

// ---- K_B -------------------------------------------------------------------------------
template <typename T>
struct RowArgs {
  const typename V2<T>::type *Hin;  // transposed half plane [plane][mx+1][ny] or null
  const T *map_in;                  // real maps [plane][ny][nx] (used when Hin == null)
  typename V2<T>::type *Hout;       // transposed half plane out or null
  T *map_out;                       // real maps out (before the window) or null
  const T *window;                  // [ny][nx] or null
  const typename V2<T>::type *tw;
  int tw_len;
  int ny, nx, mx;
  // plane p of the grid reads map_in + (p / group) * map_in_group_stride + (p % group) * ny * nx and
  // multiplies by window + (p / group) * win_group_stride: group = 1, strides ny*nx and 0 give one map
  // per plane and a window shared by all planes (the taper); the quadratic estimator multiplies the
  // two gradient legs of a realisation by its third field (group = 2, both strides 3 ny nx)
  int group = 1;
  long long map_in_group_stride = 0, win_group_stride = 0;
  // sub-plane stride inside a group (0 = ny*nx) and an optional second product term (ROW_WIN2, map_in path):
  // input = map_in * window + map_in2 * window2, with map_in2 / window2 addressed like map_in / window
  long long map_in_sub_stride = 0;
  const T *map_in2 = nullptr, *window2 = nullptr;
  // CTA -> (row tile, plane): planes vary fastest (nplanes_fast = number of planes) so that the CTAs resident
  // together work on the SAME rows of different planes and a row of a batch-shared window is fetched from DRAM
  // once per launch instead of once per plane (ncu round 1: 66 MB read per 33.5 MB map with rows fastest);
  // 0 = rows fastest, blockIdx.y = plane (ORPHX_KB_ORDER=rows)
  int nplanes_fast = 0;
  int tile_order = 0;   // persistent TMA row pass only: how its tiles are dealt to the CTAs (ox_row_tma.cuh tile_coords)
  // separable window (full pass only): window[iy][ix] = fl(win_y[iy] * win_x[ix]) exactly, e.g. the reference's
  // cosine taper (maps.py:1893-1920).  The row kernel then reads the 8*nx-byte x profile (L1-resident) and one
  // scalar per row instead of streaming the 8*ny*nx-byte window through L2 for every plane.
  const double *win_x = nullptr, *win_y = nullptr;
};

// first-stage input of the c2r transform: Z[k] = (X[k] + conj X[M-k]) + i e^{+2 pi i k/Nx} (X[k] - conj X[M-k])
template <typename T2, int MX>
struct PackLoad {
  const T2 *row;  // X[0..MX] in padded shared memory
  T2 wu;          // e^{+2 pi i u / Nx} of this thread
  // k = u + m*NT and NT/Nx = 1/32, so e^{+2 pi i k/Nx} = wu * e^{2 pi i m/32}: the 16 factors are
  // compile-time constants after unrolling (no table loads: the kernel is L1/shared-memory bound)
  __device__ __forceinline__ T2 operator()(int k, int m) const {
    constexpr double C32[16] = {1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708,
                                0.70710678118654752440, 0.55557023301960222474, 0.38268343236508977173, 0.19509032201612826785,
                                0.0, -0.19509032201612826785, -0.38268343236508977173, -0.55557023301960222474,
                                -0.70710678118654752440, -0.83146961230254523708, -0.92387953251128675613, -0.98078528040323044913};
    constexpr double S32[16] = {0.0, 0.19509032201612826785, 0.38268343236508977173, 0.55557023301960222474,
                                0.70710678118654752440, 0.83146961230254523708, 0.92387953251128675613, 0.98078528040323044913,
                                1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708,
                                0.70710678118654752440, 0.55557023301960222474, 0.38268343236508977173, 0.19509032201612826785};
    T2 xk = row[pad(k)], xm = row[pad(MX - k)];
    typedef decltype(xk.x) T;
    T2 w;
    w.x = wu.x * (T)C32[m] - wu.y * (T)S32[m];
    w.y = wu.x * (T)S32[m] + wu.y * (T)C32[m];
    T2 sum = cadd(xk, cconj(xm)), dif = csub(xk, cconj(xm));
    return cadd(sum, mul_i<+1>(cmul(w, dif)));
  }
};

// last-stage output of the c2r transform: z[n] = x[2n] + i x[2n+1]; store the map, apply the taper and
// KEEP the element in registers: the outputs u + m*NT of a thread's last stage are exactly the inputs
// of its first forward butterfly, so the real-space row never goes back to shared memory
template <typename T, bool OUT_MAP, bool WIN, bool RUNTIME, int NT>
struct WindowKeep {
  typedef typename V2<T>::type T2;
  T2 *keep;          // registers [16]
  T2 *map_row;       // global or null
  const T2 *win_row; // global or null
  int u;             // thread index within the row
  T2 *w;             // registers [16]: window operands in flight (only a butterfly's worth is live at a time)
  const double2 *winx = nullptr;  // separable window: x profile as pairs (RUNTIME mode), times wy
  double wy = 0.0;
  // RUNTIME: test the pointers per element instead of the compile-time flags.  The branches keep the
  // compiler from issuing all 16 window loads at once, which is what the full c2r -> taper -> r2c pass
  // wants (it has no registers to spare: hoisting spilled 40 B/thread and cost 15%); the one-way passes
  // have free registers and gain 1.5x from the hoisted loads.  For the full pass the engine prefetches
  // the window operands one butterfly ahead instead (pre/use).
  static constexpr bool PREFETCH = OX_KB_WINPF && RUNTIME;
  __device__ __forceinline__ void pre(int m) const {
    if (win_row != nullptr) w[m] = ldg2(win_row + u + m * NT);
  }
  __device__ __forceinline__ void use(int n, T2 z, int m) const {
    if (map_row != nullptr) st_once(map_row + n, z);
    if (win_row != nullptr) {
      z.x *= w[m].x;
      z.y *= w[m].y;
    }
    keep[m] = z;
  }
  __device__ __forceinline__ void operator()(int n, T2 z, int m) const {
    if (RUNTIME ? map_row != nullptr : OUT_MAP) st_once(map_row + n, z);
    if (RUNTIME && winx != nullptr) {
      const double2 p = __ldg(winx + n);
      z.x *= (T)(p.x * wy);   // the window value itself is formed in float64 and rounded once, as numpy forms it
      z.y *= (T)(p.y * wy);
    } else if (RUNTIME ? win_row != nullptr : WIN) {
      T2 w1 = ldg2(win_row + n);
      z.x *= w1.x;
      z.y *= w1.y;
    }
    keep[m] = z;
  }
};

// what a launch of the row kernel does, as compile-time flags: null checks inside the element loops
// are branches that stop the compiler from issuing a thread's 16 window loads together
enum { ROW_IN_H = 1, ROW_OUT_MAP = 2, ROW_WIN = 4, ROW_OUT_H = 8, ROW_WIN2 = 16 };

// R rows per CTA, each a length-MX complex FFT handled by NT = MX/16 threads
template <typename T, int MX, int R, int MODE>
__global__ void __launch_bounds__(R *(MX / 16), (R * (MX / 16) <= 256 ? ((MODE & ROW_OUT_H) ? OX_KB_MINB : OX_KB_MINB_C2R) : 1))
fused_row_kernel(RowArgs<T> a) {
  constexpr bool IN_H = MODE & ROW_IN_H, OUT_MAP = MODE & ROW_OUT_MAP, WIN = MODE & ROW_WIN, OUT_H = MODE & ROW_OUT_H;
  constexpr bool WIN2 = MODE & ROW_WIN2;
  typedef typename V2<T>::type T2;
  typedef BlockFFT<T, MX> FFT;
  constexpr int NT = FFT::NT, NTHREADS = R * NT, PS = padded_size(MX), NX = 2 * MX;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T2 *s = reinterpret_cast<T2 *>(smem_raw);  // [R][PS]
  const int tid = threadIdx.x;
  int tile = blockIdx.x;
  long long plane = blockIdx.y;
  if (a.nplanes_fast) {
    tile = blockIdx.x / a.nplanes_fast;
    plane = blockIdx.x - tile * a.nplanes_fast;
  }
  const int iy0 = tile * R;
  const int f = tid / NT, u = tid - f * NT;
  T2 *row = s + f * PS;
  const int tws_n = a.tw_len / NX;  // stride for exp(-2 pi i k / Nx)
  typename FFT::Twiddles tws;
  tws.init(a.tw, a.tw_len / MX, u);
  // the NT threads of one row synchronise among themselves only (named barriers need whole warps)
  const int bar = (NT % 32 == 0) ? 1 + f : 0;
  const long long rowoff = (long long)(iy0 + f) * MX;
  T2 keep[16];
  constexpr bool RT = IN_H && OUT_H;  // the full pass takes map_out / window as run-time options
  T2 wpf[16];
  WindowKeep<T, OUT_MAP, WIN, RT, NT> wst;
  wst.keep = keep;
  wst.w = wpf;
  wst.u = u;
  wst.map_row = (RT ? a.map_out != nullptr : OUT_MAP) ? reinterpret_cast<T2 *>(a.map_out + plane * (long long)a.ny * NX) + rowoff : nullptr;
  const long long grp = plane / a.group, sub = plane - grp * a.group;
  wst.win_row = (RT ? a.window != nullptr : WIN) ? reinterpret_cast<const T2 *>(a.window + grp * a.win_group_stride) + rowoff : nullptr;
  if (RT && a.win_x != nullptr) {
    wst.winx = reinterpret_cast<const double2 *>(a.win_x);
    wst.wy = a.win_y[iy0 + f];
  }
  if (IN_H) {
    // tile load: for each ix the R rows are R*16 B contiguous in the transposed layout
    const T2 *src = a.Hin + plane * (long long)(MX + 1) * a.ny + iy0;
    // (unrolled with constant strides: a loop that bumps the address registers stalls every
    // iteration on the write-after-read scoreboard of the previous LDGSTS)
    if constexpr ((NTHREADS / R) % 16 == 0 && (MX * R) % NTHREADS == 0) {
      constexpr int IXS = NTHREADS / R;  // ix advance per iteration; pad(ix + IXS) = pad(ix) + pad(IXS)
      const int ix0 = tid / R, r = tid - ix0 * R;
      T2 *sdst = &s[r * PS + pad(ix0)];
      const T2 *gsrc = &src[(long long)ix0 * a.ny + r];
      const long long gstride = (long long)IXS * a.ny;
#pragma unroll
      for (int i = 0; i < MX * R / NTHREADS; i++) cp_async<sizeof(T2)>(sdst + i * pad(IXS), gsrc + i * gstride);
      if (tid < R) cp_async<sizeof(T2)>(&s[tid * PS + pad(MX)], &src[(long long)MX * a.ny + tid]);  // Nyquist column
    } else {
      for (int e = tid; e < (MX + 1) * R; e += NTHREADS) {
        int ix = e / R, r = e - ix * R;
        cp_async<sizeof(T2)>(&s[r * PS + pad(ix)], &src[(long long)ix * a.ny + r]);
      }
    }
    cp_async_wait_all();
    __syncthreads();
    T2 wu = a.tw[u * tws_n];
    wu.y = -wu.y;  // e^{+2 pi i u/Nx}
    PackLoad<T2, MX> ld{row, wu};
    FFT::template run<+1, true, false>(row, tws, u, bar, ld, wst);
  } else {
    // real map rows viewed as z[n] = x[2n] + i x[2n+1]
    const long long sstride = a.map_in_sub_stride ? a.map_in_sub_stride : (long long)a.ny * NX;
    const long long moff = grp * a.map_in_group_stride + sub * sstride + (long long)(iy0 + f) * NX;
    const T2 *src = reinterpret_cast<const T2 *>(a.map_in + moff);
    if (WIN2) {
      const T2 *src2 = reinterpret_cast<const T2 *>(a.map_in2 + moff);
      const T2 *win2 = reinterpret_cast<const T2 *>(a.window2 + grp * a.win_group_stride) + rowoff;
#pragma unroll
      for (int m = 0; m < 16; m++) {
        const int n = u + m * NT;
        const T2 z1 = src[n], w1 = ldg2(wst.win_row + n), z2 = src2[n], w2 = ldg2(win2 + n);
        T2 z;
        z.x = z1.x * w1.x + z2.x * w2.x;
        z.y = z1.y * w1.y + z2.y * w2.y;
        keep[m] = z;
      }
    } else {
#pragma unroll
      for (int m = 0; m < 16; m++) {
        const int n = u + m * NT;
        wst(n, src[n], m);
      }
    }
  }
  if (!OUT_H) return;
  {
    // forward transform fed from registers; IN_SMEM = true: the other threads of the row may still be
    // reading the inverse transform's last exchange, so the first stage synchronises before it writes
    RegLoad<T2> ld{keep};
    SmemStore<T2> st{row};
    FFT::template run<-1, true, true>(row, tws, u, bar, ld, st);
  }
  __syncthreads();
  // transposed store with the r2c unpacking fused in, two outputs per pair (k, M-k) of inputs:
  //   X[k]   = 1/2 [(Z[k] + conj Z[M-k]) - i w_k (Z[k] - conj Z[M-k])],  w_k = e^{-2 pi i k/Nx}
  //   X[M-k] = 1/2 conj[(Z[k] + conj Z[M-k]) + i w_k (Z[k] - conj Z[M-k])]      (w_{M-k} = -conj w_k)
  // with Z[M] = Z[0]; k = 0 yields X[0] and the Nyquist column X[M]
  T2 *dst = a.Hout + plane * (long long)(MX + 1) * a.ny + iy0;
#pragma unroll 4
  for (int e = tid; e < (MX / 2 + 1) * R; e += NTHREADS) {
    int k = e / R, r = e - k * R;
    const T2 *zr = s + r * PS;
    T2 zk = zr[pad(k)], zm = zr[pad(k == 0 ? 0 : MX - k)];
    T2 w = ldg2(a.tw + k * tws_n);
    T2 sum = cadd(zk, cconj(zm)), dif = csub(zk, cconj(zm));
    T2 pw = mul_i<+1>(cmul(w, dif));
    T2 x0, x1;
    x0.x = (T)0.5 * (sum.x - pw.x);
    x0.y = (T)0.5 * (sum.y - pw.y);
    x1.x = (T)0.5 * (sum.x + pw.x);
    x1.y = -(T)0.5 * (sum.y + pw.y);
    st_once(dst + (long long)k * a.ny + r, x0);
    if (2 * k != MX) st_once(dst + (long long)(MX - k) * a.ny + r, x1);
  }
}

// rows per CTA: 64-byte segments of the transposed layout (4 x double2 / 8 x float2) while two CTAs
// still fit in an SM's shared memory -- with a single resident CTA nothing overlaps its tile load
// (measured at nx = 4096: 2 rows of 35 KB each beat 4 rows); never below 32-byte segments
template <typename T, int MX>
struct RowCfg {
  typedef typename V2<T>::type T2;
  static constexpr int WANT = sizeof(T2) == 16 ? OX_KB_ROWS64 : 8;
  static constexpr int MINR = (int)(OX_KB_MINSEG / sizeof(T2)) > 0 ? (int)(OX_KB_MINSEG / sizeof(T2)) : 1;
  static constexpr size_t ROW = sizeof(T2) * padded_size(MX);
  static constexpr int FIT = (int)((SMEM_MAX / 2) / ROW);  // rows per CTA that leave room for a second CTA
  static constexpr int R = FIT >= WANT ? WANT : (FIT >= MINR ? (FIT >= 4 ? 4 : (FIT >= 2 ? 2 : 1)) : MINR);
};

}  // namespace oxk
#include "ox_row_tma.cuh"
#include "ox_row_w32.cuh"
namespace oxk {

template <typename T, int MX, int MODE>
int launch_row_mode(RowArgs<T> &a, long long nplanes) {
  typedef typename V2<T>::type T2;
  constexpr int R = RowCfg<T, MX>::R;
  {
    // Blackwell path: persistent CTAs fed by the TMA unit (ox_row_tma.cuh); ORPHX_KB=legacy keeps the kernel below
    bool launched = false;
    OX_TRY((launch_row_w32<T, MX, MODE>(a, nplanes, &launched)));   // nx = 2048: one warp per row, radix 32
    if (launched) return OX_OK;
    OX_TRY((launch_row_tma<T, MX, MODE>(a, nplanes, &launched)));
    if (launched) return OX_OK;
  }
  size_t smem = sizeof(T2) * R * padded_size(MX);
  OX_REQUIRE(smem <= SMEM_MAX, "fused row: %d rows of %d need %zu B of shared memory", R, MX, smem);
  OX_REQUIRE(a.ny % R == 0, "ny must be a multiple of %d", R);
  auto k = fused_row_kernel<T, MX, R, MODE>;
  OX_TRY(set_smem(k, smem));
  static const bool rows_fastest = [] { const char *e = getenv("ORPHX_KB_ORDER"); return e && !strcmp(e, "rows"); }();
  dim3 grid(a.ny / R, (unsigned)nplanes);
  if (!rows_fastest && (long long)(a.ny / R) * nplanes < (1LL << 31)) {
    a.nplanes_fast = (int)nplanes;
    grid = dim3((unsigned)((a.ny / R) * nplanes), 1);
  }
  const size_t wbytes = sizeof(T) * (size_t)a.ny * a.nx;
  if (a.window && !a.win_x && a.win_group_stride == 0 && nplanes > 1 && l2_persist_budget() >= wbytes) {
    // the window is shared by every plane of the launch but is evicted from L2 between its uses by the
    // planes streaming through (ncu: K_B re-read its 32 MB from DRAM for every map): keep it resident
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(R * (MX / 16));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = g_stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
    attr[0].val.accessPolicyWindow.base_ptr = const_cast<T *>(a.window);
    attr[0].val.accessPolicyWindow.num_bytes = wbytes;
    attr[0].val.accessPolicyWindow.hitRatio = 1.0f;
    attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    OX_CUDA(cudaLaunchKernelEx(&cfg, k, a));
  } else {
    k<<<grid, R * (MX / 16), smem, g_stream>>>(a);
  }
  OX_KERNEL_CHECK();
  return OX_OK;
}

// the including translation unit lists the modes it launches (each is a separate instantiation)
template <int... MODES>
struct RowModes {};

template <typename T, int MX>
int launch_row(RowArgs<T> &, long long, int mode, RowModes<>) {
  set_error("fused row pass: mode %d is not instantiated in this translation unit", mode);
  return OX_ERR_UNSUPPORTED;
}
template <typename T, int MX, int M0, int... REST>
int launch_row(RowArgs<T> &a, long long nplanes, int mode, RowModes<M0, REST...>) {
  if (mode == M0) return launch_row_mode<T, MX, M0>(a, nplanes);
  return launch_row<T, MX>(a, nplanes, mode, RowModes<REST...>());
}

// ---- generic column pass ------------------------------------------------------------------
// One CTA per (column ix of the transposed half plane, plane): ld(iy, m) supplies the first-stage
// inputs (global loads, optionally multiplied by per-pixel factors), FFT along y in direction DIR,
// st(iy, value, m) consumes the natural-order outputs.  The functors are built on the device from
// (args, ix, plane) so that one kernel template serves every elementwise fusion of the estimator.
template <typename T, int LY, int DIR, class Ops>
__global__ void __launch_bounds__(LY / 16, (LY / 16 <= 256 ? 512 / (LY / 16) : 1))
fused_col_kernel(Ops ops, const typename V2<T>::type *__restrict__ tw, int tw_len, int nplanes) {
  typedef typename V2<T>::type T2;
  typedef BlockFFT<T, LY> FFT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T2 *s = reinterpret_cast<T2 *>(smem_raw);
  const int u = threadIdx.x;
  typename FFT::Twiddles tws;
  tws.init(tw, tw_len / LY, u);
  // planes vary fastest across consecutive CTAs: CTAs that share a column of a batch-shared table (or of
  // the same input, for the three legs of the estimator) are resident together and reuse it through L2
  const int ix = blockIdx.x / nplanes, plane = blockIdx.x - ix * nplanes;
  typename Ops::Load ld = ops.load(ix, plane);
  typename Ops::Store st = ops.store(ix, plane);
  FFT::template run<DIR, false, false>(s, tws, u, 0, ld, st);
}

template <typename T, int LY, int DIR, class Ops>
int launch_col(const Ops &ops, const void *tw, int tw_len, int ncols, long long nplanes) {
  typedef typename V2<T>::type T2;
  size_t smem = sizeof(T2) * padded_size(LY);
  OX_REQUIRE(smem <= SMEM_MAX, "fused column pass: %d elements need %zu B of shared memory", LY, smem);
  auto k = fused_col_kernel<T, LY, DIR, Ops>;
  OX_TRY(set_smem(k, smem));
  OX_REQUIRE((long long)ncols * nplanes < (1LL << 31), "fused column pass: grid of %d x %lld CTAs is too large", ncols, nplanes);
  k<<<(unsigned)(ncols * nplanes), LY / 16, smem, g_stream>>>(ops, (const T2 *)tw, tw_len, (int)nplanes);
  OX_KERNEL_CHECK();
  return OX_OK;
}

// dispatch on the runtime column length
#define OX_COL_DISPATCH(T, DIR, ops, tw, tw_len, ncols, nplanes, ny, st)                        \
  switch (ny) {                                                                                  \
    case 512: st = oxk::launch_col<T, 512, DIR>(ops, tw, tw_len, ncols, nplanes); break;         \
    case 1024: st = oxk::launch_col<T, 1024, DIR>(ops, tw, tw_len, ncols, nplanes); break;       \
    case 2048: st = oxk::launch_col<T, 2048, DIR>(ops, tw, tw_len, ncols, nplanes); break;       \
    case 4096: st = oxk::launch_col<T, 4096, DIR>(ops, tw, tw_len, ncols, nplanes); break;       \
    case 8192: st = oxk::launch_col<T, 8192, DIR>(ops, tw, tw_len, ncols, nplanes); break;       \
    default: ox::set_error("fused column pass: unsupported ny=%d", ny); st = OX_ERR_UNSUPPORTED;  \
  }

template <typename T, class MODES>
int launch_row_any(RowArgs<T> &a, long long nplanes, MODES modes) {
  int mode = (a.Hin ? ROW_IN_H : 0) | (a.map_out ? ROW_OUT_MAP : 0) | (a.window ? ROW_WIN : 0) | (a.Hout ? ROW_OUT_H : 0) |
             (a.map_in2 ? ROW_WIN2 : 0);
  if (a.Hin && a.Hout) mode = ROW_IN_H | ROW_OUT_H;  // map_out / window are run-time options of the full pass
  switch (a.nx / 2) {
    case 128: return launch_row<T, 128>(a, nplanes, mode, modes);
    case 256: return launch_row<T, 256>(a, nplanes, mode, modes);
    case 512: return launch_row<T, 512>(a, nplanes, mode, modes);
    case 1024: return launch_row<T, 1024>(a, nplanes, mode, modes);
    case 2048: return launch_row<T, 2048>(a, nplanes, mode, modes);
    case 4096: return launch_row<T, 4096>(a, nplanes, mode, modes);
  }
  set_error("fused row pass: unsupported nx=%d", a.nx);
  return OX_ERR_UNSUPPORTED;
}

}  // namespace oxk
