// stats.bin2D on the device: digitize (bit-exact vs np.digitize(right=True)), integer
// counts, deterministic annular binning of full-plane data, and the fused half-plane
// conj(k1).k2 -> bandpower kernel used by FourierCalc.power2d + bin2D.bin.
//
// Determinism: no floating-point atomics anywhere.  Pixels are mapped to
// (block, warp, lane, iteration) by a fixed function of the map size only; lanes of
// a warp that hit the same slot are combined by a fixed shuffle tree
// (reduce_peers), each warp owns a private shared-memory slot array, warps are
// summed in index order, blocks are summed in index order by the finalize kernel.
#include <math.h>

#include "ox_common.cuh"

using namespace ox;

namespace {

constexpr int BIN_THREADS = 256;
constexpr int BIN_WARPS = BIN_THREADS / 32;
constexpr int BIN_CHUNK = 8192;  // pixels per block: fixes the summation order for a given map size

// ---- digitize ---------------------------------------------------------------------
// np.digitize(x, edges, right=True) for increasing edges == number of edges < x; NaN -> nedges
__device__ __forceinline__ int digitize_right(double x, const double *__restrict__ edges, int nedges) {
  if (x != x) return nedges;
  int lo = 0, hi = nedges;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (edges[mid] < x)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}

__global__ void digitize_kernel(const double *__restrict__ x, long long n, const double *__restrict__ edges, int nedges,
                                uint16_t *__restrict__ idx, unsigned long long *__restrict__ counts) {
  extern __shared__ unsigned int s_cnt[];
  for (int s = threadIdx.x; s <= nedges; s += blockDim.x) s_cnt[s] = 0;
  __syncthreads();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int d = digitize_right(x[i], edges, nedges);
    idx[i] = (uint16_t)d;
    atomicAdd(&s_cnt[d], 1u);
  }
  __syncthreads();
  for (int s = threadIdx.x; s <= nedges; s += blockDim.x)
    if (s_cnt[s]) atomicAdd(&counts[s], (unsigned long long)s_cnt[s]);
}

__global__ void digitize_geom_kernel(const double *__restrict__ ly, const double *__restrict__ lx, int ny, int nx,
                                     const double *__restrict__ edges, int nedges, uint16_t *__restrict__ idx,
                                     unsigned long long *__restrict__ counts) {
  extern __shared__ unsigned int s_cnt[];
  for (int s = threadIdx.x; s <= nedges; s += blockDim.x) s_cnt[s] = 0;
  __syncthreads();
  long long n = (long long)ny * nx;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int iy = (int)(i / nx), ix = (int)(i - (long long)iy * nx);
    double y = ly[iy], xx = lx[ix];
    double m = __dsqrt_rn(__dadd_rn(__dmul_rn(y, y), __dmul_rn(xx, xx)));
    int d = digitize_right(m, edges, nedges);
    idx[i] = (uint16_t)d;
    atomicAdd(&s_cnt[d], 1u);
  }
  __syncthreads();
  for (int s = threadIdx.x; s <= nedges; s += blockDim.x)
    if (s_cnt[s]) atomicAdd(&counts[s], (unsigned long long)s_cnt[s]);
}

// half-plane slot index with the Hermitian weight folded into bit 15; also checks that the
// mirrored pixel p' = (-iy, -ix) falls into the same slot (needed for the weight-2 shortcut)
__global__ void half_index_kernel(const uint16_t *__restrict__ idx, int ny, int nx, int nxh, uint16_t *__restrict__ idxh,
                                  int *__restrict__ mismatch) {
  long long n = (long long)ny * nxh;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int iy = (int)(i / nxh), ix = (int)(i - (long long)iy * nxh);
    int my = iy ? ny - iy : 0, mx = ix ? nx - ix : 0;
    uint16_t a = idx[(long long)iy * nx + ix], b = idx[(long long)my * nx + mx];
    if (a != b) atomicAdd(mismatch, 1);
    bool self_col = (ix == 0) || (2 * ix == nx);  // p' lies in the same half-plane column
    idxh[i] = (uint16_t)(a | (self_col ? 0 : 0x8000));
  }
}

__global__ void countf64_kernel(const unsigned long long *__restrict__ counts, int nslots, double *__restrict__ inv) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < nslots) inv[s] = (double)counts[s];
}

// ---- deterministic in-warp reduce-by-key ----------------------------------------------
// After the call the lowest lane of every group of equal keys holds the group's sums.
template <int NV>
__device__ __forceinline__ void reduce_peers(unsigned peers, double (&v)[NV]) {
  const int lane = threadIdx.x & 31;
  unsigned rel = __popc(peers & ((1u << lane) - 1u));
  peers &= (lane == 31) ? 0u : (0xfffffffeu << lane);  // peers above me
  while (__any_sync(0xffffffffu, peers)) {
    int next = __ffs(peers);  // 1-based lane of my next higher peer, 0 = none
    int src = next ? next - 1 : lane;
#pragma unroll
    for (int k = 0; k < NV; k++) {
      double t = __shfl_sync(0xffffffffu, v[k], src);
      if (next) v[k] += t;
    }
    unsigned done = rel & 1u;  // odd-ranked peers have just been absorbed
    peers &= __ballot_sync(0xffffffffu, !done);
    rel >>= 1;
  }
}

// ---- general full-plane binning --------------------------------------------------------
// sums[m][s] = sum data*w, cnts[m][s] = sum w over kept pixels of slot s (see orphx.h)
template <typename T, bool HAS_W, bool MASK_NAN>
__global__ void __launch_bounds__(BIN_THREADS)
bin_full_kernel(const T *__restrict__ data, const T *__restrict__ weights, const uint16_t *__restrict__ idx, long long n,
                int nslots, int nblk, double *__restrict__ partial /*[nmaps][nblk][2][nslots]*/) {
  extern __shared__ double s_bins[];  // [BIN_WARPS][2][nslots+1]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stride = nslots + 1;
  double *mysum = s_bins + (size_t)warp * 2 * stride;
  double *mycnt = mysum + stride;
  for (int s = lane; s < 2 * stride; s += 32) mysum[s] = 0.0;
  __syncwarp();
  const long long m = blockIdx.y;
  const T *d = data + m * n;
  long long begin = (long long)blockIdx.x * BIN_CHUNK;
  long long end = begin + BIN_CHUNK < n ? begin + BIN_CHUNK : n;
  for (long long base = begin; base < end; base += BIN_THREADS) {
    long long i = base + threadIdx.x;
    unsigned key = nslots;  // trash slot
    double v[2] = {0.0, 0.0};
    if (i < end) {
      double x = (double)d[i];
      bool keep = MASK_NAN ? (x == x) : true;
      if (keep) {
        key = idx[i];
        double w = HAS_W ? (double)weights[i] : 1.0;
        v[0] = HAS_W ? x * w : x;
        v[1] = w;
      }
    }
    unsigned peers = __match_any_sync(0xffffffffu, key);
    bool leader = (__ffs(peers) - 1) == lane;
    reduce_peers<2>(peers, v);
    if (leader) {
      mysum[key] += v[0];
      mycnt[key] += v[1];
    }
    __syncwarp();
  }
  __syncthreads();
  double *out = partial + ((size_t)m * nblk + blockIdx.x) * 2 * nslots;
  for (int s = threadIdx.x; s < 2 * nslots; s += BIN_THREADS) {
    int which = s / nslots, slot = s - which * nslots;
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < BIN_WARPS; w++) t += s_bins[(size_t)w * 2 * stride + which * stride + slot];
    out[s] = t;
  }
}

// out[m][v] = sum over blocks (index order) of partial[m][blk][v]
__global__ void reduce_blocks_kernel(const double *__restrict__ partial, int nblk, int nvals, double *__restrict__ out) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  long long m = blockIdx.y;
  if (v >= nvals) return;
  const double *p = partial + (size_t)m * nblk * nvals + v;
  double t = 0.0;
  for (int b = 0; b < nblk; b++) t += p[(size_t)b * nvals];
  out[(size_t)m * nvals + v] = t;
}

// ---- fused half-plane power + bin --------------------------------------------------------
// k1,k2: [nbatch][NC][ny*nxh] complex; for NC==3 with ROT the (Q,U)->(E,B) rotation of
// FourierCalc.iqu2teb (maps.py:1614-1615) is applied in registers.  Per pixel and spectrum
// (i<=j): Re(conj(k1_i) k2_j) * hermitian weight, accumulated per slot.
template <int NC>
struct NSpec {
  static constexpr int value = NC * (NC + 1) / 2;
};

template <typename T2, int NC, bool ROT, bool CROSS, bool SKIP_CROSS>
__global__ void __launch_bounds__(BIN_THREADS)
power_bin_half_kernel(const T2 *__restrict__ k1, const T2 *__restrict__ k2, const uint16_t *__restrict__ idxh,
                      const double *__restrict__ ly, const double *__restrict__ lx, int nxh, long long nh, int nslots,
                      int nblk, double rot_sgn, double *__restrict__ partial /*[nbatch][nblk][NS][nslots]*/) {
  constexpr int NS = SKIP_CROSS ? NC : NSpec<NC>::value;
  extern __shared__ double s_bins[];  // [BIN_WARPS][NS][nslots+1]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stride = nslots + 1;
  double *mine = s_bins + (size_t)warp * NS * stride;
  for (int s = lane; s < NS * stride; s += 32) mine[s] = 0.0;
  __syncwarp();
  const long long m = blockIdx.y;
  const int ny = (int)(nh / nxh);
  const T2 *a = k1 + m * NC * nh;
  const T2 *b = CROSS ? k2 + m * NC * nh : a;
  long long begin = (long long)blockIdx.x * BIN_CHUNK;
  long long end = begin + BIN_CHUNK < nh ? begin + BIN_CHUNK : nh;
  for (long long base = begin; base < end; base += BIN_THREADS) {
    long long i = base + threadIdx.x;
    unsigned key = nslots;
    double v[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) v[s] = 0.0;
    unsigned raw = 0;
    if (i < end) {
      raw = idxh[i];
      key = raw & 0x7fffu;
      // slots 0 and nslots-1 are dropped by bin2D's [1:-1] (stats.py:796-797)
      if (key == 0u || key == (unsigned)(nslots - 1)) key = nslots;
    }
    if (__all_sync(0xffffffffu, key == (unsigned)nslots)) continue;
    if (key != (unsigned)nslots) {
      double w = (raw & 0x8000u) ? 2.0 : 1.0;
      double ar[NC], ai[NC], br[NC], bi[NC];
#pragma unroll
      for (int c = 0; c < NC; c++) {
        T2 z = a[(long long)c * nh + i];
        ar[c] = (double)z.x;
        ai[c] = (double)z.y;
        if (CROSS) {
          T2 y = b[(long long)c * nh + i];
          br[c] = (double)y.x;
          bi[c] = (double)y.y;
        }
      }
      double c = 1.0, s = 0.0;
      bool nyq_pair = false;
      if (ROT && NC == 3) {
        int iy = (int)(i / nxh), ix = (int)(i - (long long)iy * nxh);
        double y = ly[iy], x = lx[ix];
        double l2 = y * y + x * x;
        if (l2 > 0.0) {
          double inv = 1.0 / l2;
          c = (y * y - x * x) * inv;            // cos 2*atan2(-lx,ly)
          s = rot_sgn * (-2.0 * x * y) * inv;   // sgn * sin 2*atan2(-lx,ly)
        }
        // Nyquist row of an interior column: the mirrored pixel p' = (iy, nx - ix) has the same ly and the
        // opposite lx, so its rotation has the opposite sine and its E/B power differs from the one at p:
        // the pair is accumulated as one term with +s and one with -s instead of twice the term at p
        nyq_pair = (2 * iy == ny) && w == 2.0;
      }
      if (!CROSS) {
#pragma unroll
        for (int c2 = 0; c2 < NC; c2++) { br[c2] = ar[c2]; bi[c2] = ai[c2]; }
      }
      const int nterm = nyq_pair ? 2 : 1;
      const double wt = nyq_pair ? 1.0 : w;
      for (int term = 0; term < nterm; term++) {
        double pr[NC], pi[NC], qr[NC], qi[NC];
#pragma unroll
        for (int c2 = 0; c2 < NC; c2++) { pr[c2] = ar[c2]; pi[c2] = ai[c2]; qr[c2] = br[c2]; qi[c2] = bi[c2]; }
        if (ROT && NC == 3) {
          const double st = term ? -s : s;
          pr[1] = c * ar[1] - st * ar[2]; pi[1] = c * ai[1] - st * ai[2];
          pr[2] = st * ar[1] + c * ar[2]; pi[2] = st * ai[1] + c * ai[2];
          qr[1] = c * br[1] - st * br[2]; qi[1] = c * bi[1] - st * bi[2];
          qr[2] = st * br[1] + c * br[2]; qi[2] = st * bi[1] + c * bi[2];
        }
        if (SKIP_CROSS) {
#pragma unroll
          for (int c2 = 0; c2 < NC; c2++) v[c2] += (pr[c2] * qr[c2] + pi[c2] * qi[c2]) * wt;
        } else {
          int sidx = 0;
#pragma unroll
          for (int p = 0; p < NC; p++)
#pragma unroll
            for (int q = p; q < NC; q++) v[sidx++] += (pr[p] * qr[q] + pi[p] * qi[q]) * wt;
        }
      }
    }
    unsigned peers = __match_any_sync(0xffffffffu, key);
    bool leader = (__ffs(peers) - 1) == lane;
    reduce_peers<NS>(peers, v);
    if (leader) {
#pragma unroll
      for (int s = 0; s < NS; s++) mine[s * stride + key] += v[s];
    }
    __syncwarp();
  }
  __syncthreads();
  double *out = partial + ((size_t)m * nblk + blockIdx.x) * NS * nslots;
  for (int t = threadIdx.x; t < NS * nslots; t += BIN_THREADS) {
    int sp = t / nslots, slot = t - sp * nslots;
    double acc = 0.0;
#pragma unroll
    for (int w = 0; w < BIN_WARPS; w++) acc += s_bins[(size_t)w * NS * stride + sp * stride + slot];
    out[t] = acc;
  }
}

// bandpowers[m][sp][bin] = normfact * (sum_blk partial[m][blk][sp][bin+1]) / count[bin+1]
__global__ void bandpower_finalize_kernel(const double *__restrict__ partial, int nblk, int ns, int nslots,
                                          const double *__restrict__ count, double normfact, double *__restrict__ bp) {
  const int nbins = nslots - 2;
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  long long m = blockIdx.y;
  if (t >= ns * nbins) return;
  int sp = t / nbins, bin = t - sp * nbins;
  const double *p = partial + (size_t)m * nblk * ns * nslots + (size_t)sp * nslots + bin + 1;
  double acc = 0.0;
  for (int b = 0; b < nblk; b++) acc += p[(size_t)b * ns * nslots];
  bp[(size_t)m * ns * nbins + t] = (acc * normfact) / count[bin + 1];
}

int grid_1d(long long n, int block) {
  long long want = (n + block - 1) / block;
  long long cap = (long long)ox::sm_count() * 8;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

int finish_binner(ox_binner *b) {
  int nslots = b->nslots;
  b->h_counts.resize(nslots);
  OX_CUDA(cudaMemcpyAsync(b->h_counts.data(), b->counts.p, sizeof(long long) * nslots, cudaMemcpyDeviceToHost, g_stream));
  OX_TRY(b->countf.ensure(sizeof(double) * nslots));
  countf64_kernel<<<(nslots + 127) / 128, 128, 0, g_stream>>>(b->counts.as<unsigned long long>(), nslots,
                                                             b->countf.as<double>());
  OX_KERNEL_CHECK();
  OX_CUDA(cudaStreamSynchronize(g_stream));
  return OX_OK;
}

int check_edges(const double *edges, int nedges) {
  OX_REQUIRE(edges && nedges >= 2, "bin2D needs at least two bin edges");
  OX_REQUIRE(nedges + 1 <= 0x7fff, "bin2D: at most 32766 edges supported (got %d)", nedges);
  for (int i = 1; i < nedges; i++)
    OX_REQUIRE(edges[i] > edges[i - 1], "bin2D: bin_edges must be strictly increasing (edge %d)", i);
  return OX_OK;
}

template <typename F>
int set_smem(F kernel, size_t bytes) {
  if (bytes > 48 * 1024) OX_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return OX_OK;
}

template <typename T>
int launch_bin_full(const T *data, const T *w, const uint16_t *idx, long long n, int nslots, int nblk, long long nmaps,
                    int flags, double *partial) {
  size_t smem = sizeof(double) * BIN_WARPS * 2 * (nslots + 1);
  OX_REQUIRE(smem <= 200 * 1024, "bin2D: %d slots need %zu bytes of shared memory (max 200 KiB)", nslots, smem);
  dim3 grid(nblk, (unsigned)nmaps);
  bool mask = flags & OX_FLAG_MASK_NAN;
#define OX_LAUNCH(HW, MN)                                                                          \
  do {                                                                                             \
    OX_TRY(set_smem(bin_full_kernel<T, HW, MN>, smem));                                            \
    bin_full_kernel<T, HW, MN><<<grid, BIN_THREADS, smem, g_stream>>>(data, w, idx, n, nslots, nblk, partial); \
  } while (0)
  if (w) {
    OX_LAUNCH(true, false);  // the reference ignores mask_nan when weights are given (stats.py:802-804)
  } else if (mask) {
    OX_LAUNCH(false, true);
  } else {
    OX_LAUNCH(false, false);
  }
#undef OX_LAUNCH
  OX_KERNEL_CHECK();
  return OX_OK;
}

template <typename T2, int NC>
int launch_power_bin(const void *k1, const void *k2, ox_geometry *g, ox_binner *b, int nbatch, int flags, int nblk,
                     double *partial) {
  const bool rot = (flags & OX_FLAG_ROT) && NC == 3;
  const bool cross = k2 != nullptr && k2 != k1;
  const bool skip = (flags & OX_FLAG_SKIP_CROSS) && NC > 1;
  const int ns = skip ? NC : NC * (NC + 1) / 2;
  size_t smem = sizeof(double) * BIN_WARPS * ns * (b->nslots + 1);
  OX_REQUIRE(smem <= 200 * 1024, "power_bin: %d slots x %d spectra need %zu bytes of shared memory (max 200 KiB)",
             b->nslots, ns, smem);
  dim3 grid(nblk, nbatch);
  long long nh = (long long)g->ny * g->nxh;
  double sgn = (flags & OX_FLAG_IAU) ? 1.0 : -1.0;
#define OX_LAUNCH(R, C, S)                                                                                     \
  do {                                                                                                         \
    OX_TRY(set_smem(power_bin_half_kernel<T2, NC, R, C, S>, smem));                                            \
    power_bin_half_kernel<T2, NC, R, C, S><<<grid, BIN_THREADS, smem, g_stream>>>(                             \
        (const T2 *)k1, (const T2 *)k2, b->idxh.as<uint16_t>(), g->ly.as<double>(), g->lx.as<double>(), g->nxh, nh, \
        b->nslots, nblk, sgn, partial);                                                                        \
  } while (0)
  if (NC == 1) {
    if (cross) OX_LAUNCH(false, true, false); else OX_LAUNCH(false, false, false);
  } else if (rot) {
    if (cross) { if (skip) OX_LAUNCH(true, true, true); else OX_LAUNCH(true, true, false); }
    else { if (skip) OX_LAUNCH(true, false, true); else OX_LAUNCH(true, false, false); }
  } else {
    if (cross) { if (skip) OX_LAUNCH(false, true, true); else OX_LAUNCH(false, true, false); }
    else { if (skip) OX_LAUNCH(false, false, true); else OX_LAUNCH(false, false, false); }
  }
#undef OX_LAUNCH
  OX_KERNEL_CHECK();
  return OX_OK;
}

}  // namespace

namespace ox {

int power_bin_half(ox_geometry *g, ox_binner *b, int dtype, int ncomp, const void *kh1, const void *kh2, int nbatch,
                   int flags, double normfact, DevBuf &partial, double *bp_dev) {
  OX_REQUIRE(b->has_half, "this bin2D was not built from a Hermitian-symmetric geometry; the fused half-plane path needs ox_binner_create_geom");
  OX_REQUIRE(b->ny == g->ny && b->nx == g->nx, "binner/geometry shape mismatch");
  OX_REQUIRE(ncomp >= 1 && ncomp <= 3, "power_bin: ncomp must be 1..3 (got %d)", ncomp);
  long long nh = (long long)g->ny * g->nxh;
  int nblk = (int)((nh + BIN_CHUNK - 1) / BIN_CHUNK);
  const bool skip = (flags & OX_FLAG_SKIP_CROSS) && ncomp > 1;
  int ns = skip ? ncomp : ncomp * (ncomp + 1) / 2;
  OX_TRY(partial.ensure(sizeof(double) * (size_t)nbatch * nblk * ns * b->nslots));
  int st = OX_OK;
  if (dtype == OX_F64) {
    if (ncomp == 1) st = launch_power_bin<double2, 1>(kh1, kh2, g, b, nbatch, flags, nblk, partial.as<double>());
    else if (ncomp == 2) st = launch_power_bin<double2, 2>(kh1, kh2, g, b, nbatch, flags, nblk, partial.as<double>());
    else st = launch_power_bin<double2, 3>(kh1, kh2, g, b, nbatch, flags, nblk, partial.as<double>());
  } else {
    if (ncomp == 1) st = launch_power_bin<float2, 1>(kh1, kh2, g, b, nbatch, flags, nblk, partial.as<double>());
    else if (ncomp == 2) st = launch_power_bin<float2, 2>(kh1, kh2, g, b, nbatch, flags, nblk, partial.as<double>());
    else st = launch_power_bin<float2, 3>(kh1, kh2, g, b, nbatch, flags, nblk, partial.as<double>());
  }
  OX_TRY(st);
  int nbins = b->nslots - 2;
  dim3 grid((ns * nbins + 127) / 128, nbatch);
  double nf = (flags & OX_FLAG_PIXEL_UNITS) ? 1.0 : normfact;
  bandpower_finalize_kernel<<<grid, 128, 0, g_stream>>>(partial.as<double>(), nblk, ns, b->nslots,
                                                        b->countf.as<double>(), nf, bp_dev);
  OX_KERNEL_CHECK();
  return OX_OK;
}

}  // namespace ox

extern "C" {

int ox_binner_create(const double *modrmap, int where, long long n, const double *edges, int nedges, ox_binner **out) {
  OX_REQUIRE(modrmap && out && n > 0, "ox_binner_create: bad arguments");
  OX_TRY(check_edges(edges, nedges));
  ox_binner *b = new ox_binner;
  b->n = n;
  b->nedges = nedges;
  b->nslots = nedges + 1;
  int st = OX_OK;
  const void *dmod = nullptr;
  ox::DevBuf tmp;
  auto fail = [&](int s) { delete b; return s; };
  if ((st = b->edges.ensure(sizeof(double) * nedges)) != OX_OK) return fail(st);
  if ((st = b->idx.ensure(sizeof(uint16_t) * n)) != OX_OK) return fail(st);
  if ((st = b->counts.ensure(sizeof(long long) * b->nslots)) != OX_OK) return fail(st);
  if ((st = stage_in(modrmap, where, sizeof(double) * n, tmp, &dmod)) != OX_OK) return fail(st);
  cudaMemcpyAsync(b->edges.p, edges, sizeof(double) * nedges, cudaMemcpyHostToDevice, g_stream);
  cudaMemsetAsync(b->counts.p, 0, sizeof(long long) * b->nslots, g_stream);
  digitize_kernel<<<grid_1d(n, 256), 256, sizeof(unsigned) * b->nslots, g_stream>>>(
      (const double *)dmod, n, b->edges.as<double>(), nedges, b->idx.as<uint16_t>(), b->counts.as<unsigned long long>());
  g_launches++;
  if (cudaGetLastError() != cudaSuccess) { set_error("digitize kernel launch failed"); return fail(OX_ERR_CUDA); }
  if ((st = finish_binner(b)) != OX_OK) return fail(st);
  *out = b;
  return OX_OK;
}

int ox_binner_create_geom(ox_geometry *g, const double *edges, int nedges, ox_binner **out) {
  OX_REQUIRE(g && out, "ox_binner_create_geom: bad arguments");
  OX_TRY(check_edges(edges, nedges));
  ox_binner *b = new ox_binner;
  long long n = (long long)g->ny * g->nx;
  b->n = n;
  b->nedges = nedges;
  b->nslots = nedges + 1;
  b->ny = g->ny; b->nx = g->nx; b->nxh = g->nxh;
  int st = OX_OK;
  auto fail = [&](int s) { delete b; return s; };
  if ((st = b->edges.ensure(sizeof(double) * nedges)) != OX_OK) return fail(st);
  if ((st = b->idx.ensure(sizeof(uint16_t) * n)) != OX_OK) return fail(st);
  if ((st = b->counts.ensure(sizeof(long long) * b->nslots)) != OX_OK) return fail(st);
  long long nh = (long long)g->ny * g->nxh;
  if ((st = b->idxh.ensure(sizeof(uint16_t) * nh)) != OX_OK) return fail(st);
  ox::DevBuf mism;
  if ((st = mism.ensure(sizeof(int))) != OX_OK) return fail(st);
  cudaMemcpyAsync(b->edges.p, edges, sizeof(double) * nedges, cudaMemcpyHostToDevice, g_stream);
  cudaMemsetAsync(b->counts.p, 0, sizeof(long long) * b->nslots, g_stream);
  cudaMemsetAsync(mism.p, 0, sizeof(int), g_stream);
  digitize_geom_kernel<<<grid_1d(n, 256), 256, sizeof(unsigned) * b->nslots, g_stream>>>(
      g->ly.as<double>(), g->lx.as<double>(), g->ny, g->nx, b->edges.as<double>(), nedges, b->idx.as<uint16_t>(),
      b->counts.as<unsigned long long>());
  g_launches++;
  half_index_kernel<<<grid_1d(nh, 256), 256, 0, g_stream>>>(b->idx.as<uint16_t>(), g->ny, g->nx, g->nxh,
                                                           b->idxh.as<uint16_t>(), mism.as<int>());
  g_launches++;
  int h_mism = 0;
  if (cudaMemcpyAsync(&h_mism, mism.p, sizeof(int), cudaMemcpyDeviceToHost, g_stream) != cudaSuccess ||
      cudaStreamSynchronize(g_stream) != cudaSuccess) {
    set_error("ox_binner_create_geom: %s", cudaGetErrorString(cudaGetLastError()));
    return fail(OX_ERR_CUDA);
  }
  b->has_half = (h_mism == 0);
  if ((st = finish_binner(b)) != OX_OK) return fail(st);
  *out = b;
  return OX_OK;
}

int ox_binner_destroy(ox_binner *b) {
  delete b;
  return OX_OK;
}

__global__ void widen_idx_kernel(const uint16_t *__restrict__ idx, long long n, long long *__restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = idx[i];
}

int ox_binner_digitized(ox_binner *b, long long *out_host) {
  OX_REQUIRE(b && out_host, "null pointer");
  ox::DevBuf tmp;
  OX_TRY(tmp.ensure(sizeof(long long) * b->n));
  widen_idx_kernel<<<grid_1d(b->n, 256), 256, 0, g_stream>>>(b->idx.as<uint16_t>(), b->n, tmp.as<long long>());
  OX_KERNEL_CHECK();
  return stage_out(out_host, OX_HOST, tmp.p, sizeof(long long) * b->n);
}

int ox_binner_counts(ox_binner *b, long long *out_host) {
  OX_REQUIRE(b && out_host, "null pointer");
  memcpy(out_host, b->h_counts.data(), sizeof(long long) * b->nslots);
  return OX_OK;
}

int ox_binner_bin(ox_binner *b, const void *data, int dtype, int where, long long nmaps, const void *weights, int flags,
                  double *sums, double *counts, int out_where) {
  OX_REQUIRE(b && data && sums && counts && nmaps > 0, "ox_binner_bin: bad arguments");
  OX_REQUIRE(dtype == OX_F64 || dtype == OX_F32, "bad dtype");
  size_t es = elem_size(dtype);
  const void *dd = nullptr, *dw = nullptr;
  OX_TRY(stage_in(data, where, es * b->n * nmaps, b->stage, &dd));
  ox::DevBuf wtmp;
  if (weights) OX_TRY(stage_in(weights, where, es * b->n, wtmp, &dw));
  int nblk = (int)((b->n + BIN_CHUNK - 1) / BIN_CHUNK);
  int nvals = 2 * b->nslots;
  OX_TRY(b->partial.ensure(sizeof(double) * (size_t)nmaps * nblk * nvals));
  OX_TRY(b->scratch.ensure(sizeof(double) * (size_t)nmaps * nvals));
  if (dtype == OX_F64)
    OX_TRY(launch_bin_full<double>((const double *)dd, (const double *)dw, b->idx.as<uint16_t>(), b->n, b->nslots, nblk,
                                   nmaps, flags, b->partial.as<double>()));
  else
    OX_TRY(launch_bin_full<float>((const float *)dd, (const float *)dw, b->idx.as<uint16_t>(), b->n, b->nslots, nblk,
                                  nmaps, flags, b->partial.as<double>()));
  dim3 grid((nvals + 127) / 128, (unsigned)nmaps);
  reduce_blocks_kernel<<<grid, 128, 0, g_stream>>>(b->partial.as<double>(), nblk, nvals, b->scratch.as<double>());
  OX_KERNEL_CHECK();
  // scratch is [m][2][nslots]; the ABI returns sums[m][nslots] and counts[m][nslots]
  for (long long m = 0; m < nmaps; m++) {
    const double *src = b->scratch.as<double>() + m * nvals;
    if (out_where == OX_DEVICE) {
      OX_CUDA(cudaMemcpyAsync(sums + m * b->nslots, src, sizeof(double) * b->nslots, cudaMemcpyDeviceToDevice, g_stream));
      OX_CUDA(cudaMemcpyAsync(counts + m * b->nslots, src + b->nslots, sizeof(double) * b->nslots,
                              cudaMemcpyDeviceToDevice, g_stream));
    } else {
      OX_CUDA(cudaMemcpyAsync(sums + m * b->nslots, src, sizeof(double) * b->nslots, cudaMemcpyDeviceToHost, g_stream));
      OX_CUDA(cudaMemcpyAsync(counts + m * b->nslots, src + b->nslots, sizeof(double) * b->nslots,
                              cudaMemcpyDeviceToHost, g_stream));
    }
  }
  if (out_where == OX_HOST || weights) OX_CUDA(cudaStreamSynchronize(g_stream));
  return OX_OK;
}

}  // extern "C"
