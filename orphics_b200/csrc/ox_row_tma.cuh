// Persistent TMA version of the row pass (K_B and the c2r row passes of the estimator), 16-byte elements.
//
// The one-tile-per-CTA row kernel (fused_row_kernel) loads its tile with per-element LDGSTS and then waits for
// it; with two CTAs per SM nothing but the other CTA's arithmetic hides that wait, and the register file is full
// (128 x 256 x 2), so more CTAs are not an option.  Here ONE persistent CTA per SM runs two independent groups of
// R x MX/16 threads (the same 16 warps) over THREE shared-memory slots:
//
//   * a tile (R rows x (MX+1) columns of the transposed half plane = R x 16 B segments, one per column) is fetched
//     by the TMA unit -- cp.async.bulk.tensor boxes of 256 columns x R rows with the 64 B / 32 B swizzle, plus one
//     cp.async.bulk for the Nyquist column -- into whichever slot is free, completing on that slot's mbarrier;
//   * a group that finishes a tile hands its slot to the TMA for the next tile of the CTA and takes over the slot
//     whose tile is already in flight or landed: the load of tile i+2 overlaps the transforms of tiles i and i+1,
//     and no thread ever issues a load instruction for the tile;
//   * the first FFT stage reads the swizzled tile directly (conflict-free: 8 consecutive columns of one row fall
//     into 8 different 16-byte bank groups), a group-wide named barrier separates those reads from the stage's
//     writes, which rebuild the slot as the padded per-row work area of the FFT engine (ox_fft.cuh).
//
// Tiles are dealt to the CTAs round-robin in plane-fastest order, so the CTAs resident at any time work on the
// same rows of different planes and a row of the batch-shared window is read from DRAM once per launch.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace oxk {

// ---- PTX wrappers ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// 3-D tiled TMA load: box at coordinates (c0, c1, c2) of the tensor map -> shared memory, completes on bar
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, unsigned long long *bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// plain bulk copy global -> shared (bytes a multiple of 16), completes on bar
__device__ __forceinline__ void bulk_load(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// position (in 16-byte elements) of element (column k, row r) of a TMA tile whose R-row segments were written
// with the R*16-byte swizzle: 16-byte chunk index ^= address bits 7.. (CU_TENSOR_MAP_SWIZZLE_64B / _32B)
template <int R>
__device__ __forceinline__ int tile_pos(int k, int r) {
  if (R == 4) return k * 4 + (r ^ ((k >> 1) & 3));
  if (R == 2) return k * 2 + (r ^ ((k >> 2) & 1));
  return k * R + r;
}

// first-stage input of the c2r transform read straight from the swizzled TMA tile (see PackLoad)
template <typename T2, int MX, int R>
struct PackLoadTile {
  const T2 *tile;
  int r;
  T2 wu;
  __device__ __forceinline__ T2 operator()(int k, int m) const {
    constexpr double C32[16] = {1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708,
                                0.70710678118654752440, 0.55557023301960222474, 0.38268343236508977173, 0.19509032201612826785,
                                0.0, -0.19509032201612826785, -0.38268343236508977173, -0.55557023301960222474,
                                -0.70710678118654752440, -0.83146961230254523708, -0.92387953251128675613, -0.98078528040323044913};
    constexpr double S32[16] = {0.0, 0.19509032201612826785, 0.38268343236508977173, 0.55557023301960222474,
                                0.70710678118654752440, 0.83146961230254523708, 0.92387953251128675613, 0.98078528040323044913,
                                1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708,
                                0.70710678118654752440, 0.55557023301960222474, 0.38268343236508977173, 0.19509032201612826785};
    T2 xk = tile[tile_pos<R>(k, r)], xm = tile[tile_pos<R>(MX - k, r)];
    typedef decltype(xk.x) T;
    T2 w;
    w.x = wu.x * (T)C32[m] - wu.y * (T)S32[m];
    w.y = wu.x * (T)S32[m] + wu.y * (T)C32[m];
    T2 sum = cadd(xk, cconj(xm)), dif = csub(xk, cconj(xm));
    return cadd(sum, mul_i<+1>(cmul(w, dif)));
  }
};

// last-stage output of the c2r transform in the persistent kernel (see WindowKeep): store the map, apply the window
// and keep the element in registers.  The x profile of a separable window lives in shared memory (the carve-out
// leaves ~29 KB of L1: the 16 KB profile and the 32 KB twiddle table evicted each other and 3 of 4 table loads
// went to L2 -- ncu round 2: 118 M sectors of table loads per launch, 26% L1 hits, long-scoreboard stalls).
template <typename T, bool OUT_H, int NT>
struct RowTmaStore {
  typedef typename V2<T>::type T2;
  T2 *keep;               // registers [16]
  T2 *map_row;            // global or null
  const T2 *win_row;      // global or null (general window)
  const double2 *swinx;   // shared memory or null: x profile of a separable window as pairs, times wy
  double wy;
  unsigned flat;          // bit m: the profile is exactly 1 for every thread's m-th output (no load needed)
  __device__ __forceinline__ void operator()(int n, T2 z, int m) const {
    if (!OUT_H || map_row != nullptr) st_once(map_row + n, z);
    if (OUT_H) {
      if (swinx != nullptr) {
        if (flat & (1u << m)) {
          z.x *= (T)wy;
          z.y *= (T)wy;
        } else {
          const double2 p = swinx[n];
          z.x *= (T)(p.x * wy);   // the window value itself is formed in float64 and rounded once, as numpy forms it
          z.y *= (T)(p.y * wy);
        }
      } else if (win_row != nullptr) {
        T2 w1 = ldg2(win_row + n);
        z.x *= w1.x;
        z.y *= w1.y;
      }
      keep[m] = z;
    }
  }
};

// The i-th tile of CTA blockIdx.x -> (row tile, plane); false past the CTA's last tile.
//   tile_order G >= 2 (default: G = 8 / R, the row tiles that share a 128-byte line): BLOCKS of G adjacent row tiles of one plane
//     go to one CTA as consecutive tiles -- for 4-row tiles a pair (2q, 2q+1) that its two groups work on at the same time.  The 64-byte segments of two adjacent 4-row tiles are the two halves of one
//     128-byte line; the TMA unit reads whole lines from L2 for a 64-byte box row (ncu: 125 M sectors for 67 M of data), so
//     with the halves requested microseconds apart by one SM the second request hits L2 and the DRAM traffic is the data,
//     once (it was 1.01-1.13 x the data depending on timing when the neighbour tile belonged to another CTA).  Pairs are dealt
//     to the CTAs planes-fastest, so a 2-D window shared by the planes is still read from DRAM once per launch.
//   tile_order 0: single tiles, planes fastest;  1: single tiles, rows fastest inside a plane.
template <typename T>
__device__ __forceinline__ bool tile_coords(const RowArgs<T> &a, int i, int nplanes, int ntiles, int &rowtile, int &plane) {
  if (a.tile_order >= 2) {   // blocks of tile_order adjacent row tiles (2 for 4-row tiles, 4 for 2-row tiles: one 128-byte line)
    const int G = a.tile_order;
    const int blk = blockIdx.x + (i / G) * gridDim.x;
    if (G * blk >= ntiles) return false;
    const int rp = blk / nplanes;
    plane = blk - rp * nplanes;
    rowtile = G * rp + (i % G);
    return true;
  }
  const int t = blockIdx.x + i * gridDim.x;   // (ntiles + gridDim.x < 2^31, checked by the launcher)
  if (t >= ntiles) return false;
  if (a.tile_order == 1) {
    const int per = ntiles / nplanes;
    plane = t / per;
    rowtile = t - plane * per;
  } else {
    rowtile = t / nplanes;
    plane = t - rowtile * nplanes;
  }
  return true;
}

// NG groups of R x MX/16 threads over NS = NG + spare shared-memory slots.  Default: 2 groups of 4-row tiles over 3 slots
// (2-row tiles where 4 rows do not fit: the estimator's c2r pass at nx = 4096).
template <int MX, int R, int NG = 2, int NS = 3>
struct RowTmaCfg {
  static constexpr int NT = MX / 16, GROUP = R * NT, NTHREADS = NG * GROUP, NGROUPS = NG, NSLOTS = NS;
  // per-row work area of the FFT engine: pad(MX) + the Nyquist element, rounded up to 2 (mod 8) elements so that
  // equal indices of neighbouring rows fall into different 16-byte bank groups (tighter than padded_size())
  static constexpr int PS0 = MX + MX / 16 + 1, PS = PS0 + ((2 - PS0 % 8) + 8) % 8;
  static constexpr size_t WORK = 16 * (size_t)R * PS;                      // padded work area
  static constexpr size_t TILE = 16 * (size_t)R * (MX + 1);               // raw tile
  static constexpr size_t SALIGN = R == 4 ? 512 : 256;                    // (the swizzle pattern repeats every 512 / 256 B)
  static constexpr size_t SLOT = (((WORK > TILE ? WORK : TILE) + SALIGN - 1) / SALIGN) * SALIGN;
  static constexpr size_t HDR = 128;                                      // mbarriers [0,64), flat-block mask [64,68), fill counters [68,100)
  static constexpr size_t WINX = 8 * (size_t)(2 * MX);                    // x profile of a separable window
  static constexpr size_t UTW = 16 * (size_t)(MX / 4 + 1);                // exp(-2 pi i k / Nx), k <= Nx/8
  static constexpr size_t TW16 = 16 * 16;                                 // compact copy of the 16 second-stage twiddles
  static constexpr size_t TABS = NS * SLOT;                               // offset of the header behind the slots
  static constexpr size_t SMEM_C2R = NS * SLOT + HDR + UTW + TW16 + 1024; // (+ alignment slack)
  static constexpr size_t SMEM_FULL = NS * SLOT + HDR + UTW + TW16 + WINX + 1024;
  static constexpr int BOXW = 256;                                        // columns per TMA box
  static_assert(MX % BOXW == 0, "MX must be a multiple of the TMA box width");
  static_assert(NS > NG && NS <= 8 && 1 + NG + NG * R <= 16, "slots / named barriers");
};

// MODE: ROW_IN_H | ROW_OUT_H (full pass; map_out / window are run-time options) or ROW_IN_H | ROW_OUT_MAP (c2r only)
//
// Schedule (static, no bookkeeping in memory): the CTA's i-th tile is t = blockIdx.x + i * gridDim.x; it lives in slot
// i % 3 and is transformed by group i % 2.  The group that finishes tile i refills its slot with tile i + 3 (which
// the OTHER group will transform after its tile i + 1) and moves on to tile i + 2, whose load was started a tile
// and a half earlier.  The k-th fill of a slot completes phase k of the slot's mbarrier.
template <typename T, int MX, int R, int MODE, int NG = 2, int NS = 3>
__global__ void __launch_bounds__(RowTmaCfg<MX, R, NG, NS>::NTHREADS, 1)
fused_row_tma_kernel(RowArgs<T> a, const __grid_constant__ CUtensorMap tmap, int nplanes, int ntiles) {
  static_assert(sizeof(T) == 8, "the TMA row pass is written for 16-byte elements");
  constexpr bool OUT_H = MODE & ROW_OUT_H;
  typedef typename V2<T>::type T2;
  typedef BlockFFT<T, MX> FFT;
  typedef RowTmaCfg<MX, R, NG, NS> Cfg;
  constexpr int NT = Cfg::NT, GROUP = Cfg::GROUP, PS = Cfg::PS, NX = 2 * MX;
  extern __shared__ unsigned char smem_dyn[];
  // one thread: fetch the i-th tile of this CTA into slot i % 3 (nothing to do past the CTA's last tile)
  auto issue = [&a, &tmap, nplanes, ntiles](unsigned char *base, int i) {
    int rowtile, plane;
    if (!tile_coords(a, i, nplanes, ntiles, rowtile, plane)) return;
    const int slot = i % NS;
    unsigned char *dst = base + slot * Cfg::SLOT;
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(base + Cfg::TABS) + slot;
    fence_proxy_async();   // the slot was last written through the generic proxy (the FFT work area)
    mbar_expect_tx(bar, (unsigned)Cfg::TILE);
#pragma unroll
    for (int j = 0; j < MX / Cfg::BOXW; j++)
      tma_load_3d(dst + (size_t)j * Cfg::BOXW * R * 16, &tmap, 2 * rowtile * R, j * Cfg::BOXW, plane, bar);
    const T2 *nyq = a.Hin + ((long long)plane * (MX + 1) + MX) * a.ny + rowtile * R;
    bulk_load(dst + (size_t)MX * R * 16, nyq, R * 16, bar);
    __threadfence_block();
    reinterpret_cast<volatile int *>(base + Cfg::TABS + 68)[slot] = i / NS + 1;   // fills issued into this slot so far
  };

  {
    // slots aligned to 1024 B; [3 slots][3 mbarriers, flat-block mask, 3 fill counters (64 B)][twiddles, split by parity][16 second-stage
    // twiddles][x profile of the window]
    unsigned char *base = smem_dyn + ((1024 - (smem_u32(smem_dyn) & 1023)) & 1023);
    if (threadIdx.x == 0) {
      unsigned long long *bars = reinterpret_cast<unsigned long long *>(base + Cfg::TABS);
      for (int b = 0; b < NS; b++) {
        mbar_init(bars + b, 1);
        reinterpret_cast<volatile int *>(base + Cfg::TABS + 68)[b] = 0;
      }
      fence_mbar_init();
      for (int b = 0; b < NS; b++) issue(base, b);
    }
    T2 *utw = reinterpret_cast<T2 *>(base + Cfg::TABS + Cfg::HDR);
    T2 *tw16 = reinterpret_cast<T2 *>(base + Cfg::TABS + Cfg::HDR + Cfg::UTW);
    unsigned *flat = reinterpret_cast<unsigned *>(base + Cfg::TABS + 64);
    const int tws_n = a.tw_len / NX;
    for (int e = threadIdx.x; e <= MX / 4; e += Cfg::NTHREADS) utw[FFT::SmemTwiddles::tpos(e)] = a.tw[e * tws_n];
    if (threadIdx.x < 16) tw16[threadIdx.x] = a.tw[threadIdx.x * (NX / 256) * tws_n];
    if (threadIdx.x == 0) *flat = 0u;
    __syncthreads();
    if (OUT_H && a.win_x != nullptr) {
      double *swx = reinterpret_cast<double *>(base + Cfg::TABS + Cfg::HDR + Cfg::UTW + Cfg::TW16);
      for (int e = threadIdx.x; e < NX; e += Cfg::NTHREADS) swx[e] = a.win_x[e];
      // bit m of the mask: the x profile is exactly 1 on columns [2 NT m, 2 NT (m+1)) -- the m-th last-stage outputs of
      // every thread; there z * fl(1 * wy) = z * wy needs no profile load (70% of a cosine taper's columns)
      if (threadIdx.x < 16) {
        bool one = true;
        for (int c = 0; c < 2 * NT; c++) one = one && (a.win_x[threadIdx.x * 2 * NT + c] == 1.0);
        if (one) atomicOr(flat, 1u << threadIdx.x);
      }
    }
    __syncthreads();
  }

  for (int i = threadIdx.x / GROUP;; i += NG) {
    // Everything but i is rebuilt from the thread index every iteration (it passes through an empty asm so that the
    // compiler can neither hoist the dozens of derived addresses out of the loop nor keep them live across the
    // transforms: with the register file full, hoisted invariants came back from local memory inside every transform)
    int tid = threadIdx.x;
    asm volatile("" : "+r"(tid));
    const int g = tid / GROUP, gt = tid - g * GROUP;
    const int f = gt / NT, u = gt - f * NT;
    const int grp_bar = 1 + g, row_bar = 1 + NG + g * R + f;
    __builtin_assume(row_bar > 0);   // (the engine's barrier 0 = __syncthreads is never used here)
    int rowtile, plane_i;
    if (!tile_coords(a, i, nplanes, ntiles, rowtile, plane_i)) break;
    unsigned char *base = smem_dyn + ((1024 - (smem_u32(smem_dyn) & 1023)) & 1023);
    const int slot = i % NS;
    T2 *s = reinterpret_cast<T2 *>(base + slot * Cfg::SLOT);
    T2 *row = s + f * PS;
    // twiddles: fetched per stage from the shared-memory table where the plan allows it (frees ~20 registers
    // that otherwise live across both transforms; with them the kernel spilled 84 B / thread)
    constexpr bool STW = FFT::SmemTwiddles::OK;
    typename std::conditional<STW, typename FFT::SmemTwiddles, typename FFT::Twiddles>::type tws;
    if constexpr (STW) {
      tws.tab = reinterpret_cast<const T2 *>(base + Cfg::TABS + Cfg::HDR);
      tws.tab16 = reinterpret_cast<const T2 *>(base + Cfg::TABS + Cfg::HDR + Cfg::UTW);
      tws.u = u;
    } else {
      tws.init(a.tw, a.tw_len / MX, u);
    }
    const long long plane = plane_i;
    T2 keep[16];
    {
      const int iy0 = rowtile * R;
      const long long rowoff = (long long)(iy0 + f) * MX;
      RowTmaStore<T, OUT_H, NT> wst;
      wst.keep = keep;
      wst.map_row = a.map_out != nullptr ? reinterpret_cast<T2 *>(a.map_out + plane * (long long)a.ny * NX) + rowoff : nullptr;
      wst.win_row = nullptr;
      wst.swinx = nullptr;
      wst.wy = 0.0;
      wst.flat = 0u;
      if (OUT_H) {
        if (a.win_x != nullptr) {
          wst.swinx = reinterpret_cast<const double2 *>(base + Cfg::TABS + Cfg::HDR + Cfg::UTW + Cfg::TW16);
          wst.wy = a.win_y[iy0 + f];
          wst.flat = *reinterpret_cast<const unsigned *>(base + Cfg::TABS + 64);
        } else if (a.window != nullptr) {
          wst.win_row = reinterpret_cast<const T2 *>(a.window + (plane / a.group) * a.win_group_stride) + rowoff;
        }
      }
      T2 wu = reinterpret_cast<const T2 *>(base + Cfg::TABS + Cfg::HDR)[FFT::SmemTwiddles::tpos(u)];   // (u < MX/16: inside the table)
      wu.y = -wu.y;  // e^{+2 pi i u/Nx}
      // The tile has landed.  mbarrier waits name a phase by its PARITY only, so the wait for fill k must not start while
      // the slot's barrier is still in phase k-1 (it would be taken for the completed phase k-2 and the group would read
      // the previous tile): that can happen when this group runs a tile and a half ahead of the other one while the TMA
      // unit is backed up -- the load of tile i-3, issued by this group two tiles ago and to be consumed by the other
      // group, has not landed yet (seen with 2-row tiles, whose 32-byte box rows make the TMA unit the bottleneck).  The
      // issuing thread therefore publishes the number of fills issued per slot, and a waiter first sees fill k issued
      // (true long before in the normal case: one shared-memory load), which puts the barrier in phase k or past it.
      {
        unsigned long long *bar = reinterpret_cast<unsigned long long *>(base + Cfg::TABS) + slot;
        const int k = i / NS;
        const volatile int *fills = reinterpret_cast<const volatile int *>(base + Cfg::TABS + 68);
        while (fills[slot] <= k) {}          // fill k has been issued: the barrier is in phase k (or past it)
        __threadfence_block();
        mbar_wait(bar, (unsigned)(k & 1));
      }
      PackLoadTile<T2, MX, R> ld{s, f, wu};
      // first-stage reads come from the tile (all rows interleaved) -> group-wide barrier before the writes
      FFT::template run<+1, true, false>(row, tws, u, row_bar, ld, wst, grp_bar, GROUP);
    }
    if (OUT_H) {
      {
        RegLoad<T2> ld{keep};
        SmemStore<T2> st{row};
        FFT::template run<-1, true, true>(row, tws, u, row_bar, ld, st);
      }
      named_sync(grp_bar, GROUP);
      // transposed store with the r2c unpacking fused in (see fused_row_kernel); w_k from the shared-memory table,
      // second octant by symmetry: w_{Nx/4 - j} = (-Im w_j, -Re w_j) (exact: fused_make_twiddles builds it that way)
      const T2 *utw = reinterpret_cast<const T2 *>(base + Cfg::TABS + Cfg::HDR);
      T2 *dst = a.Hout + plane * (long long)(MX + 1) * a.ny + rowtile * R;
#pragma unroll 4
      for (int e = gt; e < (MX / 2 + 1) * R; e += GROUP) {
        const int k = e / R, r = e - k * R;
        const T2 *zr = s + r * PS;
        const T2 zk = zr[pad(k)], zm = zr[pad(k == 0 ? 0 : MX - k)];
        const bool hi = k > MX / 4;
        const T2 wj = utw[FFT::SmemTwiddles::tpos(hi ? MX / 2 - k : k)];
        T2 w;
        w.x = hi ? -wj.y : wj.x;
        w.y = hi ? -wj.x : wj.y;
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
    // everybody in the group is done with the slot: one thread refills it, nobody waits for that
    named_sync(grp_bar, GROUP);
    if (gt == 0) issue(base, i + NS);
  }
}

// ---- host side -----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn tma_encoder() {
  static EncodeTiledFn fn = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    cudaGetLastError();
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// which row-pass implementation: ORPHX_KB = tma (default where supported) | legacy
inline bool row_tma_enabled() {
  const char *e = getenv("ORPHX_KB");   // read per launch: the tests switch implementations inside one process
  return !(e && !strcmp(e, "legacy"));
}

// 2-row tiles (nx = 4096: the estimator's c2r pass at 4096^2, 1.72 -> 1.42 ms per 8 realisations): ORPHX_KB_R2 = 0 turns them off
inline bool row_tma_r2_enabled() {
  const char *e = getenv("ORPHX_KB_R2");
  return !(e && e[0] == '0');
}
inline int row_tma_tile_order(int dflt) {   // ORPHX_KB_TILE_ORDER = pairs | planes | rows (see tile_coords)
  const char *e = getenv("ORPHX_KB_TILE_ORDER");
  if (e && !strcmp(e, "pairs")) return 2;   // (the launcher only offers it for an even number of row tiles)
  if (e && !strcmp(e, "planes")) return 0;
  if (e && !strcmp(e, "rows")) return 1;
  return dflt;
}
inline CUtensorMapL2promotion row_tma_l2_promotion() {
  const char *e = getenv("ORPHX_KB_L2PROMO");
  const int v = e ? atoi(e) : 0;
  return v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : v == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
       : v == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE;
}

// one configuration (rows per tile, groups, slots) of the persistent TMA row pass
template <typename T, int MX, int MODE, int R, int NG, int NS>
int launch_row_tma_cfg(RowArgs<T> &a, long long nplanes, bool *launched) {
  typedef RowTmaCfg<MX, R, NG, NS> Cfg;
  constexpr size_t SMEM = (MODE & ROW_OUT_H) ? Cfg::SMEM_FULL : Cfg::SMEM_C2R;
  if constexpr (SMEM <= SMEM_MAX && Cfg::NTHREADS <= 1024 && (R == 4 || R == 2)) {
    if (!tma_encoder() || a.ny % R != 0) return OX_OK;
    const long long ntiles = (long long)(a.ny / R) * nplanes;
    if (ntiles >= (1LL << 30) || nplanes >= (1LL << 30)) return OX_OK;
    // the transposed half planes as a 3-D tensor of doubles: [plane][ix = 0..MX][2 * ny]
    CUtensorMap tmap;
    cuuint64_t dims[3] = {(cuuint64_t)2 * a.ny, (cuuint64_t)MX + 1, (cuuint64_t)nplanes};
    cuuint64_t strides[2] = {(cuuint64_t)a.ny * 16, (cuuint64_t)(MX + 1) * a.ny * 16};
    cuuint32_t box[3] = {2 * R, Cfg::BOXW, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapSwizzle swz = R == 4 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    CUresult r = tma_encoder()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void *>((const void *)a.Hin), dims, strides, box,
                               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, row_tma_l2_promotion(),
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled failed (%d) for the row pass %d x %d x %lld", (int)r, a.ny, MX, nplanes);
      return OX_ERR_CUDA;
    }
    auto k = fused_row_tma_kernel<T, MX, R, MODE, NG, NS>;
    OX_TRY(set_smem(k, SMEM));
    int grid = sm_count();   // one persistent CTA per SM
    if ((long long)grid * NG > ntiles) grid = (int)((ntiles + NG - 1) / NG);
    // (with a general 2-D window streamed from L2 the paired order measured 9% slower than single tiles: 2.29 vs 2.10 ms)
    const bool general_window = a.window != nullptr && a.win_x == nullptr;
    constexpr int G = 8 / R;   // row tiles per 128-byte line
    a.tile_order = row_tma_tile_order((a.ny / R) % G == 0 && !general_window ? G : 0);
    if ((a.ny / R) % G != 0 && a.tile_order >= 2) a.tile_order = 0;
    if (a.tile_order >= 2) a.tile_order = G;
    a.nplanes_fast = (int)nplanes;
    k<<<grid, Cfg::NTHREADS, SMEM, g_stream>>>(a, tmap, (int)nplanes, (int)ntiles);
    OX_KERNEL_CHECK();
    *launched = true;
  }
  return OX_OK;
}

// launches the persistent TMA row pass if this (type, size, mode) has one; *launched says whether it did
template <typename T, int MX, int MODE>
int launch_row_tma(RowArgs<T> &a, long long nplanes, bool *launched) {
  *launched = false;
  if constexpr (sizeof(T) == 8 && (MODE == (ROW_IN_H | ROW_OUT_H) || MODE == (ROW_IN_H | ROW_OUT_MAP)) && (MX / 16) % 32 == 0) {
    if (!row_tma_enabled()) return OX_OK;
    constexpr int R = RowCfg<T, MX>::R;
    if constexpr (R == 4 && MODE == (ROW_IN_H | ROW_OUT_H)) {
      // experiment (ORPHX_KB=tma4): FOUR groups of 2-row tiles over six slots -- every scheduler then holds one warp of each group,
      // i.e. four different phases of the tile, instead of two warps of each of two groups
      const char *e = getenv("ORPHX_KB");
      if (e && !strcmp(e, "tma4")) return launch_row_tma_cfg<T, MX, MODE, 2, 4, 6>(a, nplanes, launched);
    }
    // (R = 2, the 32-byte segments of nx = 4096 in 70 KB slots: with single tiles per CTA it measured 2.05 ms against 1.72 ms
    // for the one-tile kernel on the estimator's c2r pass; with blocks of four adjacent row tiles per CTA 1.42 ms)
    if (R == 2 && !row_tma_r2_enabled()) return OX_OK;
    return launch_row_tma_cfg<T, MX, MODE, R, 2, 3>(a, nplanes, launched);
  }
  return OX_OK;
}

}  // namespace oxk
