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

struct RowTmaCtl {
  unsigned long long full[3];  // one mbarrier per slot: the tile's bytes have landed
  int next_buf;                // slot holding the tile in flight / landed that nobody owns yet (-1: being handed over)
  int issued;                  // tiles of this CTA issued so far
  int tile_of[3], par_of[3], nloads[3];
  int grp_slot[2], grp_tile[2], grp_par[2];
};

template <int MX, int R>
struct RowTmaCfg {
  static constexpr int NT = MX / 16, GROUP = R * NT, NTHREADS = 2 * GROUP;
  static constexpr size_t WORK = 16 * (size_t)R * padded_size(MX);        // padded per-row work area
  static constexpr size_t TILE = 16 * (size_t)R * (MX + 1);               // raw tile
  static constexpr size_t SLOT = (((WORK > TILE ? WORK : TILE) + 1023) / 1024) * 1024;
  static constexpr size_t SMEM = 3 * SLOT + sizeof(RowTmaCtl) + 1024;     // (+ alignment slack)
  static constexpr int BOXW = 256;                                        // columns per TMA box
  static_assert(MX % BOXW == 0, "MX must be a multiple of the TMA box width");
};

// MODE: ROW_IN_H | ROW_OUT_H (full pass; map_out / window are run-time options) or ROW_IN_H | ROW_OUT_MAP (c2r only)
template <typename T, int MX, int R, int MODE>
__global__ void __launch_bounds__(RowTmaCfg<MX, R>::NTHREADS, (RowTmaCfg<MX, R>::NTHREADS <= 256 ? 2 : 1))
fused_row_tma_kernel(RowArgs<T> a, const __grid_constant__ CUtensorMap tmap, int nplanes, int ntiles) {
  static_assert(sizeof(T) == 8, "the TMA row pass is written for 16-byte elements");
  constexpr bool OUT_MAP = MODE & ROW_OUT_MAP, WIN = MODE & ROW_WIN, OUT_H = MODE & ROW_OUT_H;
  typedef typename V2<T>::type T2;
  typedef BlockFFT<T, MX> FFT;
  typedef RowTmaCfg<MX, R> Cfg;
  constexpr int NT = Cfg::NT, GROUP = Cfg::GROUP, PS = padded_size(MX), NX = 2 * MX;
  extern __shared__ unsigned char smem_dyn[];
  // slots aligned to 1024 B (the swizzle pattern repeats every 512 B of shared-memory address)
  unsigned char *base = smem_dyn + ((1024 - (smem_u32(smem_dyn) & 1023)) & 1023);
  RowTmaCtl *ctl = reinterpret_cast<RowTmaCtl *>(base + 3 * Cfg::SLOT);
  // one thread: fetch the i-th tile of this CTA into `slot` (or mark the end of the CTA's tiles)
  auto issue = [&a, &tmap, nplanes, ntiles](unsigned char *base, RowTmaCtl *ctl, int slot, int i) {
    const unsigned tile_bytes = (unsigned)Cfg::TILE;
    const long long t = (long long)blockIdx.x + (long long)i * gridDim.x;
    if (t >= ntiles) {
      ctl->tile_of[slot] = -1;
      return;
    }
    const int rowtile = (int)(t / nplanes), plane = (int)(t - (long long)rowtile * nplanes);
    ctl->tile_of[slot] = (int)t;
    ctl->par_of[slot] = ctl->nloads[slot] & 1;
    ctl->nloads[slot]++;
    unsigned char *dst = base + slot * Cfg::SLOT;
    fence_proxy_async();   // the slot was last written through the generic proxy (the FFT work area)
    mbar_expect_tx(&ctl->full[slot], tile_bytes);
#pragma unroll
    for (int j = 0; j < MX / Cfg::BOXW; j++)
      tma_load_3d(dst + (size_t)j * Cfg::BOXW * R * 16, &tmap, 2 * rowtile * R, j * Cfg::BOXW, plane, &ctl->full[slot]);
    const T2 *nyq = a.Hin + ((long long)plane * (MX + 1) + MX) * a.ny + rowtile * R;
    bulk_load(dst + (size_t)MX * R * 16, nyq, R * 16, &ctl->full[slot]);
  };

  if (threadIdx.x == 0) {
    for (int b = 0; b < 3; b++) {
      mbar_init(&ctl->full[b], 1);
      ctl->nloads[b] = 0;
    }
    fence_mbar_init();
    fence_proxy_async();
    issue(base, ctl, 0, 0);
    issue(base, ctl, 1, 1);
    issue(base, ctl, 2, 2);
    ctl->issued = 3;
    ctl->next_buf = 2;
    for (int q = 0; q < 2; q++) {
      ctl->grp_slot[q] = q;
      ctl->grp_tile[q] = ctl->tile_of[q];
      ctl->grp_par[q] = ctl->par_of[q];
    }
  }
  __syncthreads();

  while (true) {
    // Everything is rebuilt from the thread index every iteration (it passes through an empty asm so that the
    // compiler can neither hoist the dozens of derived addresses out of the loop nor keep them live across the
    // transforms: with the register file full, hoisted invariants came back from local memory -- L2 round trips --
    // inside every transform)
    int tid = threadIdx.x;
    asm volatile("" : "+r"(tid));
    const int g = tid / GROUP, gt = tid - g * GROUP;
    const int f = gt / NT, u = gt - f * NT;
    const int grp_bar = 1 + g, row_bar = 3 + g * R + f;
    const int tws_n = a.tw_len / NX;  // stride for exp(-2 pi i k / Nx)
    unsigned char *base = smem_dyn + ((1024 - (smem_u32(smem_dyn) & 1023)) & 1023);
    RowTmaCtl *ctl = reinterpret_cast<RowTmaCtl *>(base + 3 * Cfg::SLOT);
    typename FFT::Twiddles tws;
    tws.init(a.tw, a.tw_len / MX, u);
    const int slot = ctl->grp_slot[g];
    int t = ctl->grp_tile[g];
    if (t < 0) break;
    T2 *s = reinterpret_cast<T2 *>(base + slot * Cfg::SLOT);
    T2 *row = s + f * PS;
    T2 keep[16];
    {
      const int rowtile = t / nplanes;
      const long long plane = t - (long long)rowtile * nplanes;
      const int iy0 = rowtile * R;
      const long long rowoff = (long long)(iy0 + f) * MX;
      T2 wpf[16];
      constexpr bool RT = OUT_H;  // the full pass takes map_out / window as run-time options
      WindowKeep<T, OUT_MAP, WIN, RT, NT> wst;
      wst.keep = keep;
      wst.w = wpf;
      wst.u = u;
      wst.map_row = (RT ? a.map_out != nullptr : OUT_MAP) ? reinterpret_cast<T2 *>(a.map_out + plane * (long long)a.ny * NX) + rowoff : nullptr;
      const long long grp = plane / a.group;
      wst.win_row = (RT ? a.window != nullptr : WIN) ? reinterpret_cast<const T2 *>(a.window + grp * a.win_group_stride) + rowoff : nullptr;
      if (RT && a.win_x != nullptr) {
        wst.winx = reinterpret_cast<const double2 *>(a.win_x);
        wst.wy = a.win_y[iy0 + f];
      }
      T2 wu = a.tw[u * tws_n];
      wu.y = -wu.y;  // e^{+2 pi i u/Nx}
      mbar_wait(&ctl->full[slot], (unsigned)ctl->grp_par[g]);   // the TMA has delivered the tile
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
      t = *(volatile int *)&ctl->grp_tile[g];   // (re-read: nothing of the tile's bookkeeping stays live across the transforms)
      const int rowtile = t / nplanes;
      const long long plane = t - (long long)rowtile * nplanes;
      T2 *dst = a.Hout + plane * (long long)(MX + 1) * a.ny + rowtile * R;
#pragma unroll 4
      for (int e = gt; e < (MX / 2 + 1) * R; e += GROUP) {
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
    // everybody in the group is done with the slot: hand it to the TMA, take over the prefetched one
    named_sync(grp_bar, GROUP);
    if (gt == 0) {
      int nb;
      do {
        nb = atomicExch(&ctl->next_buf, -1);
      } while (nb < 0);
      __threadfence_block();
      const int i = atomicAdd(&ctl->issued, 1);
      const int myslot = *(volatile int *)&ctl->grp_slot[g];
      issue(base, ctl, myslot, i);
      __threadfence_block();
      atomicExch(&ctl->next_buf, myslot);
      ctl->grp_slot[g] = nb;
      ctl->grp_tile[g] = ctl->tile_of[nb];
      ctl->grp_par[g] = ctl->par_of[nb];
    }
    named_sync(grp_bar, GROUP);
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

// launches the persistent TMA row pass if this (type, size, mode) has one; *launched says whether it did
template <typename T, int MX, int MODE>
int launch_row_tma(RowArgs<T> &a, long long nplanes, bool *launched) {
  *launched = false;
  if constexpr (sizeof(T) == 8 && (MODE == (ROW_IN_H | ROW_OUT_H) || MODE == (ROW_IN_H | ROW_OUT_MAP)) && (MX / 16) % 32 == 0) {
    constexpr int R = RowCfg<T, MX>::R;
    typedef RowTmaCfg<MX, R> Cfg;
    if constexpr (Cfg::SMEM <= SMEM_MAX && Cfg::NTHREADS <= 1024 && (R == 1 || R == 2 || R == 4)) {
      if (!row_tma_enabled() || !tma_encoder() || a.ny % R != 0) return OX_OK;
      const long long ntiles = (long long)(a.ny / R) * nplanes;
      if (ntiles >= (1LL << 31) || nplanes >= (1LL << 31)) return OX_OK;
      // the transposed half planes as a 3-D tensor of doubles: [plane][ix = 0..MX][2 * ny]
      CUtensorMap tmap;
      cuuint64_t dims[3] = {(cuuint64_t)2 * a.ny, (cuuint64_t)MX + 1, (cuuint64_t)nplanes};
      cuuint64_t strides[2] = {(cuuint64_t)a.ny * 16, (cuuint64_t)(MX + 1) * a.ny * 16};
      cuuint32_t box[3] = {2 * R, Cfg::BOXW, 1};
      cuuint32_t estr[3] = {1, 1, 1};
      const CUtensorMapSwizzle swz = R == 4 ? CU_TENSOR_MAP_SWIZZLE_64B : (R == 2 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE);
      CUresult r = tma_encoder()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void *>((const void *)a.Hin), dims, strides, box,
                                 estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for the row pass %d x %d x %lld", (int)r, a.ny, MX, nplanes);
        return OX_ERR_CUDA;
      }
      auto k = fused_row_tma_kernel<T, MX, R, MODE>;
      OX_TRY(set_smem(k, Cfg::SMEM));
      // one persistent CTA per SM (two where 256-thread CTAs and their three slots fit twice)
      const int per_sm = (Cfg::NTHREADS <= 256 && 2 * Cfg::SMEM <= SMEM_MAX) ? 2 : 1;
      int grid = sm_count() * per_sm;
      if ((long long)grid * 2 > ntiles) grid = (int)((ntiles + 1) / 2);
      a.nplanes_fast = (int)nplanes;
      k<<<grid, Cfg::NTHREADS, Cfg::SMEM, g_stream>>>(a, tmap, (int)nplanes, (int)ntiles);
      OX_KERNEL_CHECK();
      *launched = true;
    }
  }
  return OX_OK;
}

}  // namespace oxk
