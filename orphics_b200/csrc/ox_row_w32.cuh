// Warp-per-row version of the persistent TMA row pass for nx = 2048 (MX = 1024 = 32 x 32), 16-byte elements.
//
// ncu on the 16-values-per-thread kernel (ox_row_tma.cuh, profiles/r02_*): the L1/shared-memory data pipe is 80% busy --
// three FFT stages mean two shared-memory exchanges per transform, ~15 passes of the tile through shared memory per
// tile, plus ten named barriers.  Here one WARP owns a row and each lane holds 32 elements: the transform is two
// radix-32 stages with ONE exchange, the exchange only needs __syncwarp, and a tile costs 9 passes:
//   tile -> [pack + radix-32] -> exchange -> [twiddle + radix-32] -> map store, x window (registers)
//        -> [radix-32] -> exchange -> [twiddle + radix-32] -> shared memory -> (k, M-k) unpacking + transposed store.
// Work area layout: pad32(e) = e + e/32 (stride-32 writes and contiguous reads are both conflict free for 16-byte
// elements), rows 1058 elements apart.  Everything around it (TMA tiles into three rotating slots, two groups per
// CTA, static schedule, shared-memory tables) is ox_row_tma.cuh's.
#pragma once

namespace oxk {

__host__ __device__ constexpr int out32(int m) { return 8 * (m & 3) + (m >> 2); }
__host__ __device__ constexpr int pad32(int e) { return e + (e >> 5); }

// cos / sin of 2 pi p / 64, p = 0..31 (the 32-point DFT uses the even entries)
#define OX_C64 {1.0, 0.99518472667219688624, 0.98078528040323044913, 0.95694033573220886494, 0.92387953251128675613, \
                0.88192126434835502971, 0.83146961230254523708, 0.77301045336273696081, 0.70710678118654752440, \
                0.63439328416364549822, 0.55557023301960222474, 0.47139673682599764856, 0.38268343236508977173, \
                0.29028467725446236764, 0.19509032201612826785, 0.09801714032956060199, 0.0, \
                -0.09801714032956060199, -0.19509032201612826785, -0.29028467725446236764, -0.38268343236508977173, \
                -0.47139673682599764856, -0.55557023301960222474, -0.63439328416364549822, -0.70710678118654752440, \
                -0.77301045336273696081, -0.83146961230254523708, -0.88192126434835502971, -0.92387953251128675613, \
                -0.95694033573220886494, -0.98078528040323044913, -0.99518472667219688624}
#define OX_S64 {0.0, 0.09801714032956060199, 0.19509032201612826785, 0.29028467725446236764, 0.38268343236508977173, \
                0.47139673682599764856, 0.55557023301960222474, 0.63439328416364549822, 0.70710678118654752440, \
                0.77301045336273696081, 0.83146961230254523708, 0.88192126434835502971, 0.92387953251128675613, \
                0.95694033573220886494, 0.98078528040323044913, 0.99518472667219688624, 1.0, \
                0.99518472667219688624, 0.98078528040323044913, 0.95694033573220886494, 0.92387953251128675613, \
                0.88192126434835502971, 0.83146961230254523708, 0.77301045336273696081, 0.70710678118654752440, \
                0.63439328416364549822, 0.55557023301960222474, 0.47139673682599764856, 0.38268343236508977173, \
                0.29028467725446236764, 0.19509032201612826785, 0.09801714032956060199}

// 32-point DFT as 4 x 8: input natural order v[n]; X[m] is left at v[out32(m)]
//   X[k1 + 4 k2] = sum_{n2<8} w32^{n2 k1} ( sum_{n1<4} x[8 n1 + n2] w4^{n1 k1} ) w8^{n2 k2}
template <int DIR, typename T2>
__device__ __forceinline__ void dft32(T2 *v) {
  typedef decltype(v[0].x) T;
  constexpr double C[32] = OX_C64, S[32] = OX_S64;
#pragma unroll
  for (int n2 = 0; n2 < 8; n2++) dft4<DIR>(v[n2], v[8 + n2], v[16 + n2], v[24 + n2]);  // -> v[8 k1 + n2]
#pragma unroll
  for (int k1 = 1; k1 < 4; k1++) {
#pragma unroll
    for (int n2 = 1; n2 < 8; n2++) {
      const int p = n2 * k1;  // w32^p, p <= 21
      if (p == 8) v[8 * k1 + n2] = mul_i<DIR>(v[8 * k1 + n2]);
      else if (p == 16) { v[8 * k1 + n2].x = -v[8 * k1 + n2].x; v[8 * k1 + n2].y = -v[8 * k1 + n2].y; }
      else {
        const int q = 2 * p;   // angle 2 pi q / 64; the tables stop at pi
        v[8 * k1 + n2] = mul_cs<DIR>(v[8 * k1 + n2], (T)(q < 32 ? C[q & 31] : -C[q & 31]), (T)(q < 32 ? S[q & 31] : -S[q & 31]));
      }
    }
  }
#pragma unroll
  for (int k1 = 0; k1 < 4; k1++) dft8<DIR>(&v[8 * k1]);  // -> v[8 k1 + k2]
}

// x[r] *= w^r, r = 1..31 (w = w1, conjugated for DIR > 0); powers by products of depth <= 5
template <int DIR, typename T2>
__device__ __forceinline__ void twiddle32(T2 *x, T2 w1) {
  if (DIR > 0) w1.y = -w1.y;
  const T2 w2 = cmul(w1, w1), w3 = cmul(w2, w1), w4 = cmul(w2, w2), w8 = cmul(w4, w4), w16 = cmul(w8, w8);
  x[1] = cmul(x[1], w1);
  x[2] = cmul(x[2], w2);
  x[3] = cmul(x[3], w3);
#pragma unroll
  for (int b = 1; b < 8; b++) {
    T2 wb;
    if (b == 1) wb = w4;
    if (b == 2) wb = w8;
    if (b == 3) wb = cmul(w8, w4);
    if (b == 4) wb = w16;
    if (b == 5) wb = cmul(w16, w4);
    if (b == 6) wb = cmul(w16, w8);
    if (b == 7) wb = cmul(cmul(w16, w8), w4);
    x[4 * b] = cmul(x[4 * b], wb);
    x[4 * b + 1] = cmul(x[4 * b + 1], cmul(wb, w1));
    x[4 * b + 2] = cmul(x[4 * b + 2], cmul(wb, w2));
    x[4 * b + 3] = cmul(x[4 * b + 3], cmul(wb, w3));
  }
}

struct RowW32Cfg {
  static constexpr int MX = 1024, R = 4, GROUP = 32 * R, NTHREADS = 2 * GROUP;
  static constexpr int PS = 1058;                                         // pad32(1023) + 1 = 1055, rounded to 2 (mod 8)
  static constexpr size_t WORK = 16 * (size_t)R * PS;
  static constexpr size_t TILE = 16 * (size_t)R * (MX + 1);
  static constexpr size_t SLOT = (((WORK > TILE ? WORK : TILE) + 511) / 512) * 512;
  static constexpr size_t UTW = 16 * (size_t)(MX / 4 + 1), WINX = 8 * (size_t)(2 * MX);
  static constexpr size_t SMEM = 3 * SLOT + 64 + UTW + WINX + 1024;
  static constexpr int BOXW = 256;
};

// MODE: ROW_IN_H | ROW_OUT_H (full pass; map_out / window are run-time options) or ROW_IN_H | ROW_OUT_MAP (c2r only)
template <int MODE>
__global__ void __launch_bounds__(RowW32Cfg::NTHREADS, 1)
fused_row_w32_kernel(RowArgs<double> a, const __grid_constant__ CUtensorMap tmap, int nplanes, int ntiles) {
  typedef double T;
  typedef double2 T2;
  typedef RowW32Cfg Cfg;
  constexpr bool OUT_H = MODE & ROW_OUT_H;
  constexpr int MX = Cfg::MX, R = Cfg::R, GROUP = Cfg::GROUP, PS = Cfg::PS, NX = 2 * MX;
  extern __shared__ unsigned char smem_dyn[];
  auto issue = [&a, &tmap, nplanes, ntiles](unsigned char *base, int i) {
    const long long t = (long long)blockIdx.x + (long long)i * gridDim.x;
    if (t >= ntiles) return;
    const int rowtile = (int)(t / nplanes), plane = (int)(t - (long long)rowtile * nplanes);
    const int slot = i % 3;
    unsigned char *dst = base + slot * Cfg::SLOT;
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(base + 3 * Cfg::SLOT) + slot;
    fence_proxy_async();   // the slot was last written through the generic proxy (the FFT work area)
    mbar_expect_tx(bar, (unsigned)Cfg::TILE);
#pragma unroll
    for (int j = 0; j < MX / Cfg::BOXW; j++)
      tma_load_3d(dst + (size_t)j * Cfg::BOXW * R * 16, &tmap, 2 * rowtile * R, j * Cfg::BOXW, plane, bar);
    const T2 *nyq = a.Hin + ((long long)plane * (MX + 1) + MX) * a.ny + rowtile * R;
    bulk_load(dst + (size_t)MX * R * 16, nyq, R * 16, bar);
    __threadfence_block();
    reinterpret_cast<volatile int *>(base + 3 * Cfg::SLOT + 40)[slot] = i / 3 + 1;   // fills issued into this slot so far
  };

  {
    unsigned char *base = smem_dyn + ((1024 - (smem_u32(smem_dyn) & 1023)) & 1023);
    if (threadIdx.x == 0) {
      unsigned long long *bars = reinterpret_cast<unsigned long long *>(base + 3 * Cfg::SLOT);
      for (int b = 0; b < 3; b++) {
        mbar_init(bars + b, 1);
        reinterpret_cast<volatile int *>(base + 3 * Cfg::SLOT + 40)[b] = 0;
      }
      fence_mbar_init();
      issue(base, 0);
      issue(base, 1);
      issue(base, 2);
    }
    T2 *utw = reinterpret_cast<T2 *>(base + 3 * Cfg::SLOT + 64);
    const int tws_n = a.tw_len / NX;
    for (int e = threadIdx.x; e <= MX / 4; e += Cfg::NTHREADS) utw[e] = a.tw[e * tws_n];
    if (OUT_H && a.win_x != nullptr) {
      double *swx = reinterpret_cast<double *>(base + 3 * Cfg::SLOT + 64 + Cfg::UTW);
      for (int e = threadIdx.x; e < NX; e += Cfg::NTHREADS) swx[e] = a.win_x[e];
    }
    __syncthreads();
  }

  for (int i = threadIdx.x / GROUP;; i += 2) {
    int tid = threadIdx.x;
    asm volatile("" : "+r"(tid));   // (nothing derived from the thread index is hoisted out of the loop)
    const int g = tid / GROUP, gt = tid - g * GROUP;
    const int f = gt >> 5, j = gt & 31;
    const int grp_bar = 1 + g;
    const int t = blockIdx.x + i * gridDim.x;
    if (t >= ntiles) break;
    unsigned char *base = smem_dyn + ((1024 - (smem_u32(smem_dyn) & 1023)) & 1023);
    const int slot = i % 3;
    T2 *s = reinterpret_cast<T2 *>(base + slot * Cfg::SLOT);
    T2 *row = s + f * PS;
    const T2 *utw = reinterpret_cast<const T2 *>(base + 3 * Cfg::SLOT + 64);
    const int rowtile = t / nplanes;
    const long long plane = t - rowtile * nplanes;
    const int iy0 = rowtile * R;
    T2 v[32];
    {
      // first inverse stage: Z[k] = (X[k] + conj X[M-k]) + i e^{+2 pi i k/Nx} (X[k] - conj X[M-k]), k = j + 32 m, read from
      // the swizzled tile; e^{+2 pi i k/Nx} = e^{+2 pi i j/Nx} e^{2 pi i m/64}
      constexpr double C[32] = OX_C64, S[32] = OX_S64;
      T2 wu = utw[j];
      wu.y = -wu.y;
      {   // the tile has landed (fill k-1 first: see ox_row_tma.cuh)
        unsigned long long *bar = reinterpret_cast<unsigned long long *>(base + 3 * Cfg::SLOT) + slot;
        const int k = i / 3;
        const volatile int *fills = reinterpret_cast<const volatile int *>(base + 3 * Cfg::SLOT + 40);
        while (fills[slot] <= k) {}          // fill k has been issued: the barrier is in phase k (or past it)
        __threadfence_block();
        mbar_wait(bar, (unsigned)(k & 1));
      }
#pragma unroll
      for (int m = 0; m < 32; m++) {
        const int k = j + 32 * m;
        const T2 xk = s[tile_pos<R>(k, f)], xm = s[tile_pos<R>(MX - k, f)];
        T2 w;
        w.x = wu.x * C[m] - wu.y * S[m];
        w.y = wu.x * S[m] + wu.y * C[m];
        const T2 sum = cadd(xk, cconj(xm)), dif = csub(xk, cconj(xm));
        v[m] = cadd(sum, mul_i<+1>(cmul(w, dif)));
      }
      dft32<+1>(v);
      named_sync(grp_bar, GROUP);   // every warp of the group has read the (row-interleaved) tile: the slot becomes the work area
#pragma unroll
      for (int r = 0; r < 32; r++) row[33 * j + r] = v[out32(r)];
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 32; r++) v[r] = row[j + 33 * r];
      twiddle32<+1>(v, utw[2 * j]);
      dft32<+1>(v);   // z[j + 32 r] = x[2n] + i x[2n+1] at v[out32(r)]
    }
    T2 x[32];
    {
      const long long rowoff = (long long)(iy0 + f) * MX;
      T2 *map_row = a.map_out != nullptr ? reinterpret_cast<T2 *>(a.map_out + plane * (long long)a.ny * NX) + rowoff : nullptr;
      if (!OUT_H) {
#pragma unroll
        for (int r = 0; r < 32; r++) st_once(map_row + j + 32 * r, v[out32(r)]);
      } else {
        const T2 *win_row = nullptr;
        const double2 *swinx = nullptr;
        double wy = 0.0;
        if (a.win_x != nullptr) {
          swinx = reinterpret_cast<const double2 *>(base + 3 * Cfg::SLOT + 64 + Cfg::UTW);
          wy = a.win_y[iy0 + f];
        } else if (a.window != nullptr) {
          win_row = reinterpret_cast<const T2 *>(a.window + (plane / a.group) * a.win_group_stride) + rowoff;
        }
#pragma unroll
        for (int r = 0; r < 32; r++) {
          const int n = j + 32 * r;
          T2 z = v[out32(r)];
          if (map_row != nullptr) st_once(map_row + n, z);
          if (swinx != nullptr) {
            const double2 p = swinx[n];
            z.x *= p.x * wy;   // the window value itself is formed in float64 and rounded once, as numpy forms it
            z.y *= p.y * wy;
          } else if (win_row != nullptr) {
            const T2 w1 = ldg2(win_row + n);
            z.x *= w1.x;
            z.y *= w1.y;
          }
          x[r] = z;
        }
      }
    }
    if (OUT_H) {
      dft32<-1>(x);
      __syncwarp();   // the lanes of this row have finished reading the inverse transform's exchange
#pragma unroll
      for (int r = 0; r < 32; r++) row[33 * j + r] = x[out32(r)];
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 32; r++) x[r] = row[j + 33 * r];
      twiddle32<-1>(x, utw[2 * j]);
      dft32<-1>(x);   // Z[j + 32 r] at x[out32(r)]
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 32; r++) row[j + 33 * r] = x[out32(r)];   // = row[pad32(j + 32 r)]
      named_sync(grp_bar, GROUP);
      // transposed store with the r2c unpacking fused in (see fused_row_kernel), two outputs per pair (k, M-k)
      T2 *dst = a.Hout + plane * (long long)(MX + 1) * a.ny + iy0;
#pragma unroll 4
      for (int e = gt; e < (MX / 2 + 1) * R; e += GROUP) {
        const int k = e / R, r = e - k * R;
        const T2 *zr = s + r * PS;
        const T2 zk = zr[pad32(k)], zm = zr[pad32(k == 0 ? 0 : MX - k)];
        const bool hi = k > MX / 4;
        const T2 wj = utw[hi ? MX / 2 - k : k];
        T2 w;
        w.x = hi ? -wj.y : wj.x;
        w.y = hi ? -wj.x : wj.y;
        const T2 sum = cadd(zk, cconj(zm)), dif = csub(zk, cconj(zm));
        const T2 pw = mul_i<+1>(cmul(w, dif));
        T2 x0, x1;
        x0.x = 0.5 * (sum.x - pw.x);
        x0.y = 0.5 * (sum.y - pw.y);
        x1.x = 0.5 * (sum.x + pw.x);
        x1.y = -0.5 * (sum.y + pw.y);
        st_once(dst + (long long)k * a.ny + r, x0);
        if (2 * k != MX) st_once(dst + (long long)(MX - k) * a.ny + r, x1);
      }
    }
    // everybody in the group is done with the slot: one thread refills it, nobody waits for that
    named_sync(grp_bar, GROUP);
    if (gt == 0) issue(base, i + 3);
  }
}

// ORPHX_KB=w32 selects this kernel (an experiment kept for the record: it halves the shared-memory traffic -- L1 data
// pipe 43% busy against 80% -- but with 254 registers only 8 warps fit an SM, and it measures 1.87 ms per 64 maps
// against 1.58 ms for the 16-values-per-thread kernel, 2.65 against 2.10 with a general 2-D window;
// profiles/r02_variants.txt)
inline bool row_w32_enabled() {
  const char *e = getenv("ORPHX_KB");
  return e && !strcmp(e, "w32");
}

template <typename T, int MX, int MODE>
int launch_row_w32(RowArgs<T> &a, long long nplanes, bool *launched) {
  *launched = false;
  if constexpr (sizeof(T) == 8 && MX == 1024 && (MODE == (ROW_IN_H | ROW_OUT_H) || MODE == (ROW_IN_H | ROW_OUT_MAP))) {
    typedef RowW32Cfg Cfg;
    static_assert(Cfg::SMEM <= SMEM_MAX, "shared memory");
    if (!row_w32_enabled() || !tma_encoder() || a.ny % Cfg::R != 0) return OX_OK;
    const long long ntiles = (long long)(a.ny / Cfg::R) * nplanes;
    if (ntiles >= (1LL << 30) || nplanes >= (1LL << 30)) return OX_OK;
    CUtensorMap tmap;
    cuuint64_t dims[3] = {(cuuint64_t)2 * a.ny, (cuuint64_t)MX + 1, (cuuint64_t)nplanes};
    cuuint64_t strides[2] = {(cuuint64_t)a.ny * 16, (cuuint64_t)(MX + 1) * a.ny * 16};
    cuuint32_t box[3] = {2 * Cfg::R, Cfg::BOXW, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = tma_encoder()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void *>((const void *)a.Hin), dims, strides, box,
                               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled failed (%d) for the row pass %d x %d x %lld", (int)r, a.ny, MX, nplanes);
      return OX_ERR_CUDA;
    }
    auto k = fused_row_w32_kernel<MODE>;
    OX_TRY(set_smem(k, Cfg::SMEM));
    int grid = sm_count();   // one persistent CTA per SM
    if ((long long)grid * 2 > ntiles) grid = (int)((ntiles + 1) / 2);
    a.nplanes_fast = (int)nplanes;
    k<<<grid, Cfg::NTHREADS, Cfg::SMEM, g_stream>>>(a, tmap, (int)nplanes, (int)ntiles);
    OX_KERNEL_CHECK();
    *launched = true;
  }
  return OX_OK;
}

}  // namespace oxk
