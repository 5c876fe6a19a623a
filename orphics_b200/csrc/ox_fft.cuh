// Block-level FFT engine for the fused sim -> FFT -> power -> bin kernels (ox_fused.cu).
//
// A power-of-two complex FFT of length L = 16^a * rem (rem in {1,2,4,8}) is executed by
// L/16 threads, each holding a "unit" of 16 elements in registers.  Stages are Stockham
// autosort passes: radix 16 while possible, then one radix-rem stage; between stages the
// data is exchanged through shared memory (write to the autosorted positions, barrier,
// contiguous read).  Shared-memory bandwidth is what bounds these kernels on B200 (ncu:
// L1/shared pipe 75% busy with the earlier radix-8 engine, FP64 pipe 20%), so the design
// minimises round trips:
//   * radix 16 -> 2 exchanges for L <= 2048 (the radix-8 version needed 3-4);
//   * the first stage takes its input from a caller-supplied functor (registers: generated
//     noise, global memory, or a packed view of shared memory) and the last stage hands its
//     output to a functor (global store, window multiply, binning) -- no extra passes;
//   * buffers are padded by one element every 16 (pad(e) = e + e/16): the stride-16 writes
//     of the first stage and every contiguous read are bank-conflict free for 16 B and 8 B
//     elements.
// Twiddles: a thread works on the same butterflies in every transform, so the first-order
// twiddle w = exp(-2 pi i k / (Ns R)) of each stage is loaded once from a global table
// (tw[j] = exp(-2 pi i j / LT), built on the host in long double) into registers; w^2..w^15
// are formed by complex multiplications.  DIR = -1 forward, +1 backward (unnormalised).
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

namespace oxfft {

// A last-stage StoreOp may declare `static constexpr bool PREFETCH = true` and provide pre(m) (start the global
// loads it needs for output m) and use(f, v, m) (consume output m with the prefetched operands): the engine then
// issues the loads of butterfly q+1 while butterfly q is being stored, instead of one exposed latency per element.
template <class S, class = void>
struct has_prefetch : std::false_type {};
template <class S>
struct has_prefetch<S, std::void_t<decltype(S::PREFETCH)>> : std::integral_constant<bool, S::PREFETCH> {};

template <typename T>
struct V2;
template <>
struct V2<double> {
  typedef double2 type;
};
template <>
struct V2<float> {
  typedef float2 type;
};

__host__ __device__ constexpr int pad(int e) { return e + (e >> 4); }
// elements to allocate for one padded length-L buffer that may also hold index L (the
// Nyquist element of a real transform); sizes are kept = 2 (mod 8) so that consecutive
// double2 buffers start 32 bytes apart modulo 128 and equal indices of neighbouring rows
// fall into different banks
__host__ __device__ constexpr int padded_size(int L) { return ((L + (L >> 4) + 1 + 7) / 8) * 8 + 2; }

template <typename T2>
__device__ __forceinline__ T2 cadd(T2 a, T2 b) {
  T2 r;
  r.x = a.x + b.x;
  r.y = a.y + b.y;
  return r;
}
template <typename T2>
__device__ __forceinline__ T2 csub(T2 a, T2 b) {
  T2 r;
  r.x = a.x - b.x;
  r.y = a.y - b.y;
  return r;
}
template <typename T2>
__device__ __forceinline__ T2 cmul(T2 a, T2 b) {
  T2 r;
  r.x = a.x * b.x - a.y * b.y;
  r.y = a.x * b.y + a.y * b.x;
  return r;
}
template <typename T2>
__device__ __forceinline__ T2 cconj(T2 a) {
  a.y = -a.y;
  return a;
}
// multiply by e^{DIR * i pi/2} = DIR * i
template <int DIR, typename T2>
__device__ __forceinline__ T2 mul_i(T2 a) {
  T2 r;
  if (DIR > 0) {
    r.x = -a.y;
    r.y = a.x;
  } else {
    r.x = a.y;
    r.y = -a.x;
  }
  return r;
}
// multiply by (c + DIR*i*s)
template <int DIR, typename T2, typename T>
__device__ __forceinline__ T2 mul_cs(T2 a, T c, T s) {
  T2 r;
  if (DIR > 0) {
    r.x = a.x * c - a.y * s;
    r.y = a.y * c + a.x * s;
  } else {
    r.x = a.x * c + a.y * s;
    r.y = a.y * c - a.x * s;
  }
  return r;
}

template <int DIR, typename T2>
__device__ __forceinline__ void dft2(T2 &a, T2 &b) {
  T2 t = a;
  a = cadd(t, b);
  b = csub(t, b);
}

// 4-point DFT, natural order in and out
template <int DIR, typename T2>
__device__ __forceinline__ void dft4(T2 &x0, T2 &x1, T2 &x2, T2 &x3) {
  T2 a0 = cadd(x0, x2), a2 = csub(x0, x2);
  T2 a1 = cadd(x1, x3), a3 = mul_i<DIR>(csub(x1, x3));
  x0 = cadd(a0, a1);
  x1 = cadd(a2, a3);
  x2 = csub(a0, a1);
  x3 = csub(a2, a3);
}

// 8-point DFT, natural order in and out
template <int DIR, typename T2>
__device__ __forceinline__ void dft8(T2 *v) {
  typedef decltype(v[0].x) T;
  const T h = (T)0.70710678118654752440;
  T2 a0 = cadd(v[0], v[4]), a4 = csub(v[0], v[4]);
  T2 a1 = cadd(v[1], v[5]), a5 = csub(v[1], v[5]);
  T2 a2 = cadd(v[2], v[6]), a6 = csub(v[2], v[6]);
  T2 a3 = cadd(v[3], v[7]), a7 = csub(v[3], v[7]);
  a5 = mul_cs<DIR>(a5, h, h);
  a6 = mul_i<DIR>(a6);
  a7 = mul_cs<DIR>(a7, -h, h);
  T2 b0 = cadd(a0, a2), b2 = csub(a0, a2), b1 = cadd(a1, a3), b3 = mul_i<DIR>(csub(a1, a3));
  T2 b4 = cadd(a4, a6), b6 = csub(a4, a6), b5 = cadd(a5, a7), b7 = mul_i<DIR>(csub(a5, a7));
  v[0] = cadd(b0, b1); v[4] = csub(b0, b1);
  v[2] = cadd(b2, b3); v[6] = csub(b2, b3);
  v[1] = cadd(b4, b5); v[5] = csub(b4, b5);
  v[3] = cadd(b6, b7); v[7] = csub(b6, b7);
}

// 16-point DFT as 4 x 4.  Input natural order v[n]; the result X[m] is left at
// v[4*(m&3) + (m>>2)] (use out16(m)); the transposition is resolved at compile time.
__host__ __device__ constexpr int out16(int m) { return 4 * (m & 3) + (m >> 2); }

template <int DIR, typename T2>
__device__ __forceinline__ void dft16(T2 *v) {
  typedef decltype(v[0].x) T;
  const T c1 = (T)0.92387953251128675613, s1 = (T)0.38268343236508977173, h = (T)0.70710678118654752440;
  // X[k1 + 4 k2] = sum_n2 w16^{n2 k1} ( sum_n1 x[4 n1 + n2] w4^{n1 k1} ) w4^{n2 k2}
#pragma unroll
  for (int n2 = 0; n2 < 4; n2++) dft4<DIR>(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);  // -> v[4 k1 + n2]
  // twiddles w16^{n2 k1}
  v[5] = mul_cs<DIR>(v[5], c1, s1);    // k1=1,n2=1 : w^1
  v[6] = mul_cs<DIR>(v[6], h, h);      // k1=1,n2=2 : w^2
  v[7] = mul_cs<DIR>(v[7], s1, c1);    // k1=1,n2=3 : w^3
  v[9] = mul_cs<DIR>(v[9], h, h);      // k1=2,n2=1 : w^2
  v[10] = mul_i<DIR>(v[10]);           // k1=2,n2=2 : w^4
  v[11] = mul_cs<DIR>(v[11], -h, h);   // k1=2,n2=3 : w^6
  v[13] = mul_cs<DIR>(v[13], s1, c1);  // k1=3,n2=1 : w^3
  v[14] = mul_cs<DIR>(v[14], -h, h);   // k1=3,n2=2 : w^6
  v[15] = mul_cs<DIR>(v[15], -c1, -s1);  // k1=3,n2=3 : w^9 = -w^1
#pragma unroll
  for (int k1 = 0; k1 < 4; k1++) dft4<DIR>(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);  // -> v[4 k1 + k2]
}

// barrier over the threads of one transform: id 0 = the whole CTA (__syncthreads), id 1..15 = a
// named barrier shared by the `count` threads working on this transform, so that independent
// transforms in one CTA do not wait for each other (count must be a multiple of 32)
__device__ __forceinline__ void fft_sync(int bar, int count) {
  if (bar == 0) __syncthreads();
  else asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(count) : "memory");
}

template <int L>
struct Plan16 {
  static constexpr int log2() {
    int n = 0, l = L;
    while (l > 1) { l >>= 1; n++; }
    return n;
  }
  static constexpr int N16 = log2() / 4;             // radix-16 stages
  static constexpr int REM = 1 << (log2() % 4);     // radix of the final stage (1 = none)
  static constexpr int QREM = REM > 1 ? 16 / REM : 0;
  // first-order twiddles per thread: one per radix-16 stage after the first, QREM for the final stage
  static constexpr int NW = (N16 > 1 ? N16 - 1 : 0) + (N16 > 0 ? QREM : 0);
};

// default functors: plain padded shared-memory access
// default functors: plain padded shared-memory access.  A thread's first-stage inputs and its
// last-stage outputs are both the elements u + m*NT, m = 0..15; m is passed to the functors as
// a compile-time-constant (after unrolling) register index.
template <typename T2>
struct SmemLoad {
  const T2 *s;
  __device__ __forceinline__ T2 operator()(int e, int /*m*/) const { return s[pad(e)]; }
};
template <typename T2>
struct SmemStore {
  T2 *s;
  __device__ __forceinline__ void operator()(int f, T2 v, int /*m*/) const { s[pad(f)] = v; }
};

// FFT of length L in shared memory `s` (padded_size(L) elements) by NT = L/16 threads, thread
// index u in [0, NT).  Every thread of the barrier group must call run().
//   LoadOp  ld(e, m)    -> input element e = u + m*NT (natural order) of the first stage
//   StoreOp st(f, v, m) -> consumes output element f = u + m*NT (natural order) of the last stage
// IN_SMEM : the first stage reads shared memory that other threads of the group may still be
//           writing/reading -> the caller must have synchronised; also tells the engine that a
//           barrier is needed between the first stage's reads and its writes.
// OUT_SMEM: the StoreOp writes the transform's own buffer `s` -> a final barrier is issued.
template <typename T, int L>
struct BlockFFT {
  typedef typename V2<T>::type T2;
  typedef Plan16<L> P;
  static constexpr int NT = L / 16;
  static constexpr int NW = P::NW;
  static_assert(L >= 32 && (L & (L - 1)) == 0, "L must be a power of two >= 32");

  struct Twiddles {
    T2 w[NW > 0 ? NW : 1];
    __device__ __forceinline__ T2 get(int i) const { return w[i]; }
    // tw: table exp(-2 pi i j / (L*tw_stride)), j < L*tw_stride
    __device__ __forceinline__ void init(const T2 *__restrict__ tw, int tw_stride, int u) {
      int o = 0;
      int Ns = 16;
#pragma unroll
      for (int st = 1; st < P::N16; st++) {  // radix-16 stages after the first: j = u
        const int k = u & (Ns - 1);
        w[o++] = tw[k * (L / (Ns * 16)) * tw_stride];
        Ns *= 16;
      }
      if (P::REM > 1) {
#pragma unroll
        for (int q = 0; q < P::QREM; q++) {
          const int j = u + q * NT;
          const int k = j & (Ns - 1);
          w[o++] = tw[k * (L / (Ns * P::REM)) * tw_stride];
        }
      }
    }
  };

  // The same first-order twiddles fetched where a stage needs them from a shared-memory table
  // tab[j] = exp(-2 pi i j / (2L)), j <= L/4 (second octant by the exact symmetry of fused_make_twiddles' table:
  // tab[L/2 - j] = (-Im tab[j], -Re tab[j])), instead of living in ~4 NW registers across every transform of a
  // persistent kernel.  Needs every exponent below L/2 in units of 1/(2L): REM in {1, 4, 8}.
  struct SmemTwiddles {
    static constexpr bool OK = P::REM == 1 || P::REM == 4 || P::REM == 8;
    // The table is stored split by parity -- tab[tpos(j)], tpos(j) = j/2 for even j, L/8 + 1 + j/2 for odd j -- so that
    // the last stage's exponents (even, consecutive lanes 2 apart) are contiguous 16-byte elements, and the 16
    // exponents of the first radix-16 twiddle stage (8 (L/1024) apart) have a compact copy tab16[k].
    static constexpr int TLEN = L / 4 + 1;
    static __host__ __device__ constexpr int tpos(int j) { return (j & 1) ? L / 8 + 1 + (j >> 1) : (j >> 1); }
    const T2 *tab;
    const T2 *tab16;   // tab16[k] = exp(-2 pi i k / 256), k < 16 (used when N16 >= 2)
    int u;
    __device__ __forceinline__ T2 get(int i) const {
      // i is a compile-time constant after unrolling: i < N16-1 -> radix-16 stage i+1, else butterfly q of the last stage
      if (P::N16 >= 2 && i == 0) return tab16[u & 15];
      int Ns = 16, idx = 0;
      bool found = false;
#pragma unroll
      for (int st = 1; st < P::N16; st++) {
        if (i == st - 1) {
          idx = 2 * (u & (Ns - 1)) * (L / (Ns * 16));
          found = true;
        }
        Ns *= 16;
      }
      if (!found) {
        const int q = i - (P::N16 > 1 ? P::N16 - 1 : 0);
        const int j = u + q * NT;
        idx = 2 * (j & (Ns - 1)) * (L / (Ns * (P::REM > 1 ? P::REM : 1)));
      }
      const bool hi = idx > L / 4;
      const T2 t = tab[tpos(hi ? L / 2 - idx : idx)];
      T2 r;
      r.x = hi ? -t.y : t.x;
      r.y = hi ? -t.x : t.y;
      return r;
    }
  };

  // multiply x[1..R-1] by w^1..w^(R-1)
  template <int DIR, int R>
  static __device__ __forceinline__ void apply_twiddles(T2 *x, T2 w1) {
    if (DIR > 0) w1.y = -w1.y;
    x[1] = cmul(x[1], w1);
    if (R >= 4) {
      T2 w2 = cmul(w1, w1), w3 = cmul(w2, w1);
      x[2] = cmul(x[2], w2);
      x[3] = cmul(x[3], w3);
      if (R >= 8) {
        T2 w4 = cmul(w2, w2);
        x[4] = cmul(x[4], w4);
        x[5] = cmul(x[5], cmul(w4, w1));
        x[6] = cmul(x[6], cmul(w3, w3));
        T2 w7 = cmul(w4, w3);
        x[7] = cmul(x[7], w7);
        if (R == 16) {
          T2 w8 = cmul(w4, w4);
          x[8] = cmul(x[8], w8);
          x[9] = cmul(x[9], cmul(w8, w1));
          x[10] = cmul(x[10], cmul(w8, w2));
          x[11] = cmul(x[11], cmul(w8, w3));
          x[12] = cmul(x[12], cmul(w8, w4));
          x[13] = cmul(x[13], cmul(w7, cmul(w3, w3)));
          x[14] = cmul(x[14], cmul(w7, w7));
          x[15] = cmul(x[15], cmul(w8, w7));
        }
      }
    }
  }

  // one Stockham stage of radix R (Q = 16/R butterflies per unit), Ns = product of previous radices
  template <int DIR, int R, int Ns, int WOFF, bool FIRST, bool LAST, bool SYNC_BEFORE_WRITE, bool SYNC_AFTER_WRITE,
            class TW, class LoadOp, class StoreOp>
  static __device__ __forceinline__ void stage(T2 *__restrict__ s, const TW &tws, int u, int bar, const LoadOp &ld,
                                               const StoreOp &st, int bar_bw = -1, int cnt_bw = 0) {
    constexpr int Q = 16 / R;
    constexpr bool PF = LAST && Q > 1 && has_prefetch<StoreOp>::value;
    T2 v[16];
    if constexpr (PF) {
#pragma unroll
      for (int r = 0; r < R; r++) st.pre(r * Q);  // operands of the first butterfly's outputs
    }
#pragma unroll
    for (int q = 0; q < Q; q++) {
#pragma unroll
      for (int r = 0; r < R; r++) {
        const int e = u + (q + r * Q) * NT;  // = j + r*L/R with j = u + q*NT
        v[q * R + r] = FIRST ? ld(e, q + r * Q) : s[pad(e)];
      }
      if (Ns > 1) apply_twiddles<DIR, R>(&v[q * R], tws.get(WOFF + q));
      if (R == 16) dft16<DIR>(&v[0]);
      if (R == 8) dft8<DIR>(&v[q * 8]);
      if (R == 4) dft4<DIR>(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
      if (R == 2) dft2<DIR>(v[q * 2], v[q * 2 + 1]);
    }
    if (SYNC_BEFORE_WRITE) {
      // (bar_bw >= 0: the first stage read a buffer shared with OTHER transforms -- e.g. a TMA tile holding
      // several rows interleaved -- so its reads are fenced by the wider barrier of cnt_bw threads)
      if (bar_bw >= 0) fft_sync(bar_bw, cnt_bw);
      else fft_sync(bar, NT);
    }
#pragma unroll
    for (int q = 0; q < Q; q++) {
      const int j = u + q * NT;
      const int k = j & (Ns - 1);
      const int d = (j - k) * R + k;
      if constexpr (PF) {
        if (q + 1 < Q) {
#pragma unroll
          for (int r = 0; r < R; r++) st.pre(q + 1 + r * Q);
        }
      }
#pragma unroll
      for (int r = 0; r < R; r++) {
        const T2 val = v[q * R + (R == 16 ? out16(r) : r)];
        if constexpr (PF) st.use(d + r * Ns, val, q + r * Q);
        else if (LAST) st(d + r * Ns, val, q + r * Q);  // here Ns = L/R, so d + r*Ns = u + (q + r*Q)*NT
        else s[pad(d + r * Ns)] = val;
      }
    }
    if (SYNC_AFTER_WRITE) fft_sync(bar, NT);
  }

  template <int DIR, int I, int Ns, bool IN_SMEM, bool OUT_SMEM, class TW, class LoadOp, class StoreOp>
  static __device__ __forceinline__ void from(T2 *__restrict__ s, const TW &tws, int u, int bar, const LoadOp &ld,
                                              const StoreOp &st, int bar0 = -1, int cnt0 = 0) {
    constexpr int rem = L / Ns;
    constexpr bool first = (I == 0);
    // the reads of a stage must complete before its writes unless the reads did not touch smem
    constexpr bool sync_bw = first ? IN_SMEM : true;
    if constexpr (rem >= 16) {
      constexpr bool last = (rem == 16);
      constexpr bool sync_aw = last ? OUT_SMEM : true;
      // when the last stage does not write smem nobody waits for its reads either
      stage<DIR, 16, Ns, (I > 0 ? I - 1 : 0), first, last, (last && !OUT_SMEM) ? false : sync_bw, sync_aw>(s, tws, u, bar, ld, st,
                                                                                                          first ? bar0 : -1, cnt0);
      if constexpr (!last) from<DIR, I + 1, Ns * 16, IN_SMEM, OUT_SMEM>(s, tws, u, bar, ld, st);
    } else if constexpr (rem > 1) {
      stage<DIR, rem, Ns, (P::N16 > 1 ? P::N16 - 1 : 0), first, true, OUT_SMEM ? sync_bw : false, OUT_SMEM>(s, tws, u, bar, ld, st,
                                                                                                            first ? bar0 : -1, cnt0);
    }
  }

  template <int DIR, bool IN_SMEM, bool OUT_SMEM, class TW, class LoadOp, class StoreOp>
  // bar0 / cnt0: optional wider named barrier (id, thread count) fencing the FIRST stage's reads from its writes
  static __device__ __forceinline__ void run(T2 *__restrict__ s, const TW &tws, int u, int bar, const LoadOp &ld,
                                             const StoreOp &st, int bar0 = -1, int cnt0 = 0) {
    from<DIR, 0, 1, IN_SMEM, OUT_SMEM>(s, tws, u, bar, ld, st, bar0, cnt0);
  }

  // plain in-place transform of s (input already in s and synchronised; output in s, synchronised)
  template <int DIR>
  static __device__ __forceinline__ void run_inplace(T2 *__restrict__ s, const Twiddles &tws, int u, int bar) {
    SmemLoad<T2> ld{s};
    SmemStore<T2> st{s};
    run<DIR, true, true>(s, tws, u, bar, ld, st);
  }
};

}  // namespace oxfft
