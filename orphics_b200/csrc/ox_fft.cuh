// Block-level FFT engine for the fused sim -> FFT -> power -> bin kernels (ox_fused.cu).
//
// A power-of-two complex FFT of length L lives in shared memory; L/8 "units" of 8
// elements are processed per stage by a Stockham autosort radix-8 (then 4 or 2) pass:
//   read 8 strided elements -> twiddle -> butterfly in registers -> barrier ->
//   write to the autosorted positions -> barrier.
// Reads of a stage are contiguous across lanes; the stride-R writes of the first stage
// would be 8..32-way bank conflicts, so the buffer is padded by one element every 8
// (pad(e) = e + e/8), which makes every access pattern of every stage conflict-free for
// both 16-byte (double2) and 8-byte (float2) elements.
//
// Twiddles: a thread works on the same butterflies in every transform it takes part in,
// so its first-order twiddle w = exp(-2 pi i k / (Ns R)) of every stage is loaded ONCE from
// a global table (tw[j] = exp(-2 pi i j / LT), built on the host in long double) into
// registers (Twiddles::init); w^2..w^7 are formed by complex multiplications in the stage.
// (Loading all seven per butterfly from the table made the kernels L1-bound: ncu showed
// 4.7x more L1 sectors for twiddles than for data.)  DIR = -1 forward, +1 backward.
#pragma once
#include <cuda_runtime.h>

namespace oxfft {

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

__host__ __device__ constexpr int pad(int e) { return e + (e >> 3); }
// elements to allocate for one padded length-L buffer that may also hold index L (the
// Nyquist element of a real transform); the +2 makes consecutive buffers start 32 bytes
// apart modulo 128 for double2 so that equal indices of neighbouring rows hit different banks
__host__ __device__ constexpr int padded_size(int L) { return L + (L >> 3) + 2; }

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

template <int DIR, typename T2>
__device__ __forceinline__ void bfly2(T2 (&v)[8], int o) {
  T2 a = v[o], b = v[o + 1];
  v[o] = cadd(a, b);
  v[o + 1] = csub(a, b);
}

// radix-4 DFT of v[o..o+3] (natural order in and out)
template <int DIR, typename T2>
__device__ __forceinline__ void bfly4(T2 (&v)[8], int o) {
  T2 a0 = cadd(v[o], v[o + 2]), a2 = csub(v[o], v[o + 2]);
  T2 a1 = cadd(v[o + 1], v[o + 3]), a3 = mul_i<DIR>(csub(v[o + 1], v[o + 3]));
  v[o] = cadd(a0, a1);
  v[o + 1] = cadd(a2, a3);
  v[o + 2] = csub(a0, a1);
  v[o + 3] = csub(a2, a3);
}

// radix-8 DFT of v[0..7] (natural order in and out)
template <int DIR, typename T2>
__device__ __forceinline__ void bfly8(T2 (&v)[8]) {
  typedef decltype(v[0].x) T;
  const T h = (T)0.70710678118654752440;
  T2 a0 = cadd(v[0], v[4]), a4 = csub(v[0], v[4]);
  T2 a1 = cadd(v[1], v[5]), a5 = csub(v[1], v[5]);
  T2 a2 = cadd(v[2], v[6]), a6 = csub(v[2], v[6]);
  T2 a3 = cadd(v[3], v[7]), a7 = csub(v[3], v[7]);
  // a5 *= e^{DIR i pi/4}, a6 *= e^{DIR i pi/2}, a7 *= e^{DIR 3 i pi/4}
  T2 t;
  if (DIR > 0) {
    t.x = h * (a5.x - a5.y); t.y = h * (a5.x + a5.y); a5 = t;
    t.x = -h * (a7.x + a7.y); t.y = h * (a7.x - a7.y); a7 = t;
  } else {
    t.x = h * (a5.x + a5.y); t.y = h * (a5.y - a5.x); a5 = t;
    t.x = h * (a7.y - a7.x); t.y = -h * (a7.x + a7.y); a7 = t;
  }
  a6 = mul_i<DIR>(a6);
  T2 b0 = cadd(a0, a2), b2 = csub(a0, a2), b1 = cadd(a1, a3), b3 = mul_i<DIR>(csub(a1, a3));
  T2 b4 = cadd(a4, a6), b6 = csub(a4, a6), b5 = cadd(a5, a7), b7 = mul_i<DIR>(csub(a5, a7));
  v[0] = cadd(b0, b1); v[4] = csub(b0, b1);
  v[2] = cadd(b2, b3); v[6] = csub(b2, b3);
  v[1] = cadd(b4, b5); v[5] = csub(b4, b5);
  v[3] = cadd(b6, b7); v[7] = csub(b6, b7);
}

// barrier over the threads of one transform: id 0 = the whole CTA (__syncthreads), id 1..15 = a
// named barrier shared by the `count` threads working on this transform, so that independent
// transforms in one CTA do not wait for each other
__device__ __forceinline__ void fft_sync(int bar, int count) {
  if (bar == 0) __syncthreads();
  else asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(count) : "memory");
}

// compile-time stage plan of a length-L transform: radix 8 while possible, then 4 or 2
template <int L>
struct Plan {
  // number of first-order twiddles a unit needs over all stages with Ns > 1
  template <int Ns>
  static constexpr int count() {
    constexpr int rem = L / Ns;
    if constexpr (rem >= 8) return (Ns > 1 ? 1 : 0) + count<Ns * 8>();
    else if constexpr (rem == 4) return 2;
    else if constexpr (rem == 2) return 4;
    else return 0;
  }
  static constexpr int NW = count<1>();
};

// In-place shared-memory FFT of length L executed by NT = L/8/BPT threads per transform
// (thread index t in [0,NT)); every thread of the CTA must call run() (it contains
// __syncthreads()).  The caller fills s[pad(e)] (natural order), synchronises, calls run();
// on return (after a final barrier) s[pad(f)] holds the transform in natural order.
template <typename T, int L, int BPT>
struct BlockFFT {
  typedef typename V2<T>::type T2;
  static constexpr int UNITS = L / 8;
  static constexpr int NT = UNITS / BPT;
  static constexpr int NW = Plan<L>::NW * BPT;
  static_assert(L >= 8 && (L & (L - 1)) == 0, "L must be a power of two >= 8");
  static_assert(UNITS % BPT == 0, "BPT must divide L/8");

  // per-thread first-order (forward) twiddles of every stage, held in registers
  struct Twiddles {
    T2 w[NW > 0 ? NW : 1];

    template <int Ns, int OFF>
    __device__ __forceinline__ void fill(const T2 *__restrict__ tw, int tw_stride, int t) {
      constexpr int rem = L / Ns;
      constexpr int R = rem >= 8 ? 8 : rem;
      constexpr int Q = 8 / R;
      if constexpr (rem >= 2) {
        if constexpr (Ns > 1) {
#pragma unroll
          for (int b = 0; b < BPT; b++)
#pragma unroll
            for (int q = 0; q < Q; q++) {
              const int j = t + b * NT + q * UNITS;
              const int k = j & (Ns - 1);
              w[OFF + b * Q + q] = tw[k * (L / (Ns * R)) * tw_stride];
            }
        }
        // another stage follows only after a radix-8 stage that leaves a remainder
        if constexpr (rem >= 16) fill<Ns * 8, OFF + (Ns > 1 ? BPT : 0)>(tw, tw_stride, t);
      }
    }
    // tw: table exp(-2 pi i j / (L*tw_stride))
    __device__ __forceinline__ void init(const T2 *__restrict__ tw, int tw_stride, int t) {
      fill<1, 0>(tw, tw_stride, t);
    }
  };

  // one Stockham stage of radix R with Ns = product of the previous radices
  template <int DIR, int R, int Ns, int OFF>
  static __device__ __forceinline__ void stage(T2 *__restrict__ s, const Twiddles &tws, int t, int bar) {
    constexpr int Q = 8 / R;  // radix-R butterflies per unit
    T2 v[BPT][8];
#pragma unroll
    for (int b = 0; b < BPT; b++) {
      const int u = t + b * NT;
#pragma unroll
      for (int q = 0; q < Q; q++) {
        const int j = u + q * UNITS;
#pragma unroll
        for (int r = 0; r < R; r++) v[b][q * R + r] = s[pad(j + r * (L / R))];
        if (Ns > 1) {
          T2 w1 = tws.w[OFF + b * Q + q];
          if (DIR > 0) w1.y = -w1.y;
          T2 *x = &v[b][q * R];
          x[1] = cmul(x[1], w1);
          if (R >= 4) {
            T2 w2 = cmul(w1, w1), w3 = cmul(w2, w1);
            x[2] = cmul(x[2], w2);
            x[3] = cmul(x[3], w3);
            if (R == 8) {
              T2 w4 = cmul(w2, w2);
              x[4] = cmul(x[4], w4);
              x[5] = cmul(x[5], cmul(w4, w1));
              x[6] = cmul(x[6], cmul(w3, w3));
              x[7] = cmul(x[7], cmul(w4, w3));
            }
          }
        }
      }
      if (R == 8) bfly8<DIR>(v[b]);
      if (R == 4) { bfly4<DIR>(v[b], 0); bfly4<DIR>(v[b], 4); }
      if (R == 2) { bfly2<DIR>(v[b], 0); bfly2<DIR>(v[b], 2); bfly2<DIR>(v[b], 4); bfly2<DIR>(v[b], 6); }
    }
    fft_sync(bar, NT);
#pragma unroll
    for (int b = 0; b < BPT; b++) {
      const int u = t + b * NT;
#pragma unroll
      for (int q = 0; q < Q; q++) {
        const int j = u + q * UNITS;
        const int k = j & (Ns - 1);
        const int d = (j - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; r++) s[pad(d + r * Ns)] = v[b][q * R + r];
      }
    }
    fft_sync(bar, NT);
  }

  template <int DIR, int Ns, int OFF>
  static __device__ __forceinline__ void from(T2 *__restrict__ s, const Twiddles &tws, int t, int bar) {
    constexpr int rem = L / Ns;
    if constexpr (rem >= 8) {
      stage<DIR, 8, Ns, OFF>(s, tws, t, bar);
      from<DIR, Ns * 8, OFF + (Ns > 1 ? BPT : 0)>(s, tws, t, bar);
    } else if constexpr (rem == 4) {
      stage<DIR, 4, Ns, OFF>(s, tws, t, bar);
    } else if constexpr (rem == 2) {
      stage<DIR, 2, Ns, OFF>(s, tws, t, bar);
    }
  }

  // bar: see fft_sync.  The caller must have made its writes to s visible to the NT threads
  // of this transform (same barrier) before calling.
  template <int DIR>
  static __device__ __forceinline__ void run(T2 *__restrict__ s, const Twiddles &tws, int t, int bar = 0) {
    from<DIR, 1, 0>(s, tws, t, bar);
  }
};

}  // namespace oxfft
