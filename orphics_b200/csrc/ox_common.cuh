// Shared internals of liborphx.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/orphx.h"

namespace ox {

// ---- error plumbing ---------------------------------------------------------
void set_error(const char *fmt, ...);
extern cudaStream_t g_stream;
extern long long g_launches;

#define OX_CUDA(call)                                                                         \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      ox::set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, __LINE__,   \
                    cudaGetErrorString(e_));                                                  \
      return OX_ERR_CUDA;                                                                     \
    }                                                                                         \
  } while (0)

#define OX_CUFFT(call)                                                                        \
  do {                                                                                        \
    cufftResult r_ = (call);                                                                  \
    if (r_ != CUFFT_SUCCESS) {                                                                \
      ox::set_error("cuFFT error %d at %s:%d", (int)r_, __FILE__, __LINE__);                  \
      return OX_ERR_CUFFT;                                                                    \
    }                                                                                         \
  } while (0)

#define OX_REQUIRE(cond, ...)                                                                 \
  do {                                                                                        \
    if (!(cond)) {                                                                            \
      ox::set_error(__VA_ARGS__);                                                             \
      return OX_ERR_INVALID;                                                                  \
    }                                                                                         \
  } while (0)

#define OX_TRY(call)                                                                          \
  do {                                                                                        \
    int s_ = (call);                                                                          \
    if (s_ != OX_OK) return s_;                                                               \
  } while (0)

#define OX_KERNEL_CHECK()                                                                     \
  do {                                                                                        \
    ox::g_launches++;                                                                         \
    OX_CUDA(cudaGetLastError());                                                              \
  } while (0)

// ---- device buffer with RAII -------------------------------------------------
struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
  DevBuf() {}
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  int ensure(size_t n) {
    if (n <= bytes) return OX_OK;
    release();
    cudaError_t e = cudaMalloc(&p, n);
    if (e != cudaSuccess) {
      set_error("cudaMalloc of %zu bytes failed: %s", n, cudaGetErrorString(e));
      p = nullptr;
      return OX_ERR_NOMEM;
    }
    bytes = n;
    return OX_OK;
  }
  template <typename T>
  T *as() const {
    return reinterpret_cast<T *>(p);
  }
};

// stage a caller buffer on the device if it is host memory; returns the device pointer
int stage_in(const void *src, int where, size_t bytes, DevBuf &scratch, const void **dev);
int stage_out(void *dst, int where, const void *dev_src, size_t bytes);

template <typename T>
struct Vec2;
template <>
struct Vec2<double> {
  typedef double2 type;
};
template <>
struct Vec2<float> {
  typedef float2 type;
};

static inline size_t elem_size(int dtype) { return dtype == OX_F32 ? 4 : 8; }

// number of SMs of the current device (cached)
int sm_count();
// records a CUDA event named after the stage that just finished (no-op unless ox_profile_begin is active)
void stage_mark(const char *name);

}  // namespace ox

// ---- handle definitions --------------------------------------------------------
struct ox_geometry {
  int ny = 0, nx = 0, nxh = 0;
  double area = 0;
  ox::DevBuf ly, lx;  // double[ny], double[nx]
  std::vector<double> h_ly, h_lx;
};

struct ox_binner {
  long long n = 0;
  int nedges = 0, nslots = 0;
  ox::DevBuf edges;   // double[nedges]
  ox::DevBuf idx;     // uint16[n]  full-plane slot index
  ox::DevBuf counts;  // int64[nslots]
  std::vector<long long> h_counts;
  // half-plane (geometry-derived binners only)
  bool has_half = false;
  int ny = 0, nx = 0, nxh = 0;
  ox::DevBuf idxh;      // uint16[ny*nxh]: slot | (weight==2 ? 0x8000 : 0)
  ox::DevBuf countf;  // double[nslots]: the integer slot counts as float64 (divisor of the bandpower sums)
  ox::DevBuf scratch, partial, stage;
};

struct FFTPlans {
  // cuFFT plans keyed by number of planes, sharing one work area
  int ny = 0, nx = 0, dtype = 0;
  std::map<int, cufftHandle> r2c, c2r, c2c;
  ox::DevBuf work;
  ~FFTPlans();
  int get(std::map<int, cufftHandle> &cache, cufftType type, int nplanes, cufftHandle *out);
  int exec_r2c(int nplanes, void *in, void *out);
  int exec_c2r(int nplanes, void *in, void *out);
  int exec_c2c(int nplanes, void *in, void *out, int direction);
};

struct ox_simplan {
  ox_geometry *g = nullptr;
  int ncomp = 1, dtype = OX_F64, max_batch = 1;
  ox::DevBuf covsqrt;  // T[ncomp][ncomp][ny][nx]
  FFTPlans fft;
  ox::DevBuf kh;     // half-plane complex [max_batch][ncomp][ny][nxh]
  ox::DevBuf maps;   // real [max_batch][ncomp][ny][nx]
  ox::DevBuf noise;  // staging of host noise
  ox::DevBuf seeds;  // int64[max_batch]
  ox::DevBuf stage;
};

struct ox_powerplan {
  ox_geometry *g = nullptr;
  int ncomp = 1, dtype = OX_F64, max_batch = 1;
  double normfact = 0;
  FFTPlans fft;
  ox::DevBuf in1, in2;    // staged real maps
  ox::DevBuf kh1, kh2;    // half-plane complex
  ox::DevBuf full1, full2, p2d;
  ox::DevBuf partial, bp, window;
};

// state of the hand-written fused path (ox_fused.cu)
struct FusedState {
  bool ready = false, cov_symmetric = false;
  int tw_len = 0;
  ox::DevBuf tw;      // exp(-2 pi i j / tw_len)
  ox::DevBuf covT;    // covsqrt transposed [nc][nc][nx][ny]
  ox::DevBuf idxT;    // half-plane slot index transposed [nx/2+1][ny]
  ox::DevBuf Ha, Hb;  // transposed half planes [max_batch][nc][nx/2+1][ny]
};

struct ox_pipeline {
  int path = 0;       // 1 = cuFFT passes, 2 = fused hand-written FFT kernels
  FusedState fused;
  ox_simplan *s = nullptr;
  ox_powerplan *p = nullptr;
  ox_binner *b = nullptr;
  ox::DevBuf window;  // T[ny][nx] or empty
  bool has_window = false;
  ox::DevBuf win_x, win_y;  // float64 profiles of a separable window (window == outer(win_y, win_x) exactly), else empty
  bool win_separable = false;
  bool maps_valid = false;  // the last run left its real-space maps in s->maps
  int nspec = 1, nbins = 0, dim = 0;
  ox::DevBuf partial, bp;            // partial sums, bandpowers [max_batch][nspec][nbins]
  // Statistics triple packed for ONE all-reduce (stats.py:1215-1217): float64 [N | SUM[dim] | CROSS[dim][dim]],
  // N kept as an exactly representable integer
  ox::DevBuf stat;
};

// ---- cross-file internal entry points -------------------------------------------
namespace ox {
// sim: fill plan->kh for nsim sims from the chosen noise source (no FFT)
int sim_fill_half(ox_simplan *p, const long long *seeds_host, int nsim, int noise_mode, const double *noise,
                  int noise_where, int flags);
// sim: upload seeds / stage host noise for nsim sims
int sim_stage_inputs(ox_simplan *p, const long long *seeds_host, int nsim, int noise_mode, const double *noise,
                     int noise_where, const double **noise_dev);
// fused path (ox_fused.cu)
bool fused_supported(int ny, int nx, int ncomp, int dtype);
// twiddle table exp(-2 pi i j / len), j < len, in the plan dtype (long-double accurate)
int fused_make_twiddles(int len, int dtype, DevBuf &buf);
int fused_prepare(ox_simplan *s, ox_binner *b, FusedState &fs);
int fused_run(ox_pipeline *pl, int nsim, int noise_mode, const double *noise_dev, int flags, bool keep_maps,
              cudaEvent_t *ev);
// sim: kh -> real maps in p->maps
int sim_to_maps(ox_simplan *p, int nsim);
// power+bin from half-plane Fourier arrays on device -> bandpowers (device) [nbatch][nspec][nbins]
int power_bin_half(ox_geometry *g, ox_binner *b, int dtype, int ncomp, const void *kh1, const void *kh2, int nbatch,
                   int flags, double normfact, DevBuf &partial, double *bp_dev);
// dst[n] (dtype) = src[n] (float64), both on the device
int cast_from_f64(const double *src_dev, void *dst_dev, long long n, int dtype);
// in-place NCCL sum of count float64 values on the library stream (ox_comm.cu)
int comm_allreduce_f64(ox_comm *c, double *buf_dev, long long count);
// maps[n] *= window[npix] (in place, broadcast over planes)
int apply_window(int dtype, void *maps, const void *window, long long npix, long long nplanes);
}  // namespace ox
