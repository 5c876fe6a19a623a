// Runtime plumbing, geometry kernels and cuFFT plan cache of liborphx.so.
#include <math.h>
#include <stdarg.h>

#include "ox_common.cuh"

namespace ox {

static thread_local char g_err[1024] = "";
cudaStream_t g_stream = 0;
long long g_launches = 0;

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int stage_in(const void *src, int where, size_t bytes, DevBuf &scratch, const void **dev) {
  if (where == OX_DEVICE) {
    *dev = src;
    return OX_OK;
  }
  OX_TRY(scratch.ensure(bytes));
  OX_CUDA(cudaMemcpyAsync(scratch.p, src, bytes, cudaMemcpyHostToDevice, g_stream));
  *dev = scratch.p;
  return OX_OK;
}

int stage_out(void *dst, int where, const void *dev_src, size_t bytes) {
  if (where == OX_DEVICE) {
    if (dst != dev_src) OX_CUDA(cudaMemcpyAsync(dst, dev_src, bytes, cudaMemcpyDeviceToDevice, g_stream));
    return OX_OK;
  }
  OX_CUDA(cudaMemcpyAsync(dst, dev_src, bytes, cudaMemcpyDeviceToHost, g_stream));
  OX_CUDA(cudaStreamSynchronize(g_stream));
  return OX_OK;
}

// ---- stage profiler: CUDA events between named marks on the library stream (bench.py's per-kernel times)
static bool g_prof_on = false;
static std::vector<cudaEvent_t> g_prof_ev;
static std::vector<std::string> g_prof_names;

void stage_mark(const char *name) {
  if (!g_prof_on) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, g_stream);
  g_prof_ev.push_back(e);
  g_prof_names.push_back(name);
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (!cached[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace ox

using namespace ox;

extern "C" {
int ox_profile_begin(void) {
  for (auto e : g_prof_ev) cudaEventDestroy(e);
  g_prof_ev.clear();
  g_prof_names.clear();
  g_prof_on = true;
  stage_mark("begin");
  return OX_OK;
}

int ox_profile_end(int *nstages) {
  OX_REQUIRE(nstages, "null pointer");
  g_prof_on = false;
  if (!g_prof_ev.empty()) OX_CUDA(cudaEventSynchronize(g_prof_ev.back()));
  *nstages = g_prof_ev.empty() ? 0 : (int)g_prof_ev.size() - 1;
  return OX_OK;
}

int ox_profile_stage(int i, char *name, size_t len, float *ms) {
  OX_REQUIRE(name && ms && i >= 0 && i + 1 < (int)g_prof_ev.size(), "ox_profile_stage: stage %d out of range", i);
  snprintf(name, len, "%s", g_prof_names[i + 1].c_str());
  OX_CUDA(cudaEventElapsedTime(ms, g_prof_ev[i], g_prof_ev[i + 1]));
  return OX_OK;
}
}

// ---- cuFFT plan cache ------------------------------------------------------------
FFTPlans::~FFTPlans() {
  for (auto &kv : r2c) cufftDestroy(kv.second);
  for (auto &kv : c2r) cufftDestroy(kv.second);
  for (auto &kv : c2c) cufftDestroy(kv.second);
}

int FFTPlans::get(std::map<int, cufftHandle> &cache, cufftType type, int nplanes, cufftHandle *out) {
  auto it = cache.find(nplanes);
  if (it != cache.end()) {
    *out = it->second;
    return OX_OK;
  }
  cufftHandle h;
  OX_CUFFT(cufftCreate(&h));
  OX_CUFFT(cufftSetAutoAllocation(h, 0));
  int n[2] = {ny, nx};
  int nxh = nx / 2 + 1;
  long long n_ll[2] = {ny, nx};
  long long rdist = (long long)ny * nx, cdist = (long long)ny * nxh;
  long long idist, odist;
  long long inembed[2], onembed[2];
  (void)n;
  if (type == CUFFT_D2Z || type == CUFFT_R2C) {
    inembed[0] = ny; inembed[1] = nx; onembed[0] = ny; onembed[1] = nxh;
    idist = rdist; odist = cdist;
  } else if (type == CUFFT_Z2D || type == CUFFT_C2R) {
    inembed[0] = ny; inembed[1] = nxh; onembed[0] = ny; onembed[1] = nx;
    idist = cdist; odist = rdist;
  } else {
    inembed[0] = ny; inembed[1] = nx; onembed[0] = ny; onembed[1] = nx;
    idist = rdist; odist = rdist;
  }
  size_t ws = 0;
  OX_CUFFT(cufftMakePlanMany64(h, 2, n_ll, inembed, 1, idist, onembed, 1, odist, type, nplanes, &ws));
  OX_TRY(work.ensure(ws));
  // a grown work area must be re-attached to the plans created earlier
  for (auto &kv : r2c) OX_CUFFT(cufftSetWorkArea(kv.second, work.p));
  for (auto &kv : c2r) OX_CUFFT(cufftSetWorkArea(kv.second, work.p));
  for (auto &kv : c2c) OX_CUFFT(cufftSetWorkArea(kv.second, work.p));
  OX_CUFFT(cufftSetWorkArea(h, work.p));
  cache[nplanes] = h;
  *out = h;
  return OX_OK;
}

int FFTPlans::exec_r2c(int nplanes, void *in, void *out) {
  cufftHandle h;
  OX_TRY(get(r2c, dtype == OX_F32 ? CUFFT_R2C : CUFFT_D2Z, nplanes, &h));
  OX_CUFFT(cufftSetStream(h, g_stream));
  if (dtype == OX_F32)
    OX_CUFFT(cufftExecR2C(h, (cufftReal *)in, (cufftComplex *)out));
  else
    OX_CUFFT(cufftExecD2Z(h, (cufftDoubleReal *)in, (cufftDoubleComplex *)out));
  g_launches++;
  return OX_OK;
}

int FFTPlans::exec_c2r(int nplanes, void *in, void *out) {
  cufftHandle h;
  OX_TRY(get(c2r, dtype == OX_F32 ? CUFFT_C2R : CUFFT_Z2D, nplanes, &h));
  OX_CUFFT(cufftSetStream(h, g_stream));
  if (dtype == OX_F32)
    OX_CUFFT(cufftExecC2R(h, (cufftComplex *)in, (cufftReal *)out));
  else
    OX_CUFFT(cufftExecZ2D(h, (cufftDoubleComplex *)in, (cufftDoubleReal *)out));
  g_launches++;
  return OX_OK;
}

int FFTPlans::exec_c2c(int nplanes, void *in, void *out, int direction) {
  cufftHandle h;
  OX_TRY(get(c2c, dtype == OX_F32 ? CUFFT_C2C : CUFFT_Z2Z, nplanes, &h));
  OX_CUFFT(cufftSetStream(h, g_stream));
  if (dtype == OX_F32)
    OX_CUFFT(cufftExecC2C(h, (cufftComplex *)in, (cufftComplex *)out, direction));
  else
    OX_CUFFT(cufftExecZ2Z(h, (cufftDoubleComplex *)in, (cufftDoubleComplex *)out, direction));
  g_launches++;
  return OX_OK;
}

// ---- geometry kernels --------------------------------------------------------------
// modl = sqrt(ly^2 + lx^2) with explicitly rounded mul/add/sqrt (no FMA contraction) so
// that the result is bit-identical to numpy's sum(lmap**2,0)**0.5.
__global__ void modlmap_kernel(const double *__restrict__ ly, const double *__restrict__ lx, int ny, int nx,
                               double *__restrict__ out) {
  long long n = (long long)ny * nx;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int iy = (int)(i / nx), ix = (int)(i - (long long)iy * nx);
    double y = ly[iy], x = lx[ix];
    out[i] = __dsqrt_rn(__dadd_rn(__dmul_rn(y, y), __dmul_rn(x, x)));
  }
}

// queb_rotmat: a = sgn*2*atan2(-lx, ly); rot = [[c,-s],[s,c]]
__global__ void rotmat_kernel(const double *__restrict__ ly, const double *__restrict__ lx, int ny, int nx, double sgn,
                              double *__restrict__ out) {
  long long n = (long long)ny * nx;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int iy = (int)(i / nx), ix = (int)(i - (long long)iy * nx);
    double a = sgn * 2.0 * atan2(-lx[ix], ly[iy]);
    double s, c;
    sincos(a, &s, &c);
    out[i] = c;
    out[n + i] = -s;
    out[2 * n + i] = s;
    out[3 * n + i] = c;
  }
}

__global__ void mask_kspace_kernel(const double *__restrict__ ly, const double *__restrict__ lx, int ny, int nx,
                                   double lxcut, double lycut, double lmin, double lmax, int *__restrict__ out) {
  long long n = (long long)ny * nx;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int iy = (int)(i / nx), ix = (int)(i - (long long)iy * nx);
    double y = ly[iy], x = lx[ix];
    double m = __dsqrt_rn(__dadd_rn(__dmul_rn(y, y), __dmul_rn(x, x)));
    int keep = 1;
    if (lmin == lmin && m <= lmin) keep = 0;  // NaN = not given
    if (lmax == lmax && m >= lmax) keep = 0;
    if (lxcut == lxcut && fabs(x) < lxcut) keep = 0;
    if (lycut == lycut && fabs(y) < lycut) keep = 0;
    out[i] = keep;
  }
}

// order-1 interpolation of spec[s][0..nl-1] at l = modl; zero outside [0, nl-1]
__global__ void interp_spec_kernel(const double *__restrict__ ly, const double *__restrict__ lx, int ny, int nx,
                                   const double *__restrict__ spec, int nspec, int nl, double *__restrict__ out) {
  long long n = (long long)ny * nx;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int iy = (int)(i / nx), ix = (int)(i - (long long)iy * nx);
    double y = ly[iy], x = lx[ix];
    double m = __dsqrt_rn(__dadd_rn(__dmul_rn(y, y), __dmul_rn(x, x)));
    double fl = floor(m);
    bool inside = (m >= 0.0) && (m <= (double)(nl - 1));
    long long i0 = inside ? (long long)fl : 0;
    long long i1 = (i0 + 1 < nl) ? i0 + 1 : i0;
    double t = m - fl;
    for (int s = 0; s < nspec; s++) {
      double v = 0.0;
      if (inside) {
        double a = spec[(long long)s * nl + i0], b = spec[(long long)s * nl + i1];
        v = (1.0 - t) * a + t * b;
      }
      out[(long long)s * n + i] = v;
    }
  }
}

__global__ void cast_f32_kernel(const double *__restrict__ in, float *__restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = (float)in[i];
}

static inline int grid_for(long long n, int block) {
  long long want = (n + block - 1) / block;
  long long cap = (long long)ox::sm_count() * 8;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

namespace ox {
int cast_from_f64(const double *src_dev, void *dst_dev, long long n, int dtype) {
  if (dtype == OX_F64) {
    OX_CUDA(cudaMemcpyAsync(dst_dev, src_dev, sizeof(double) * n, cudaMemcpyDeviceToDevice, g_stream));
    return OX_OK;
  }
  cast_f32_kernel<<<grid_for(n, 256), 256, 0, g_stream>>>(src_dev, (float *)dst_dev, n);
  OX_KERNEL_CHECK();
  return OX_OK;
}
}  // namespace ox

// ---- exported: runtime ---------------------------------------------------------------
extern "C" {

int ox_abi_version(void) { return OX_ABI_VERSION; }
const char *ox_last_error(void) { return ox::g_err; }

int ox_device_count(int *n) {
  OX_REQUIRE(n, "null pointer");
  cudaError_t e = cudaGetDeviceCount(n);
  if (e != cudaSuccess) {
    *n = 0;
    set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return OX_ERR_CUDA;
  }
  return OX_OK;
}
int ox_set_device(int dev) {
  OX_CUDA(cudaSetDevice(dev));
  return OX_OK;
}
int ox_get_device(int *dev) {
  OX_CUDA(cudaGetDevice(dev));
  return OX_OK;
}
int ox_device_name(char *buf, size_t len) {
  int dev;
  OX_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  OX_CUDA(cudaGetDeviceProperties(&p, dev));
  snprintf(buf, len, "%s sm_%d%d %d SMs", p.name, p.major, p.minor, p.multiProcessorCount);
  return OX_OK;
}
int ox_synchronize(void) {
  OX_CUDA(cudaStreamSynchronize(g_stream));
  return OX_OK;
}
int ox_set_stream(void *s) {
  g_stream = (cudaStream_t)s;
  return OX_OK;
}
int ox_malloc(void **dptr, size_t bytes) {
  OX_REQUIRE(dptr, "null pointer");
  cudaError_t e = cudaMalloc(dptr, bytes);
  if (e != cudaSuccess) {
    set_error("cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    return OX_ERR_NOMEM;
  }
  return OX_OK;
}
int ox_free(void *dptr) {
  OX_CUDA(cudaFree(dptr));
  return OX_OK;
}
// Stream-ordered pool for the device-resident maps of the per-call API (enmap.devmap): one cudaMalloc per map
// returned would cost more than the kernels of a 2048^2 map; the pool keeps freed blocks (release threshold =
// never) and hands them out again without a device synchronisation.
int ox_malloc_pooled(void **dptr, size_t bytes) {
  OX_REQUIRE(dptr, "null pointer");
  static int pooled_dev = -1;
  int dev = 0;
  OX_CUDA(cudaGetDevice(&dev));
  if (pooled_dev != dev) {
    cudaMemPool_t pool;
    OX_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
    unsigned long long keep = ~0ULL;
    OX_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    pooled_dev = dev;
  }
  cudaError_t e = cudaMallocAsync(dptr, bytes ? bytes : 1, g_stream);
  if (e != cudaSuccess) {
    set_error("cudaMallocAsync of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    cudaGetLastError();
    return OX_ERR_NOMEM;
  }
  return OX_OK;
}
int ox_free_pooled(void *dptr) {
  if (dptr) OX_CUDA(cudaFreeAsync(dptr, g_stream));
  return OX_OK;
}

}  // extern "C"
namespace {
// out[i] = a[i] (op) b[i % nb]  (or the scalar when b == null); complex a / out with a real b
template <typename TA, typename TB>
__global__ void map_op_kernel(int op, const TA *__restrict__ a, const TB *__restrict__ b, double scalar, long long n,
                              long long nb, TA *__restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const TB y = b ? b[nb == n ? i : i % nb] : (TB)scalar;
    TA x = a[i];
    if constexpr (sizeof(TA) == sizeof(TB)) {
      switch (op) {
        case 0: x = x * y; break;
        case 1: x = x + y; break;
        case 2: x = x - y; break;
        case 3: x = x / y; break;
        case 4: x = y - x; break;
        default: x = y / x; break;
      }
    } else {  // complex (op) real: numpy's semantics for mul/div; add/sub touch the real part only
      switch (op) {
        case 0: x.x = x.x * y; x.y = x.y * y; break;
        case 1: x.x = x.x + y; break;
        case 2: x.x = x.x - y; break;
        case 3: x.x = x.x / y; x.y = x.y / y; break;
        default: x.x = y - x.x; x.y = -x.y; break;
      }
    }
    out[i] = x;
  }
}
}  // namespace
extern "C" {

// Elementwise arithmetic on device-resident maps (what `emap * taper`, `p2d / w2`, `map + noise` of the
// reference's call sequences do in numpy; maps.py:1351, lensing.py:519).  kind: 0 f64, 1 f32 (a, b, out real),
// 2 complex128 a/out with float64 b, 3 complex64 a/out with float32 b.  op: 0 a*b, 1 a+b, 2 a-b, 3 a/b, 4 b-a,
// 5 b/a (real kinds only).  b == NULL: the scalar.  b holds nb elements and repeats over a's n (nb divides n).
int ox_map_op(int op, const void *a, const void *b, double scalar, long long n, long long nb, int kind, void *out) {
  OX_REQUIRE(a && out && n >= 0, "ox_map_op: null pointer");
  OX_REQUIRE(op >= 0 && op <= 5 && kind >= 0 && kind <= 3, "ox_map_op: op=%d kind=%d", op, kind);
  OX_REQUIRE(!(kind >= 2 && op == 5), "ox_map_op: real / complex is not provided");
  if (b) OX_REQUIRE(nb > 0 && n % nb == 0, "ox_map_op: operand of %lld elements does not tile %lld", nb, n);
  if (n == 0) return OX_OK;
  const int threads = 256;
  long long want = (n + threads - 1) / threads;
  const int blocks = (int)(want < (long long)sm_count() * 16 ? want : (long long)sm_count() * 16);
  switch (kind) {
    case 0: map_op_kernel<double, double><<<blocks, threads, 0, g_stream>>>(op, (const double *)a, (const double *)b, scalar, n, nb, (double *)out); break;
    case 1: map_op_kernel<float, float><<<blocks, threads, 0, g_stream>>>(op, (const float *)a, (const float *)b, scalar, n, nb, (float *)out); break;
    case 2: map_op_kernel<double2, double><<<blocks, threads, 0, g_stream>>>(op, (const double2 *)a, (const double *)b, scalar, n, nb, (double2 *)out); break;
    default: map_op_kernel<float2, float><<<blocks, threads, 0, g_stream>>>(op, (const float2 *)a, (const float *)b, scalar, n, nb, (float2 *)out); break;
  }
  OX_KERNEL_CHECK();
  return OX_OK;
}

int ox_memset(void *dptr, int value, size_t bytes) {
  OX_CUDA(cudaMemsetAsync(dptr, value, bytes, g_stream));
  return OX_OK;
}
int ox_host_alloc(void **hptr, size_t bytes) {
  OX_CUDA(cudaHostAlloc(hptr, bytes, cudaHostAllocDefault));
  return OX_OK;
}
int ox_host_free(void *hptr) {
  OX_CUDA(cudaFreeHost(hptr));
  return OX_OK;
}
int ox_memcpy_h2d(void *dst, const void *src, size_t bytes) {
  OX_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g_stream));
  return OX_OK;
}
int ox_memcpy_d2h(void *dst, const void *src, size_t bytes) {
  OX_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_stream));
  OX_CUDA(cudaStreamSynchronize(g_stream));
  return OX_OK;
}
int ox_memcpy_d2d(void *dst, const void *src, size_t bytes) {
  OX_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, g_stream));
  return OX_OK;
}
int ox_mem_info(size_t *free_bytes, size_t *total_bytes) {
  OX_CUDA(cudaMemGetInfo(free_bytes, total_bytes));
  return OX_OK;
}

struct OxTimer {
  cudaEvent_t a, b;
};
int ox_timer_create(void **t) {
  OxTimer *x = new OxTimer;
  OX_CUDA(cudaEventCreate(&x->a));
  OX_CUDA(cudaEventCreate(&x->b));
  *t = x;
  return OX_OK;
}
int ox_timer_start(void *t) {
  OX_CUDA(cudaEventRecord(((OxTimer *)t)->a, g_stream));
  return OX_OK;
}
int ox_timer_stop(void *t) {
  OX_CUDA(cudaEventRecord(((OxTimer *)t)->b, g_stream));
  return OX_OK;
}
int ox_timer_elapsed_ms(void *t, float *ms) {
  OxTimer *x = (OxTimer *)t;
  OX_CUDA(cudaEventSynchronize(x->b));
  OX_CUDA(cudaEventElapsedTime(ms, x->a, x->b));
  return OX_OK;
}
int ox_timer_destroy(void *t) {
  OxTimer *x = (OxTimer *)t;
  cudaEventDestroy(x->a);
  cudaEventDestroy(x->b);
  delete x;
  return OX_OK;
}
int ox_launch_count(long long *n) {
  *n = g_launches;
  return OX_OK;
}

static ox::DevBuf *g_flush = nullptr;
int ox_flush_l2(void) {
  const size_t bytes = 256u << 20;  // > 126 MB L2
  if (!g_flush) g_flush = new ox::DevBuf;
  OX_TRY(g_flush->ensure(bytes));
  OX_CUDA(cudaMemsetAsync(g_flush->p, 0, bytes, g_stream));
  return OX_OK;
}

// ---- exported: geometry ------------------------------------------------------------------
int ox_geometry_create(int ny, int nx, const double *ly, const double *lx, double area, ox_geometry **out) {
  OX_REQUIRE(ny > 0 && nx > 0 && ly && lx && out, "ox_geometry_create: bad arguments (ny=%d nx=%d)", ny, nx);
  ox_geometry *g = new ox_geometry;
  g->ny = ny;
  g->nx = nx;
  g->nxh = nx / 2 + 1;
  g->area = area;
  g->h_ly.assign(ly, ly + ny);
  g->h_lx.assign(lx, lx + nx);
  int s;
  if ((s = g->ly.ensure(sizeof(double) * ny)) != OX_OK || (s = g->lx.ensure(sizeof(double) * nx)) != OX_OK) {
    delete g;
    return s;
  }
  cudaError_t e = cudaMemcpyAsync(g->ly.p, ly, sizeof(double) * ny, cudaMemcpyHostToDevice, g_stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(g->lx.p, lx, sizeof(double) * nx, cudaMemcpyHostToDevice, g_stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(g_stream);
  if (e != cudaSuccess) {
    set_error("ox_geometry_create: %s", cudaGetErrorString(e));
    delete g;
    return OX_ERR_CUDA;
  }
  *out = g;
  return OX_OK;
}

int ox_geometry_destroy(ox_geometry *g) {
  delete g;
  return OX_OK;
}

int ox_geometry_modlmap(ox_geometry *g, double *out, int where) {
  OX_REQUIRE(g && out, "null pointer");
  long long n = (long long)g->ny * g->nx;
  ox::DevBuf tmp;
  double *d = out;
  if (where == OX_HOST) {
    OX_TRY(tmp.ensure(n * sizeof(double)));
    d = tmp.as<double>();
  }
  modlmap_kernel<<<grid_for(n, 256), 256, 0, g_stream>>>(g->ly.as<double>(), g->lx.as<double>(), g->ny, g->nx, d);
  OX_KERNEL_CHECK();
  if (where == OX_HOST) OX_TRY(stage_out(out, OX_HOST, d, n * sizeof(double)));
  return OX_OK;
}

int ox_geometry_rotmat(ox_geometry *g, int flags, double *out, int where) {
  OX_REQUIRE(g && out, "null pointer");
  long long n = (long long)g->ny * g->nx;
  ox::DevBuf tmp;
  double *d = out;
  if (where == OX_HOST) {
    OX_TRY(tmp.ensure(4 * n * sizeof(double)));
    d = tmp.as<double>();
  }
  rotmat_kernel<<<grid_for(n, 256), 256, 0, g_stream>>>(g->ly.as<double>(), g->lx.as<double>(), g->ny, g->nx,
                                                       (flags & OX_FLAG_IAU) ? 1.0 : -1.0, d);
  OX_KERNEL_CHECK();
  if (where == OX_HOST) OX_TRY(stage_out(out, OX_HOST, d, 4 * n * sizeof(double)));
  return OX_OK;
}

int ox_geometry_mask_kspace(ox_geometry *g, double lxcut, double lycut, double lmin, double lmax, int *out, int where) {
  OX_REQUIRE(g && out, "null pointer");
  long long n = (long long)g->ny * g->nx;
  ox::DevBuf tmp;
  int *d = out;
  if (where == OX_HOST) {
    OX_TRY(tmp.ensure(n * sizeof(int)));
    d = tmp.as<int>();
  }
  mask_kspace_kernel<<<grid_for(n, 256), 256, 0, g_stream>>>(g->ly.as<double>(), g->lx.as<double>(), g->ny, g->nx, lxcut,
                                                            lycut, lmin, lmax, d);
  OX_KERNEL_CHECK();
  if (where == OX_HOST) OX_TRY(stage_out(out, OX_HOST, d, n * sizeof(int)));
  return OX_OK;
}

int ox_geometry_interp_spec(ox_geometry *g, const double *spec, int nspec, int nl, double *out, int where) {
  OX_REQUIRE(g && spec && out && nspec > 0 && nl > 1, "ox_geometry_interp_spec: bad arguments");
  long long n = (long long)g->ny * g->nx;
  ox::DevBuf dspec, tmp;
  OX_TRY(dspec.ensure(sizeof(double) * nspec * nl));
  OX_CUDA(cudaMemcpyAsync(dspec.p, spec, sizeof(double) * nspec * nl, cudaMemcpyHostToDevice, g_stream));
  double *d = out;
  if (where == OX_HOST) {
    OX_TRY(tmp.ensure(nspec * n * sizeof(double)));
    d = tmp.as<double>();
  }
  interp_spec_kernel<<<grid_for(n, 256), 256, 0, g_stream>>>(g->ly.as<double>(), g->lx.as<double>(), g->ny, g->nx,
                                                            dspec.as<double>(), nspec, nl, d);
  OX_KERNEL_CHECK();
  if (where == OX_HOST)
    OX_TRY(stage_out(out, OX_HOST, d, nspec * n * sizeof(double)));
  else
    OX_CUDA(cudaStreamSynchronize(g_stream));  // dspec is freed on return
  return OX_OK;
}

}  // extern "C"
