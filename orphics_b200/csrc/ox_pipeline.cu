// The north-star path as one call: MapGen.get_map -> (x taper) -> FourierCalc.power2d ->
// bin2D.bin for a batch of seeds, plus the Statistics triple (N, SUM x, SUM x x^T;
// stats.py:1085-1090) accumulated on the device for the final all-reduce.
//
// HBM passes per T-only map (s = bytes per real, N = Ny*Nx):
//   sim_fill writes k_h (sN) -> cuFFT Z2D (2 passes) -> window RMW (2sN, optional) ->
//   cuFFT D2Z (2 passes) -> power_bin reads k_h (sN) + uint16 slot index (N).
#include <stdlib.h>

#include "ox_common.cuh"

using namespace ox;

namespace {

// SUM[d] += sum_m x[m][d];  CROSS[i][j] += sum_m x[m][i] x[m][j]   (fixed m order)
__global__ void stats_accumulate_kernel(const double *__restrict__ x, int nsim, int dim, double *__restrict__ packed) {
  double *n = packed, *sum = packed + 1, *cross = packed + 1 + dim;
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long total = (long long)dim * dim;
  if (t == 0) *n += (double)nsim;
  if (t < dim) {
    double acc = sum[t];
    for (int m = 0; m < nsim; m++) acc += x[(long long)m * dim + t];
    sum[t] = acc;
  }
  if (t < total) {
    int i = (int)(t / dim), j = (int)(t - (long long)i * dim);
    double acc = cross[t];
    for (int m = 0; m < nsim; m++) acc += x[(long long)m * dim + i] * x[(long long)m * dim + j];
    cross[t] = acc;
  }
}

}  // namespace

extern "C" {

int ox_pipeline_create(ox_simplan *s, ox_powerplan *p, ox_binner *b, const double *window, int window_where,
                       ox_pipeline **out) {
  OX_REQUIRE(s && p && b && out, "ox_pipeline_create: null pointer");
  OX_REQUIRE(s->g == p->g, "MapGen and FourierCalc must share one geometry");
  OX_REQUIRE(s->dtype == p->dtype, "MapGen and FourierCalc must use the same dtype");
  OX_REQUIRE(s->ncomp == p->ncomp, "MapGen ncomp=%d but FourierCalc ncomp=%d", s->ncomp, p->ncomp);
  OX_REQUIRE(s->ncomp <= 3, "pipeline supports ncomp 1..3");
  OX_REQUIRE(b->has_half, "the bin2D must be built with ox_binner_create_geom from a Hermitian-symmetric geometry");
  OX_REQUIRE(b->ny == s->g->ny && b->nx == s->g->nx, "binner/geometry shape mismatch");
  ox_pipeline *pl = new ox_pipeline;
  pl->s = s;
  pl->p = p;
  pl->b = b;
  pl->nspec = s->ncomp * (s->ncomp + 1) / 2;
  pl->nbins = b->nslots - 2;
  pl->dim = pl->nspec * pl->nbins;
  auto fail = [&](int st) { delete pl; return st; };
  int st;
  if (window) {
    size_t n = (size_t)s->g->ny * s->g->nx;
    ox::DevBuf tmp;
    const void *d;
    if ((st = stage_in(window, window_where, sizeof(double) * n, tmp, &d)) != OX_OK) return fail(st);
    if ((st = pl->window.ensure(elem_size(s->dtype) * n)) != OX_OK) return fail(st);
    if ((st = cast_from_f64((const double *)d, pl->window.p, (long long)n, s->dtype)) != OX_OK) return fail(st);
    if (cudaStreamSynchronize(g_stream) != cudaSuccess) { set_error("window upload failed"); return fail(OX_ERR_CUDA); }
    pl->has_window = true;
    // separable?  The reference's taper (get_taper / cosine_window, maps.py:1873-1920) is an outer product of two
    // 1-D profiles.  Accept the fast path only if window[iy][ix] == fl(wy[iy] * wx[ix]) for EVERY pixel, with the
    // profiles read off the row and column through a pixel where the window is exactly 1.
    {
      const int ny = s->g->ny, nx = s->g->nx;
      std::vector<double> hw;
      const double *w = window;
      if (window_where != OX_HOST) {
        hw.resize(n);
        if (cudaMemcpy(hw.data(), window, sizeof(double) * n, cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("window download failed"); return fail(OX_ERR_CUDA); }
        w = hw.data();
      }
      const int r0 = ny / 2, c0 = nx / 2;
      bool sep = !(getenv("ORPHX_WINDOW_SEPARABLE") && getenv("ORPHX_WINDOW_SEPARABLE")[0] == '0') && w[(size_t)r0 * nx + c0] == 1.0;
      std::vector<double> wy(ny), wx(nx);
      if (sep) {
        for (int iy = 0; iy < ny; iy++) wy[iy] = w[(size_t)iy * nx + c0];
        for (int ix = 0; ix < nx; ix++) wx[ix] = w[(size_t)r0 * nx + ix];
        for (int iy = 0; iy < ny && sep; iy++) {
          const double *row = w + (size_t)iy * nx;
          const double y = wy[iy];
          for (int ix = 0; ix < nx; ix++)
            if (row[ix] != y * wx[ix]) { sep = false; break; }
        }
      }
      if (sep) {
        if ((st = pl->win_x.ensure(sizeof(double) * nx)) != OX_OK) return fail(st);
        if ((st = pl->win_y.ensure(sizeof(double) * ny)) != OX_OK) return fail(st);
        if (cudaMemcpy(pl->win_x.p, wx.data(), sizeof(double) * nx, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(pl->win_y.p, wy.data(), sizeof(double) * ny, cudaMemcpyHostToDevice) != cudaSuccess) {
          set_error("window profile upload failed");
          return fail(OX_ERR_CUDA);
        }
        pl->win_separable = true;
      }
    }
  }
  size_t d = pl->dim;
  if ((st = pl->stat.ensure(sizeof(double) * (1 + d + d * d))) != OX_OK) return fail(st);
  if ((st = pl->bp.ensure(sizeof(double) * d * s->max_batch)) != OX_OK) return fail(st);
  // hand-written fused FFT path when the geometry allows it (ORPHX_PIPELINE=cufft|fused overrides)
  pl->path = 1;
  const char *env = getenv("ORPHX_PIPELINE");
  bool want_fused = fused_supported(s->g->ny, s->g->nx, s->ncomp, s->dtype) && !(env && !strcmp(env, "cufft"));
  if (env && !strcmp(env, "fused") && !want_fused) {
    set_error("ORPHX_PIPELINE=fused but %dx%d x%d comps is not supported by the fused path", s->g->ny, s->g->nx, s->ncomp);
    return fail(OX_ERR_UNSUPPORTED);
  }
  if (want_fused) {
    if ((st = fused_prepare(s, b, pl->fused)) != OX_OK) return fail(st);
    pl->path = 2;
  }
  *out = pl;
  return ox_pipeline_stats_reset(pl);
}

int ox_pipeline_destroy(ox_pipeline *pl) {
  delete pl;
  return OX_OK;
}

int ox_pipeline_stats_reset(ox_pipeline *pl) {
  OX_REQUIRE(pl, "null pipeline");
  size_t d = pl->dim;
  OX_CUDA(cudaMemsetAsync(pl->stat.p, 0, sizeof(double) * (1 + d + d * d), g_stream));
  return OX_OK;
}

int ox_pipeline_stats(ox_pipeline *pl, double **packed_dev, int *dim) {
  OX_REQUIRE(pl, "null pipeline");
  if (packed_dev) *packed_dev = pl->stat.as<double>();
  if (dim) *dim = pl->dim;
  return OX_OK;
}

static int pipeline_run_impl(ox_pipeline *pl, const long long *seeds, int nsim, int noise_mode, const double *noise,
                             int noise_where, int flags, double *bandpowers, int out_where, cudaEvent_t *ev /*7 or null*/) {
  OX_REQUIRE(pl, "null pipeline");
  ox_simplan *s = pl->s;
  ox_powerplan *p = pl->p;
  ox_geometry *g = s->g;
  OX_REQUIRE(nsim >= 1 && nsim <= s->max_batch && nsim <= p->max_batch, "nsim=%d outside 1..max_batch", nsim);
  size_t es = elem_size(s->dtype);
  if (pl->path == 2) {
    const double *noise_dev;
    OX_TRY(sim_stage_inputs(s, seeds, nsim, noise_mode, noise, noise_where, &noise_dev));
    OX_REQUIRE(!(flags & OX_FLAG_ROT) || s->ncomp == 3, "EB->QU rotation needs ncomp == 3");
    pl->maps_valid = (flags & OX_FLAG_KEEP_MAPS) != 0;
    OX_TRY(fused_run(pl, nsim, noise_mode, noise_dev, flags, pl->maps_valid, ev));
    long long total2 = (long long)pl->dim * pl->dim;
    stats_accumulate_kernel<<<(unsigned)((total2 + 255) / 256), 256, 0, g_stream>>>(
        pl->bp.as<double>(), nsim, pl->dim, pl->stat.as<double>());
    OX_KERNEL_CHECK();
    if (ev) OX_CUDA(cudaEventRecord(ev[6], g_stream));
    if (bandpowers) OX_TRY(stage_out(bandpowers, out_where, pl->bp.p, sizeof(double) * (size_t)nsim * pl->dim));
    return OX_OK;
  }
#define OX_MARK(i) do { if (ev) OX_CUDA(cudaEventRecord(ev[i], g_stream)); } while (0)
  OX_MARK(0);
  pl->maps_valid = true;
  // 1. k_h = Hermitian part of covsqrt.noise / sqrt(Npix)          (hand-written)
  OX_TRY(sim_fill_half(s, seeds, nsim, noise_mode, noise, noise_where, flags));
  OX_MARK(1);
  // 2. real maps                                                   (cuFFT Z2D)
  OX_TRY(sim_to_maps(s, nsim));
  OX_MARK(2);
  // 3. real-space taper                                            (hand-written)
  long long npix = (long long)g->ny * g->nx;
  if (pl->has_window) OX_TRY(apply_window(s->dtype, s->maps.p, pl->window.p, npix, (long long)nsim * s->ncomp));
  OX_MARK(3);
  // 4. forward transform                                           (cuFFT D2Z)
  OX_TRY(p->kh1.ensure(2 * es * (size_t)p->max_batch * p->ncomp * g->ny * g->nxh));
  OX_TRY(p->fft.exec_r2c(nsim * s->ncomp, s->maps.p, p->kh1.p));
  OX_MARK(4);
  // 5. conj(k).k, QU->EB, annular binning, /count                   (hand-written, fused)
  OX_TRY(power_bin_half(g, pl->b, s->dtype, s->ncomp, p->kh1.p, nullptr, nsim, flags & ~OX_FLAG_SKIP_CROSS, p->normfact,
                        pl->partial, pl->bp.as<double>()));
  OX_MARK(5);
  // 6. Statistics triple
  long long total = (long long)pl->dim * pl->dim;
  stats_accumulate_kernel<<<(unsigned)((total + 255) / 256), 256, 0, g_stream>>>(
      pl->bp.as<double>(), nsim, pl->dim, pl->stat.as<double>());
  OX_KERNEL_CHECK();
  OX_MARK(6);
#undef OX_MARK
  if (bandpowers) OX_TRY(stage_out(bandpowers, out_where, pl->bp.p, sizeof(double) * (size_t)nsim * pl->dim));
  return OX_OK;
}

int ox_pipeline_run(ox_pipeline *pl, const long long *seeds, int nsim, int noise_mode, const double *noise, int noise_where,
                    int flags, double *bandpowers, int out_where) {
  return pipeline_run_impl(pl, seeds, nsim, noise_mode, noise, noise_where, flags, bandpowers, out_where, nullptr);
}

int ox_pipeline_path(ox_pipeline *pl, int *path) {
  OX_REQUIRE(pl && path, "null pointer");
  *path = pl->path;
  return OX_OK;
}

int ox_pipeline_maps(ox_pipeline *pl, void **maps_dev) {
  OX_REQUIRE(pl && maps_dev, "null pointer");
  *maps_dev = pl->maps_valid ? pl->s->maps.p : nullptr;  // null: the last (fused) run did not keep them
  return OX_OK;
}

int ox_pipeline_profile(ox_pipeline *pl, const long long *seeds, int nsim, int noise_mode, int flags, float *stage_ms) {
  OX_REQUIRE(pl && stage_ms, "null pointer");
  cudaEvent_t ev[7];
  for (int i = 0; i < 7; i++) OX_CUDA(cudaEventCreate(&ev[i]));
  int st = pipeline_run_impl(pl, seeds, nsim, noise_mode, nullptr, OX_HOST, flags, nullptr, OX_DEVICE, ev);
  if (st == OX_OK) {
    cudaError_t e = cudaEventSynchronize(ev[6]);
    for (int i = 0; i < 6 && e == cudaSuccess; i++) e = cudaEventElapsedTime(&stage_ms[i], ev[i], ev[i + 1]);
    if (e != cudaSuccess) {
      set_error("ox_pipeline_profile: %s", cudaGetErrorString(e));
      st = OX_ERR_CUDA;
    }
  }
  for (int i = 0; i < 7; i++) cudaEventDestroy(ev[i]);
  return st;
}

}  // extern "C"
