/* orphx.h -- C ABI of liborphx.so: the B200-native flat-sky Fourier hot path that
 * sits behind the orphics Python API (MapGen / FourierCalc / bin2D / qest).
 *
 * The reference (msyriac/orphics) is pure Python and has no FFI of its own: its
 * hot path calls numpy and pixell.enmap directly.  The entry points below are
 * therefore placed exactly at those call levels; each one names the reference
 * interface (file:line under /root/reference/orphics/) it replaces.  The ctypes
 * binding a maintainer would add is shown in INTEGRATION.md and lives in
 * orphics_b200/_capi.py.
 *
 * Conventions
 *   - every function returns an int status: 0 = OX_OK, negative = error; the
 *     message of the last error on the calling thread is ox_last_error().
 *   - plain pointers and sizes only.  `where` arguments say whether a data
 *     pointer is host (OX_HOST) or device (OX_DEVICE) memory; host transfers are
 *     done inside the call on the library stream.
 *   - the caller owns every buffer it passes; handles own their device memory
 *     and cuFFT plans.  Calls on one handle are not thread-safe.
 *   - dtype: OX_F64 (double / double2) or OX_F32 (float / float2) arithmetic.
 *   - maps are C-ordered [batch][ncomp][Ny][Nx]; "half-plane" Fourier arrays are
 *     [batch][ncomp][Ny][Nx/2+1] complex (the r2c layout); "full-plane" Fourier
 *     arrays are [batch][ncomp][Ny][Nx] complex (the layout numpy/pixell c2c
 *     returns).
 */
#ifndef ORPHX_H
#define ORPHX_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OX_ABI_VERSION 2

enum { OX_OK = 0, OX_ERR_INVALID = -1, OX_ERR_CUDA = -2, OX_ERR_CUFFT = -3, OX_ERR_NOMEM = -4, OX_ERR_UNSUPPORTED = -5 };
enum { OX_F64 = 0, OX_F32 = 1 };
enum { OX_HOST = 0, OX_DEVICE = 1 };

/* noise source of ox_sim_* (MapGen.get_map, maps.py:1577-1578) */
enum {
  OX_NOISE_HOST = 0,            /* caller uploads rand_gauss_harm's two standard_normal blocks: seed parity with numpy */
  OX_NOISE_PHILOX = 1,          /* Philox4x32-10 + Box-Muller keyed by (seed, component, full-plane pixel): the reference algorithm, counter RNG */
  OX_NOISE_PHILOX_HERMITIAN = 2 /* draws the Hermitian half-plane directly (half the normals); same statistics */
};

/* flag bits */
enum {
  OX_FLAG_ROT = 1,          /* QU<->EB rotation on the last two of 3 components (harm2map maps.py:1587 / iqu2teb maps.py:1614-1615) */
  OX_FLAG_SKIP_CROSS = 2,   /* power2d(skip_cross=True) maps.py:1666 */
  OX_FLAG_PIXEL_UNITS = 4,  /* f2power(pixel_units=True) maps.py:1622 */
  OX_FLAG_IAU = 8,          /* iau sign convention of queb_rotmat (maps.py:1600,1607) */
  OX_FLAG_MASK_NAN = 16,    /* bin2D.bin(mask_nan=True) stats.py:792-793 */
  OX_FLAG_HARM = 32,        /* get_map(harm=True) maps.py:1581-1582 */
  OX_FLAG_UNITARY = 64,     /* enmap.fft(normalize=True): x Npix^-1/2 */
  OX_FLAG_KEEP_MAPS = 128   /* ox_pipeline_run: also store the real-space maps (before the taper) in HBM */
};

typedef struct ox_geometry ox_geometry;
typedef struct ox_binner ox_binner;
typedef struct ox_simplan ox_simplan;
typedef struct ox_powerplan ox_powerplan;
typedef struct ox_pipeline ox_pipeline;
typedef struct ox_qeplan ox_qeplan;
typedef struct ox_comm ox_comm;
typedef struct ox_lensplan ox_lensplan;

/* ---- runtime ------------------------------------------------------------- */
int ox_abi_version(void);
const char *ox_last_error(void);
int ox_device_count(int *n);
int ox_set_device(int dev);
int ox_get_device(int *dev);
int ox_device_name(char *buf, size_t len);
int ox_synchronize(void);
/* library stream: NULL (default) = the CUDA legacy default stream, which is also
 * torch's default stream, so torch.distributed collectives order after our kernels. */
int ox_set_stream(void *cuda_stream);
int ox_malloc(void **dptr, size_t bytes);
int ox_free(void *dptr);
/* stream-ordered pool on the library stream (cudaMallocAsync, freed blocks are kept): the buffers behind
 * the device-resident maps the Python mirror returns from get_map / power2d / ... (enmap.devmap) */
int ox_malloc_pooled(void **dptr, size_t bytes);
int ox_free_pooled(void *dptr);
/* Elementwise arithmetic between device-resident maps -- the numpy expressions of the reference's call
 * sequences (`imap*mask` maps.py:1359, `beamed+noise_map` lensing.py:519, `p1d/w2`): out = a (op) b, all device
 * pointers.  kind: 0 float64, 1 float32 (a, b, out real), 2 complex128 a/out with float64 b, 3 complex64 with
 * float32 b.  op: 0 a*b, 1 a+b, 2 a-b, 3 a/b, 4 b-a, 5 b/a (real kinds).  b == NULL: the scalar.  b has nb
 * elements and repeats over a's n (a (ncomp,Ny,Nx) map times a (Ny,Nx) taper). */
int ox_map_op(int op, const void *a, const void *b, double scalar, long long n, long long nb, int kind, void *out);
int ox_memset(void *dptr, int value, size_t bytes);
int ox_host_alloc(void **hptr, size_t bytes); /* pinned */
int ox_host_free(void *hptr);
int ox_memcpy_h2d(void *dst, const void *src, size_t bytes);
int ox_memcpy_d2h(void *dst, const void *src, size_t bytes);
int ox_memcpy_d2d(void *dst, const void *src, size_t bytes);
int ox_mem_info(size_t *free_bytes, size_t *total_bytes);
/* device timers (CUDA events on the library stream) */
int ox_timer_create(void **t);
int ox_timer_start(void *t);
int ox_timer_stop(void *t);
int ox_timer_elapsed_ms(void *t, float *ms); /* synchronises on the stop event */
int ox_timer_destroy(void *t);
/* number of kernels / cuFFT executions this library has launched in this process */
int ox_launch_count(long long *n);
/* write `bytes` of zeros to a scratch buffer larger than L2 (bench hygiene) */
int ox_flush_l2(void);
/* per-stage device times of whatever runs between begin and end: the library records a CUDA event on its
 * stream after each named stage (kernel or group of kernels); stage i spent `ms` since the previous mark.
 * Measurement only (bench.py's roofline.achieved); no reference counterpart. */
int ox_profile_begin(void);
int ox_profile_end(int *nstages);
int ox_profile_stage(int i, char *name, size_t len, float *ms);

/* ---- geometry: enmap.laxes / lmap / modlmap / area (maps.py:1374,1605,1607,1938-1939)
 * ly[ny], lx[nx] are the caller's 2*pi*fftfreq axes (host); area in steradians.   */
int ox_geometry_create(int ny, int nx, const double *ly, const double *lx, double area, ox_geometry **out);
int ox_geometry_destroy(ox_geometry *g);
/* modlmap = (ly^2+lx^2)^1/2, bit-identical to numpy's sum(lmap**2,0)**0.5 */
int ox_geometry_modlmap(ox_geometry *g, double *out, int where);
/* queb_rotmat (maps.py:1607): out[2][2][ny][nx] */
int ox_geometry_rotmat(ox_geometry *g, int flags, double *out, int where);
/* maps.mask_kspace (maps.py:1936-1948); NaN = "None" for the four cuts; out int32[ny][nx] */
int ox_geometry_mask_kspace(ox_geometry *g, double lxcut, double lycut, double lmin, double lmax, int *out, int where);
/* order-1 interpolation of tabulated spectra at modlmap, zero outside the table
 * (the 2-D half of enmap.spec2flat, maps.py:1573): spec[nspec][nl] -> out[nspec][ny][nx] */
int ox_geometry_interp_spec(ox_geometry *g, const double *spec, int nspec, int nl, double *out, int where);

/* ---- stats.bin2D (stats.py:782-811) --------------------------------------- */
/* bin2D.__init__ (stats.py:783-788): digitize(modrmap.ravel(), edges, right=True) */
int ox_binner_create(const double *modrmap, int where, long long n, const double *edges, int nedges, ox_binner **out);
/* same from a geometry's modlmap; additionally enables the fused half-plane path */
int ox_binner_create_geom(ox_geometry *g, const double *edges, int nedges, ox_binner **out);
int ox_binner_destroy(ox_binner *b);
int ox_binner_digitized(ox_binner *b, long long *out_host);      /* int64[n], values 0..nedges */
int ox_binner_counts(ox_binner *b, long long *out_host);         /* int64[nedges+1] = bincount(digitized, minlength=nedges+1) */
/* bin2D.bin (stats.py:790-811), raw slot sums: for each of nmaps maps of n pixels
 *   sums[m][s]   = sum_{digitized==s, kept} data*(weights or 1)
 *   counts[m][s] = sum_{digitized==s, kept} (weights or 1)
 * s = 0..nedges; the caller applies the reference's [1:-1] trim and the division. */
int ox_binner_bin(ox_binner *b, const void *data, int dtype, int where, long long nmaps, const void *weights, int flags,
                  double *sums, double *counts, int out_where);

/* ---- maps.MapGen (maps.py:1553-1587) --------------------------------------- */
/* covsqrt[ncomp][ncomp][ny][nx] float64 (what MapGen.__init__ stores, maps.py:1564-1573) */
int ox_simplan_create(ox_geometry *g, int ncomp, const double *covsqrt, int where, int dtype, int max_batch, ox_simplan **out);
int ox_simplan_destroy(ox_simplan *p);
/* MapGen.get_map for nsim seeds.  noise (OX_NOISE_HOST only): float64
 * [nsim][2][ncomp][ny][nx] = the real block then the imaginary block of
 * rand_gauss_harm (maps.py:1578).  Output: real maps [nsim][ncomp][ny][nx] of
 * `dtype`, or with OX_FLAG_HARM the full-plane complex covsqrt*rand (maps.py:1579-1582).
 * OX_FLAG_ROT = harm2map's EB->QU rotation (scalar=False, ncomp==3). */
int ox_sim_generate(ox_simplan *p, const long long *seeds, int nsim, int noise_mode, const double *noise, int noise_where,
                    int flags, void *out, int out_where);

/* ---- maps.FourierCalc (maps.py:1594-1677) ---------------------------------- */
int ox_powerplan_create(ox_geometry *g, int ncomp, int dtype, int max_batch, ox_powerplan **out);
int ox_powerplan_destroy(ox_powerplan *p);
/* FourierCalc.iqu2teb / .fft (maps.py:1609-1617,1635-1636): real maps -> full-plane complex */
int ox_power_fft(ox_powerplan *p, const void *maps, int where, int nbatch, int flags, void *kmap_out, int out_where);
/* FourierCalc.ifft (maps.py:1632-1633): full-plane complex -> full-plane complex / Npix */
int ox_power_ifft(ox_powerplan *p, const void *kmap, int where, int nbatch, void *out, int out_where);
/* FourierCalc.f2power (maps.py:1620-1624) on n complex elements */
int ox_power_f2power(ox_powerplan *p, const void *k1, const void *k2, int where, long long n, int flags, void *out, int out_where);
/* FourierCalc.power2d (maps.py:1639-1677): p2d[nbatch][ncomp][ncomp][ny][nx] real,
 * optional full-plane kmaps.  maps2 may be NULL (auto). */
int ox_power2d(ox_powerplan *p, const void *maps1, const void *maps2, int where, int nbatch, int flags, void *p2d_out,
               void *kmap1_out, void *kmap2_out, int out_where);
/* fused power2d + bin2D.bin (maps.binned_power, maps.py:1350-1361) without
 * materialising p2d: bandpowers[nbatch][nspec][nedges-1] float64, nspec =
 * ncomp(ncomp+1)/2 ordered (0,0),(0,1)..(0,n-1),(1,1).. ; window (may be NULL) is a
 * real [ny][nx] taper multiplied into the maps first. The binner must come from
 * ox_binner_create_geom. */
int ox_power_bin(ox_powerplan *p, ox_binner *b, const void *maps1, const void *maps2, int where, int nbatch, int flags,
                 const void *window, int window_where, double *bandpowers, int out_where);

/* ---- fused sim -> FFT -> power2d -> bin2D pipeline (north-star path) -------- */
int ox_pipeline_create(ox_simplan *s, ox_powerplan *p, ox_binner *b, const double *window, int window_where, ox_pipeline **out);
int ox_pipeline_destroy(ox_pipeline *pl);
/* bandpowers[nsim][nspec][nbins]; also accumulates the Statistics triple
 * (stats.py:1085-1090): N += nsim, SUM += x, CROSS += x x^T with x = the
 * flattened [nspec*nbins] bandpower vector of each sim. */
int ox_pipeline_run(ox_pipeline *pl, const long long *seeds, int nsim, int noise_mode, const double *noise, int noise_where,
                    int flags, double *bandpowers, int out_where);
/* which implementation the pipeline uses: 1 = cuFFT passes + hand-written kernels around them,
 * 2 = fused hand-written FFT kernels (power-of-two maps; ORPHX_PIPELINE=cufft|fused overrides) */
int ox_pipeline_path(ox_pipeline *pl, int *path);
/* device pointer of the real maps [max_batch][ncomp][ny][nx] of the last run (path 1: always;
 * path 2: when OX_FLAG_KEEP_MAPS was given) */
int ox_pipeline_maps(ox_pipeline *pl, void **maps_dev);
/* one ox_pipeline_run with CUDA events between the stages; stage_ms[6] = sim_fill, cuFFT
 * inverse, window, cuFFT forward, power_bin (+finalize), statistics.  Philox modes only. */
int ox_pipeline_profile(ox_pipeline *pl, const long long *seeds, int nsim, int noise_mode, int flags, float *stage_ms);
/* device pointer of the accumulators, packed so that the reduction over ranks is ONE collective
 * (stats.py:1215-1217): float64 [N | SUM[dim] | CROSS[dim][dim]], N an exactly represented integer */
int ox_pipeline_stats(ox_pipeline *pl, double **packed_dev, int *dim);
int ox_pipeline_stats_reset(ox_pipeline *pl);

/* ---- lensing.qest (tutorials/tt_verification.ipynb:81,608,610; lensing.py:973-976) */
enum { OX_QE_TT = 0, OX_QE_EB = 1 };
/* filters are full-plane [ny][nx] float64: wxy = W_XY, wy = W_Y (filter x mask x beam, the historical
 * QuadNorm.WXY/WY), norm = the multiplier applied in kappa_from_map, N_L 2/(L(L+1)), x kmask_K.
 * real_path != 0: the filters vanish on the Nyquist row/column and are symmetric under l -> -l, so for
 * Hermitian inputs every field is real (TT) or splits into real and imaginary parts that are each the
 * transform of a Hermitian array (EB), and the chain runs on half planes with r2c/c2r transforms. */
int ox_qeplan_create(ox_geometry *g, int est, const double *wxy, const double *wy, const double *norm, int where, int dtype,
                     int max_batch, int real_path, ox_qeplan **out);
int ox_qeplan_destroy(ox_qeplan *q);
/* qest.kappa_from_map(XY, X-leg, Y-leg, alreadyFTed, returnFt) for nbatch realisations.  x = T (TT) or
 * E (EB), y = NULL/same (TT) or B (EB); real maps [nbatch][ny][nx] of the plan dtype, or with already_ft
 * full-plane complex k-maps.  Output: real kappa maps, or (return_ft) full-plane complex kappa_hat(l).
 * accumulate_meanfield: add every kappa_hat(l) (half plane) and nbatch to the plan's mean-field stack. */
int ox_qe_reconstruct(ox_qeplan *q, const void *x, const void *y, int where, int nbatch, int already_ft, int return_ft,
                      int accumulate_meanfield, void *kappa_out, int out_where);
/* device pointer of the mean-field accumulator, packed for ONE collective (Statistics.add_stack / allreduce,
 * stats.py:1134-1158,1227-1228): float64 [2*nelem: complex128 [ny][nx/2+1] sum of kappa_hat(l) | count | pad] */
int ox_qe_meanfield(ox_qeplan *q, double **packed_dev, long long *nelem);
int ox_qe_meanfield_reset(ox_qeplan *q);
/* which implementation the plan runs for Hermitian inputs: 0 = full-plane c2c chain on cuFFT (unsymmetric
 * filters, or EB on maps that are not powers of two), 1 = TT on half planes with cuFFT r2c/c2r, 2 = TT and
 * 3 = EB on half planes with the hand-written FFT passes (power-of-two maps; ORPHX_QE=cufft in the
 * environment disables 2 and 3).  already_ft inputs that are not Hermitian (checked on every pixel pair) always take the c2c chain. */
int ox_qe_path(ox_qeplan *q);

/* maps.filter_map (maps.py:1922-1923): Re(ifft(fft(m) * kfilter)) / Npix for nbatch x ncomp real maps;
 * kfilter is a real float64 full-plane [ny][nx] array (beam, l-mask, Wiener filter) */
int ox_power_filter(ox_powerplan *p, const void *maps, int where, int nbatch, const double *kfilter, int kwhere, void *out,
                    int out_where);
/* the same with a complex128 full-plane kfilter (maps.py:1923 takes any array): out = Re ifft(fft(m) kfilter) / Npix */
int ox_power_filter_complex(ox_powerplan *p, const void *maps, int where, int nbatch, const void *kfilter, int kwhere, void *out,
                            int out_where);

/* ---- generic batched c2c FFT on full-plane complex arrays (pixell.fft.fft / ifft, lensing.py:20):
 * out = scale * FFT_direction(in), direction -1 forward / +1 backward, nplanes arrays [ny][nx] */
int ox_fft_c2c(ox_powerplan *p, const void *in, int where, int nplanes, int direction, double scale, void *out, int out_where);

/* ---- split / cross-spectrum callers.  Arrays are full-plane [ny][nx]; npix = ny*nx.
 * maps.split_calc (maps.py:2295-2332): isplits/jsplits complex128 [ni|nj][npix] Fourier transforms of the
 * splits, icoadd/jcoadd complex128 [npix]; outputs float64 [npix]: total, crosses (signal estimate), noise.
 * normfact = FourierCalc.normfact (f2power, maps.py:1620-1624). */
int ox_split_calc(const void *isplits, const void *jsplits, const void *icoadd, const void *jcoadd, int where, int ni, int nj,
                  long long npix, double normfact, int alt, double *total, double *crosses, double *noise, int out_where);
/* maps.noise_from_splits (maps.py:2337-2411): splits are real maps [nsplits][ncomp][ny][nx] of the plan dtype
 * (the caller applies the reference's float32 cast); outputs float64 [ncomp][ncomp][ny][nx]: the noise model
 * (auto - cross)/nsplits and, with do_cross, the mean cross spectrum of the i < j split pairs ("cross_teb":
 * the reference never applies its Q,U -> E,B rotation there, maps.py:2359,2379 -- reproduced). */
int ox_noise_from_splits(ox_powerplan *p, const void *splits, int where, int nsplits, int ncomp, int do_cross, double *noise,
                         double *crossteb, int out_where);
/* lensing.SplitLensing.cross_estimator (lensing.py:980-1003): per-pixel combination of the stacked
 * kappa_hat(l) arrays (complex of `dtype`, [1 + 3 n + n(n-1)][npix]) in the order q(s,s); for each split i:
 * q(m_i,s), q(s,m_i), q(m_i,m_i); then for each i < j: q(m_i,m_j), q(m_j,m_i).  Output float64 [npix]. */
int ox_split_lensing_combine(const void *khat, int where, int dtype, int nsplits, long long npix, double normfact, double *out,
                             int out_where);

/* Fourier-space internal linear combination (maps.py:1952-2050), one pass over the pixels:
 * mode 0 silc, 1 cilc -> complex128 [npix]; 2 silc_noise, 3 cilc_noise -> float64 [npix].
 * kmaps complex128 [nfreq][npix] (modes 0, 1), cinv float64 [nfreq][nfreq][npix] (npix = Ny*Nx for 2-D
 * k-space matrices or the number of bins for 1-D spectra), response vectors float64 [nfreq] on the HOST
 * (response_a NULL = ones, the CMB).  np.nan_to_num semantics as in the reference. */
int ox_ilc(const void *kmaps, const double *cinv, const double *response_a, const double *response_b, int nfreq, long long npix,
           int where, int mode, void *out, int out_where);

/* enmap.multi_pow(mat, exp, axes=[0,1]) as MapGen uses it for a 2-D covariance (maps.py:1571): per-pixel power of
 * the symmetric n x n matrices mat[n][n][npix] (float64, n <= 4) through a Jacobi eigen-decomposition; for a
 * non-integer or negative exponent, eigenvalues that are negative or below 1e-13 of the largest are zeroed. */
int ox_multi_pow(const double *mat, int n, long long npix, double exponent, int where, double *out, int out_where);

/* ---- set-up of the estimator on the device (historical QuadNorm, call site tutorials/tt_verification.ipynb:81).
 * All tables are float64 full planes [ny][nx] in DEVICE memory (results too); nullable ones are noted.
 *  ox_qe_filter  out = nan_to_num(num / (lcl beam^2 + noise)) beam, zeroed where mask < 1e-3, where L > cut_gt and where
 *                L >= cut_ge: W_XY (num = uC^{XY'}, cut_gt = gradCut, cut_ge = bigell) and W_Y (num NULL = 1, cut_gt = inf).
 *                num, noise, beam, mask may be NULL (1, 0, 1, no mask).
 *  ox_qe_norm    A_L of estimator est (OX_QE_TT / OX_QE_EB) from cl = uC^{TT} (uC^{EE} for EB), w1 = W_XY beam, w2 = W_Y beam:
 *                three inverse transforms + one forward per term, 13 (TT) / 24 (EB) cuFFT c2c transforms in all;
 *                nlkk_out = N_L^{kappa kappa} (`.N.Nlkk[XY]`), al_out = the multiplier kappa_from_map applies
 *                (N_L 2 / (L (L+1))).  kmask_k may be NULL; pix_area = pixScaleX pixScaleY. */
/* res_host[3] = { max |a(l) - a(-l)|, max |a|, 1.0 if the Nyquist row or column holds a non-zero } of a device plane:
 * the test that selects the estimator's half-plane paths (symmetric filters vanishing at Nyquist) */
int ox_plane_symmetry(ox_geometry *g, const double *a_dev, double *res_host);
int ox_qe_filter(ox_geometry *g, const double *num, const double *lcl, const double *noise, const double *beam, const double *mask,
                 double cut_gt, double cut_ge, double *out);
int ox_qe_norm(ox_geometry *g, int est, const double *cl, const double *w1, const double *w2, const double *kmask_k, double bigell,
               double pix_area, double *nlkk_out, double *al_out);

/* ---- lensing of flat-sky maps (the step before the estimator in FlatLensingSims.get_sim, lensing.py:499-521).
 * A plan belongs to a geometry; py, px = pixel height and width in radians (enmap.pixshape, lensing.py:420);
 * max_planes >= the largest taylor_order used (>= 2).  float64 throughout, maps [nmaps][ny][nx].
 *  ox_lens_kappa_to_phi  lensing.kappa_to_phi (lensing.py:651-665): phi = Re ifft(2 fft(kappa) / (L (L+1))), 0 for L < 2.
 *  ox_lens_set_phi       deflection alpha = Re ifft(i l fft(phi)) and its split into nearest pixel + remainder
 *                        (flat_taylens, lensing.py:411-427), kept in the plan.
 *  ox_lens_alpha         that deflection, [2][ny][nx] = (alphaY, alphaX) in radians (alpha_from_kappa's grad phi, lensing.py:443-449).
 *  ox_lens_taylens       flat_taylens (lensing.py:395-440): imap[(iy+aY0)%Ny, (ix+aX0)%Nx] + the Taylor series in the
 *                        remainder up to taylor_order - 1, every derivative a half-plane inverse transform.
 *  ox_lens_displace      bicubic (Keys a = -1/2) interpolation at (iy + alphaY/py, ix + alphaX/px), periodic: the stand-in
 *                        for pixell.lensing.displace_map (lensing.py:512; third party, spline of order lens_order). */
int ox_lensplan_create(ox_geometry *g, double py, double px, int max_planes, ox_lensplan **out);
int ox_lensplan_destroy(ox_lensplan *p);
int ox_lens_kappa_to_phi(ox_lensplan *p, const double *kappa, int where, double *phi_out, int out_where);
int ox_lens_set_phi(ox_lensplan *p, const double *phi, int where);
int ox_lens_alpha(ox_lensplan *p, double *alpha_out, int out_where);
int ox_lens_taylens(ox_lensplan *p, const double *imap, int where, int nmaps, int taylor_order, double *out, int out_where);
int ox_lens_displace(ox_lensplan *p, const double *imap, int where, int nmaps, double *out, int out_where);

/* ---- the exchange step: Statistics.allreduce (stats.py:1184-1232, mpi4py Allreduce(SUM)) as NCCL sum
 * all-reduces over NVLink, one process per GPU.  Rank 0 obtains the 128-byte NCCL unique id and the host
 * hands it to the other ranks (torch.distributed store, MPI_Bcast, a file); every rank then creates its
 * communicator on its current device.  nranks == 1 needs no id and makes every reduction a no-op.
 * NCCL is bound at run time (dlopen libnccl.so.2, or $ORPHX_NCCL_LIB). */
#define OX_COMM_ID_BYTES 128
int ox_comm_unique_id(void *id, size_t len);
int ox_comm_create(int rank, int nranks, const void *id, size_t len, ox_comm **out);
int ox_comm_destroy(ox_comm *c);
int ox_comm_info(ox_comm *c, int *rank, int *nranks, int *nccl_version);
/* in-place sum over the ranks of count float64 values in device memory, on the library stream */
int ox_comm_allreduce_f64(ox_comm *c, double *buf_dev, long long count);
/* ONE ncclAllReduce of the pipeline's packed [N | SUM | CROSS] (stats.py:1215-1217) */
int ox_pipeline_allreduce(ox_comm *c, ox_pipeline *pl);
/* ONE ncclAllReduce of the estimator's packed [mean-field stack | count] (stats.py:1227-1228) */
int ox_qe_meanfield_allreduce(ox_comm *c, ox_qeplan *q);

#ifdef __cplusplus
}
#endif
#endif /* ORPHX_H */
