// Split / cross-spectrum callers of the hot path: O(nsplits^2) combinations of f2power and
// of the quadratic estimator that the reference evaluates pair by pair in Python, as single device passes.
//   maps.split_calc            maps.py:2295-2332
//   maps.noise_from_splits     maps.py:2337-2411
//   lensing.SplitLensing.cross_estimator   lensing.py:980-1003 (the per-pixel combination; the estimators
//                                           themselves run through ox_qe_reconstruct)
// All sums run in the reference's order per pixel in double precision (no atomics).
#include "ox_common.cuh"

using namespace ox;

namespace {

constexpr int ST = 256;
constexpr int MAXC = 6;  // components per split of noise_from_splits (I,Q,U of up to two arrays)

// ---- split_calc: full-plane complex128 in, float64 out
__global__ void split_calc_kernel(const double2 *__restrict__ is, const double2 *__restrict__ js, const double2 *__restrict__ ic,
                                  const double2 *__restrict__ jc, int ni, int nj, long long n, double nf, int alt,
                                  double *__restrict__ total, double *__restrict__ crosses, double *__restrict__ noise) {
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= n) return;
  const double2 A = ic[p], B = jc[p];
  const double tot = (A.x * B.x + A.y * B.y) * nf;  // Re(conj(A) B) x normfact (f2power, maps.py:1620)
  double cr, nz;
  if (alt) {
    double acc = 0.0;
    for (int i = 0; i < ni; i++) {
      const double2 a = is[i * n + p], b = js[i * n + p];
      const double dx = a.x - A.x, dy = a.y - A.y, ex = b.x - B.x, ey = b.y - B.y;
      acc += (dx * ex + dy * ey) * nf;
    }
    nz = acc / ((1.0 - 1.0 / ni) * (double)ni * (double)ni);
    cr = tot - nz;
  } else {
    double acc = 0.0, cnt = 0.0;
    for (int i = 0; i < ni; i++) {
      const double2 a = is[i * n + p];
      for (int j = 0; j < nj; j++) {
        if (i == j) continue;
        const double2 b = js[j * n + p];
        acc += (a.x * b.x + a.y * b.y) * nf;
        cnt += 1.0;
      }
    }
    cr = acc / cnt;
    nz = tot - cr;
  }
  total[p] = tot;
  crosses[p] = cr;
  noise[p] = nz;
}

// ---- noise_from_splits on half planes kh[nsplits][nc][ny][nxh]; outputs [nc][nc][ny][nx] float64
// auto_ab = sum_s Re(conj k_s[a] k_s[b]) / n, cross_ab = sum_{i<j} Re(conj k_i[a] k_j[b]) / npairs for a <= b
// (power2d fills the upper triangle and mirrors it, maps.py:1664-1671); noise = (auto - cross)/n.
// "cross_teb" is the same cross spectrum: the reference never applies its Q,U -> E,B rotation (maps.py:2359,2379).
template <typename T2>
__global__ void noise_from_splits_kernel(const T2 *__restrict__ kh, int nsplits, int nc, int ny, int nx, int nxh, double nf,
                                         double *__restrict__ noise, double *__restrict__ crossteb) {
  const long long nh = (long long)ny * nxh, n = (long long)ny * nx;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= nh) return;
  const int iy = (int)(t / nxh), ix = (int)(t - (long long)iy * nxh);
  double pre_x[MAXC], pre_y[MAXC], au[MAXC * MAXC], cr[MAXC * MAXC];
  for (int a = 0; a < nc; a++) pre_x[a] = pre_y[a] = 0.0;
  for (int e = 0; e < nc * nc; e++) au[e] = cr[e] = 0.0;
  for (int s = 0; s < nsplits; s++) {
    double kx[MAXC], ky[MAXC];
    for (int a = 0; a < nc; a++) {
      const T2 z = kh[((long long)s * nc + a) * nh + t];
      kx[a] = z.x;
      ky[a] = z.y;
    }
    for (int a = 0; a < nc; a++)
      for (int b = a; b < nc; b++) {
        au[a * nc + b] += (kx[a] * kx[b] + ky[a] * ky[b]) * nf;
        cr[a * nc + b] += (pre_x[a] * kx[b] + pre_y[a] * ky[b]) * nf;  // sum over earlier splits i < s
      }
    for (int a = 0; a < nc; a++) {
      pre_x[a] += kx[a];
      pre_y[a] += ky[a];
    }
  }
  const double npairs = 0.5 * nsplits * (nsplits - 1.0);
  const bool mirror = ix > 0 && 2 * ix < nx;
  const long long p = (long long)iy * nx + ix, q = (long long)(iy ? ny - iy : 0) * nx + (nx - ix);
  for (int a = 0; a < nc; a++)
    for (int b = a; b < nc; b++) {
      const double c = cr[a * nc + b] / npairs, nz = (au[a * nc + b] / nsplits - c) / nsplits;
      const long long o1 = ((long long)a * nc + b) * n, o2 = ((long long)b * nc + a) * n;
      noise[o1 + p] = nz;
      noise[o2 + p] = nz;
      if (mirror) {
        noise[o1 + q] = nz;
        noise[o2 + q] = nz;
      }
      if (crossteb) {
        crossteb[o1 + p] = c;
        crossteb[o2 + p] = c;
        if (mirror) {
          crossteb[o1 + q] = c;
          crossteb[o2 + q] = c;
        }
      }
    }
}

// ---- SplitLensing.cross_estimator: per-pixel combination of the stacked estimators
// K[0] = q(s,s); K[1+3i], K[2+3i], K[3+3i] = q(m_i,s), q(s,m_i), q(m_i,m_i); then for i < j (row-major)
// K[..] = q(m_i,m_j), q(m_j,m_i)                                                       (lensing.py:980-1003)
template <typename T2>
__global__ void split_lensing_combine_kernel(const T2 *__restrict__ K, int ns, long long n, double nf, double *__restrict__ out) {
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= n) return;
  const double fn = ns;
  double sx = 0.0, sy = 0.0, psum = 0.0, psum2 = 0.0;
  for (int i = 0; i < ns; i++) {
    const T2 a = K[(1 + 3 * i) * n + p], b = K[(2 + 3 * i) * n + p], c = K[(3 + 3 * i) * n + p];
    sx += c.x;
    sy += c.y;
    const double kx = ((double)a.x + (double)b.x) / 2.0 - (1.0 / fn) * (double)c.x;
    const double ky = ((double)a.y + (double)b.y) / 2.0 - (1.0 / fn) * (double)c.y;
    psum += (kx * kx + ky * ky) * nf;
  }
  long long idx = 1 + 3 * (long long)ns;
  for (int i = 0; i < ns; i++)
    for (int j = i + 1; j < ns; j++) {
      const T2 a = K[idx * n + p], b = K[(idx + 1) * n + p];
      idx += 2;
      const double kx = ((double)a.x + (double)b.x) / 2.0, ky = ((double)a.y + (double)b.y) / 2.0;
      psum2 += (kx * kx + ky * ky) * nf;
    }
  const T2 k0 = K[p];
  const double cx = (double)k0.x - (1.0 / (fn * fn)) * sx, cy = (double)k0.y - (1.0 / (fn * fn)) * sy;
  const double pc = (cx * cx + cy * cy) * nf;
  out[p] = (fn * fn * fn * fn * pc - 4.0 * fn * fn * psum + 4.0 * psum2) / fn / (fn - 1.0) / (fn - 2.0) / (fn - 3.0);
}

unsigned blocks(long long n) { return (unsigned)((n + ST - 1) / ST); }

}  // namespace

extern "C" {

int ox_split_calc(const void *isplits, const void *jsplits, const void *icoadd, const void *jcoadd, int where, int ni, int nj,
                  long long npix, double normfact, int alt, double *total, double *crosses, double *noise, int out_where) {
  OX_REQUIRE(isplits && jsplits && icoadd && jcoadd && total && crosses && noise, "ox_split_calc: null pointer");
  OX_REQUIRE(ni >= 1 && nj >= 1 && npix >= 1, "ox_split_calc: bad sizes");
  OX_REQUIRE(!alt || ni == nj, "split_calc(alt=True) needs as many i splits as j splits (%d vs %d)", ni, nj);
  OX_REQUIRE(alt || ni * (long long)nj > (ni < nj ? ni : nj), "split_calc(alt=False) needs at least one i != j pair");
  DevBuf b0, b1, b2, b3, o;
  const void *is, *js, *ic, *jc;
  const size_t cb = sizeof(double2) * (size_t)npix;
  OX_TRY(stage_in(isplits, where, cb * ni, b0, &is));
  OX_TRY(stage_in(jsplits, where, cb * nj, b1, &js));
  OX_TRY(stage_in(icoadd, where, cb, b2, &ic));
  OX_TRY(stage_in(jcoadd, where, cb, b3, &jc));
  double *t = total, *c = crosses, *z = noise;
  if (out_where == OX_HOST) {
    OX_TRY(o.ensure(3 * sizeof(double) * (size_t)npix));
    t = o.as<double>();
    c = t + npix;
    z = c + npix;
  }
  split_calc_kernel<<<blocks(npix), ST, 0, g_stream>>>((const double2 *)is, (const double2 *)js, (const double2 *)ic,
                                                      (const double2 *)jc, ni, nj, npix, normfact, alt, t, c, z);
  OX_KERNEL_CHECK();
  if (out_where == OX_HOST) {
    OX_TRY(stage_out(total, OX_HOST, t, sizeof(double) * (size_t)npix));
    OX_TRY(stage_out(crosses, OX_HOST, c, sizeof(double) * (size_t)npix));
    OX_TRY(stage_out(noise, OX_HOST, z, sizeof(double) * (size_t)npix));
  }
  OX_CUDA(cudaStreamSynchronize(g_stream));
  return OX_OK;
}

int ox_noise_from_splits(ox_powerplan *p, const void *splits, int where, int nsplits, int ncomp, int do_cross, double *noise,
                         double *crossteb, int out_where) {
  OX_REQUIRE(p && splits && noise, "ox_noise_from_splits: null pointer");
  OX_REQUIRE(nsplits >= 2, "noise_from_splits needs at least two splits (got %d)", nsplits);
  OX_REQUIRE(ncomp >= 1 && ncomp <= MAXC, "noise_from_splits: ncomp must be 1..%d (got %d)", MAXC, ncomp);
  OX_REQUIRE(!do_cross || crossteb, "do_cross needs an output for the cross spectrum");
  ox_geometry *g = p->g;
  const size_t es = elem_size(p->dtype);
  const long long n = (long long)g->ny * g->nx, nh = (long long)g->ny * g->nxh;
  const int planes = nsplits * ncomp;
  DevBuf in, kh, o;
  OX_TRY(in.ensure(es * (size_t)planes * n));
  OX_TRY(kh.ensure(2 * es * (size_t)planes * nh));
  OX_CUDA(cudaMemcpyAsync(in.p, splits, es * (size_t)planes * n, where == OX_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice,
                          g_stream));
  OX_TRY(p->fft.exec_r2c(planes, in.p, kh.p));   // iqu2teb(normalize=False, rot=False), maps.py:2371
  const size_t ob = sizeof(double) * (size_t)ncomp * ncomp * n;
  double *dn = noise, *dc = crossteb;
  if (out_where == OX_HOST) {
    OX_TRY(o.ensure(2 * ob));
    dn = o.as<double>();
    dc = do_cross ? dn + (size_t)ncomp * ncomp * n : nullptr;
  } else if (!do_cross) {
    dc = nullptr;
  }
  if (p->dtype == OX_F64)
    noise_from_splits_kernel<double2><<<blocks(nh), ST, 0, g_stream>>>(kh.as<double2>(), nsplits, ncomp, g->ny, g->nx, g->nxh,
                                                                       p->normfact, dn, dc);
  else
    noise_from_splits_kernel<float2><<<blocks(nh), ST, 0, g_stream>>>(kh.as<float2>(), nsplits, ncomp, g->ny, g->nx, g->nxh,
                                                                      p->normfact, dn, dc);
  OX_KERNEL_CHECK();
  if (out_where == OX_HOST) {
    OX_TRY(stage_out(noise, OX_HOST, dn, ob));
    if (do_cross) OX_TRY(stage_out(crossteb, OX_HOST, dc, ob));
  }
  OX_CUDA(cudaStreamSynchronize(g_stream));
  return OX_OK;
}

int ox_split_lensing_combine(const void *khat, int where, int dtype, int nsplits, long long npix, double normfact, double *out,
                             int out_where) {
  OX_REQUIRE(khat && out, "ox_split_lensing_combine: null pointer");
  OX_REQUIRE(nsplits >= 4, "cross_estimator needs at least four splits (got %d): its normalisation divides by (n-3)", nsplits);
  OX_REQUIRE(dtype == OX_F64 || dtype == OX_F32, "bad dtype %d", dtype);
  const long long nk = 1 + 3LL * nsplits + (long long)nsplits * (nsplits - 1);
  const size_t cb = 2 * elem_size(dtype) * (size_t)npix;
  DevBuf b, o;
  const void *K;
  OX_TRY(stage_in(khat, where, cb * (size_t)nk, b, &K));
  double *d = out;
  if (out_where == OX_HOST) {
    OX_TRY(o.ensure(sizeof(double) * (size_t)npix));
    d = o.as<double>();
  }
  if (dtype == OX_F64)
    split_lensing_combine_kernel<double2><<<blocks(npix), ST, 0, g_stream>>>((const double2 *)K, nsplits, npix, normfact, d);
  else
    split_lensing_combine_kernel<float2><<<blocks(npix), ST, 0, g_stream>>>((const float2 *)K, nsplits, npix, normfact, d);
  OX_KERNEL_CHECK();
  if (out_where == OX_HOST) OX_TRY(stage_out(out, OX_HOST, d, sizeof(double) * (size_t)npix));
  OX_CUDA(cudaStreamSynchronize(g_stream));
  return OX_OK;
}

}  // extern "C"

// ---- Fourier-space ILC (maps.py:1952-2050): per-pixel small linear algebra over nfreq <= 16 channels
namespace {

constexpr int ILC_MAXF = 16;
struct IlcResp {
  double a[ILC_MAXF], b[ILC_MAXF];
};

__device__ __forceinline__ double nan_to_num(double x) {  // np.nan_to_num: NaN -> 0, +-inf -> +-DBL_MAX
  if (isnan(x)) return 0.0;
  if (isinf(x)) return x > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308;
  return x;
}

// a^T Cinv b = sum_l a_l (sum_k b_k Cinv[k][l])   (ilc_comb_a_b, maps.py:2046-2049)
__device__ __forceinline__ double comb(const double *ra, const double *rb, const double *__restrict__ cinv, int nf, long long n,
                                       long long p) {
  double acc = 0.0;
  for (int l = 0; l < nf; l++) {
    double inner = 0.0;
    for (int k = 0; k < nf; k++) inner += rb[k] * cinv[((long long)k * nf + l) * n + p];
    acc += ra[l] * inner;
  }
  return nan_to_num(acc);
}

// r^T Cinv kmaps = sum_k r_k (sum_l Cinv[k][l] kmaps[l])   (ilc_map_term, maps.py:2042-2044)
__device__ __forceinline__ double2 map_term(const double *r, const double *__restrict__ cinv, const double2 *__restrict__ km,
                                            int nf, long long n, long long p) {
  double2 acc = make_double2(0.0, 0.0);
  for (int k = 0; k < nf; k++) {
    double ix = 0.0, iy = 0.0;
    for (int l = 0; l < nf; l++) {
      const double c = cinv[((long long)k * nf + l) * n + p];
      const double2 z = km[(long long)l * n + p];
      ix += c * z.x;
      iy += c * z.y;
    }
    acc.x += r[k] * ix;
    acc.y += r[k] * iy;
  }
  return acc;
}

// mode 0 silc, 1 cilc (complex out), 2 silc_noise, 3 cilc_noise (real out)
__global__ void ilc_kernel(const double2 *__restrict__ km, const double *__restrict__ cinv, IlcResp r, int nf, long long n, int mode,
                           double *__restrict__ out) {
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= n) return;
  if (mode == 0 || mode == 2) {
    const double noise = nan_to_num(1.0 / comb(r.a, r.a, cinv, nf, n, p));
    if (mode == 2) {
      out[p] = noise;
      return;
    }
    const double2 w = map_term(r.a, cinv, km, nf, n, p);
    // complex x real as numpy does it: (w.x + i w.y)(noise + 0 i)
    reinterpret_cast<double2 *>(out)[p] = make_double2(w.x * noise - w.y * 0.0, w.x * 0.0 + w.y * noise);
    return;
  }
  const double brb = comb(r.b, r.b, cinv, nf, n, p), arb = comb(r.a, r.b, cinv, nf, n, p), ara = comb(r.a, r.a, cinv, nf, n, p);
  if (mode == 3) {
    const double bra = comb(r.b, r.a, cinv, nf, n, p);
    const double numer = brb * brb * ara + arb * arb * brb - brb * arb * arb - arb * brb * bra;
    const double d = ara * brb - arb * arb;
    out[p] = nan_to_num(numer / (d * d));
    return;
  }
  const double2 arM = map_term(r.a, cinv, km, nf, n, p), brM = map_term(r.b, cinv, km, nf, n, p);
  const double nx = brb * arM.x - arb * brM.x, ny = brb * arM.y - arb * brM.y;
  const double norm = ara * brb - arb * arb;
  double ox, oy;
  if (norm == 0.0) {  // numpy's complex division by (0 + 0i): componentwise x / 0
    ox = nx / 0.0;
    oy = ny / 0.0;
  } else {            // numpy's complex division by (norm + 0i): multiply by the reciprocal
    const double scl = 1.0 / norm;
    ox = nx * scl;
    oy = ny * scl;
  }
  reinterpret_cast<double2 *>(out)[p] = make_double2(nan_to_num(ox), nan_to_num(oy));
}

}  // namespace

extern "C" int ox_ilc(const void *kmaps, const double *cinv, const double *response_a, const double *response_b, int nfreq,
                      long long npix, int where, int mode, void *out, int out_where) {
  OX_REQUIRE(cinv && out, "ox_ilc: null pointer");
  OX_REQUIRE(mode >= 0 && mode <= 3, "ox_ilc: mode must be 0 (silc), 1 (cilc), 2 (silc_noise) or 3 (cilc_noise)");
  OX_REQUIRE(nfreq >= 1 && nfreq <= ILC_MAXF, "ILC supports 1..%d frequency channels (got %d)", ILC_MAXF, nfreq);
  OX_REQUIRE(mode >= 2 || kmaps, "silc/cilc need the Fourier maps");
  OX_REQUIRE((mode != 1 && mode != 3) || (response_a && response_b), "cilc needs both response vectors");
  IlcResp r;
  for (int i = 0; i < ILC_MAXF; i++) {
    r.a[i] = i < nfreq ? (response_a ? response_a[i] : 1.0) : 0.0;  // default CMB response: ones (maps.py:2007-2013)
    r.b[i] = i < nfreq && response_b ? response_b[i] : 0.0;
  }
  DevBuf b0, b1, o;
  const void *dk = nullptr, *dc;
  if (kmaps) OX_TRY(stage_in(kmaps, where, sizeof(double2) * (size_t)nfreq * npix, b0, &dk));
  OX_TRY(stage_in(cinv, where, sizeof(double) * (size_t)nfreq * nfreq * npix, b1, &dc));
  const size_t ob = (mode <= 1 ? sizeof(double2) : sizeof(double)) * (size_t)npix;
  void *d = out;
  if (out_where == OX_HOST) {
    OX_TRY(o.ensure(ob));
    d = o.p;
  }
  ilc_kernel<<<blocks(npix), ST, 0, g_stream>>>((const double2 *)dk, (const double *)dc, r, nfreq, npix, mode, (double *)d);
  OX_KERNEL_CHECK();
  if (out_where == OX_HOST) OX_TRY(stage_out(out, OX_HOST, d, ob));
  OX_CUDA(cudaStreamSynchronize(g_stream));
  return OX_OK;
}

// ---- enmap.multi_pow (pixell utils.eigpow on axes [0,1]; maps.py:1571): per-pixel symmetric matrix power
namespace {

constexpr int MP_MAXN = 4;

// cyclic Jacobi on a symmetric n x n matrix held in registers: A -> diag(E), V accumulates the rotations
template <int N>
__device__ __forceinline__ void jacobi_eig(double (&A)[N][N], double (&V)[N][N]) {
#pragma unroll
  for (int i = 0; i < N; i++)
#pragma unroll
    for (int j = 0; j < N; j++) V[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 12; sweep++) {
    double off = 0.0, dia = 0.0;
#pragma unroll
    for (int i = 0; i < N; i++) {
      dia += A[i][i] * A[i][i];
#pragma unroll
      for (int j = i + 1; j < N; j++) off += A[i][j] * A[i][j];
    }
    if (!(off > 1e-36 * dia)) break;  // off-diagonal norm below 1e-18 of the diagonal (also stops on NaN / zero)
#pragma unroll
    for (int p = 0; p < N; p++)
#pragma unroll
      for (int q = p + 1; q < N; q++) {
        const double apq = A[p][q];
        if (apq == 0.0) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
        for (int k = 0; k < N; k++) {  // A <- A J
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
#pragma unroll
        for (int k = 0; k < N; k++) {  // A <- J^T A
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
#pragma unroll
        for (int k = 0; k < N; k++) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
}

// out = V diag(f(E)) V^T with f = x^e; for non-integer or negative e eigenvalues that are negative or tiny
// relative to the largest are set to zero first (the eigpow convention the reference relies on for covsqrt)
template <int N>
__global__ void multi_pow_kernel(const double *__restrict__ mat, long long n, double e, int clip, double *__restrict__ out) {
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= n) return;
  double A[N][N], V[N][N], E[N];
#pragma unroll
  for (int i = 0; i < N; i++)
#pragma unroll
    for (int j = 0; j < N; j++) A[i][j] = mat[((long long)i * N + j) * n + p];
  if (N == 1) {
    const double a = A[0][0];
    out[p] = clip ? (a > 0.0 ? pow(fabs(a), e) : 0.0) : pow(a, e);
    return;
  }
#pragma unroll
  for (int i = 0; i < N; i++)
#pragma unroll
    for (int j = i + 1; j < N; j++) A[i][j] = A[j][i] = 0.5 * (A[i][j] + A[j][i]);
  jacobi_eig<N>(A, V);
  double emax = 0.0;
#pragma unroll
  for (int i = 0; i < N; i++) emax = fmax(emax, fabs(A[i][i]));
#pragma unroll
  for (int i = 0; i < N; i++) {
    const double x = A[i][i];
    if (clip) {
      const bool bad = (x < emax * 1e-15 * 100.0) || (x < 2.2250738585072014e-308 * 1e4);
      E[i] = bad ? 0.0 : pow(fabs(x), e);
    } else {
      E[i] = pow(x, e);
    }
  }
#pragma unroll
  for (int i = 0; i < N; i++)
#pragma unroll
    for (int j = 0; j < N; j++) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < N; k++) acc += V[i][k] * E[k] * V[j][k];
      out[((long long)i * N + j) * n + p] = acc;
    }
}

}  // namespace

extern "C" int ox_multi_pow(const double *mat, int n, long long npix, double exponent, int where, double *out, int out_where) {
  OX_REQUIRE(mat && out, "ox_multi_pow: null pointer");
  OX_REQUIRE(n >= 1 && n <= MP_MAXN, "multi_pow supports 1..%d components (got %d)", MP_MAXN, n);
  OX_REQUIRE(npix >= 1, "multi_pow: empty map");
  const size_t bytes = sizeof(double) * (size_t)n * n * npix;
  DevBuf b, o;
  const void *d;
  OX_TRY(stage_in(mat, where, bytes, b, &d));
  double *dst = out;
  if (out_where == OX_HOST) {
    OX_TRY(o.ensure(bytes));
    dst = o.as<double>();
  }
  const int clip = (exponent != floor(exponent) || exponent < 0.0) ? 1 : 0;
  const unsigned g = blocks(npix);
  switch (n) {
    case 1: multi_pow_kernel<1><<<g, ST, 0, g_stream>>>((const double *)d, npix, exponent, clip, dst); break;
    case 2: multi_pow_kernel<2><<<g, ST, 0, g_stream>>>((const double *)d, npix, exponent, clip, dst); break;
    case 3: multi_pow_kernel<3><<<g, ST, 0, g_stream>>>((const double *)d, npix, exponent, clip, dst); break;
    default: multi_pow_kernel<4><<<g, ST, 0, g_stream>>>((const double *)d, npix, exponent, clip, dst); break;
  }
  OX_KERNEL_CHECK();
  if (out_where == OX_HOST) OX_TRY(stage_out(out, OX_HOST, dst, bytes));
  OX_CUDA(cudaStreamSynchronize(g_stream));
  return OX_OK;
}
