/* ==========================================================================
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT PATH.
 *
 * CPU restatement (plain C) of the hot path of
 * osada-yum/CUDA_Fortran_MC_simulation_spin: the checkerboard Monte Carlo
 * sweep and the energy / magnetisation reductions of the Ising 2D, Ising 3D,
 * q-state clock and XY (periodic) modules.  Every function cites the
 * reference file:line it follows.  Array layouts, index arithmetic (1-based,
 * halo "norishiro" cells included), table construction order and comparators
 * are the reference's; uniforms are an INPUT array, exactly as the reference
 * reads them from its cuRAND-filled device arrays.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures
 * (test/check.f90:1-5 is a placeholder) and cannot be built here (CUDA
 * Fortran, no Fortran compiler in the image).  What pins this file instead
 * is (i) the cited source lines, (ii) the physics known-answers in
 * tests/test_oracle_*.py and (iii) oracle-vs-oracle cross checks
 * (tableall vs dual lattice, helical energy vs brute-force bond sums).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 *
 * Build: make -C oracle   (gcc -O3 -fopenmp -shared)
 * ========================================================================== */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "philox.h"

#define ORC_API __attribute__((visibility("default")))

ORC_API int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
ORC_API void orc_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

ORC_API void orc_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    orc_philox4x32_10(ctr, key, out);
}

/* ==========================================================================
 * Ising 2D  (src/ising2d_gpu_m.f90)
 * storage: spins(1-nx : nall+nx), int32, values +1/-1.  C array s[] holds it
 * contiguously; element spins(idx) is s[idx - (1-nx)] = s[idx + nx - 1].
 * ========================================================================== */
#define I2(idx) s[(idx) + nx - 1]

/* update_exparr_ising2d_gpu, src/ising2d_gpu_m.f90:122-131.
 * exparr(-8:8): 1.0 everywhere, then exp(-beta*diff) for diff = 1..8.
 * out[d + 8] = exparr(d). */
ORC_API void orc_ising2d_exparr(double beta, double out[17])
{
    for (int i = 0; i < 17; ++i) out[i] = 1.0;
    for (int diff = 1; diff <= 8; ++diff) out[diff + 8] = exp(-beta * diff);
}

/* update_norishiro_sub, src/ising2d_gpu_m.f90:95-106 */
ORC_API void orc_ising2d_norishiro(int64_t nx, int64_t ny, int32_t *s)
{
    const int64_t nall = nx * ny;
    for (int64_t idx = 1; idx <= nx; ++idx) {
        I2(nall + idx) = I2(idx);
        I2(idx - nx) = I2(nall - nx + idx);
    }
}

/* set_allup_spin, src/ising2d_gpu_m.f90:63-66 (halo included) */
ORC_API void orc_ising2d_set_allup(int64_t nx, int64_t ny, int32_t *s)
{
    for (int64_t j = 0; j < nx * ny + 2 * nx; ++j) s[j] = 1;
}

/* set_random_spin + set_random_spin_sub, src/ising2d_gpu_m.f90:68-84 */
ORC_API void orc_ising2d_set_random(int64_t nx, int64_t ny, int32_t *s, const double *randoms)
{
    const int64_t nall = nx * ny;
    for (int64_t idx = 1; idx <= nall; ++idx) I2(idx) = (randoms[idx - 1] < 0.5) ? 1 : -1;
    orc_ising2d_norishiro(nx, ny, s);
}

/* one colour pass: update_sub + calc_delta_energy,
 * src/ising2d_gpu_m.f90:148-162,191-196.  offset = 1 (odd idx) or 2 (even). */
static void ising2d_pass(int64_t nx, int64_t ny, int32_t *s, const double *randoms,
                         const double exparr[17], int offset)
{
    const int64_t nall = nx * ny;
#pragma omp parallel for schedule(static)
    for (int64_t idx = offset; idx <= nall; idx += 2) {
        int32_t de = 2 * I2(idx) * (I2(idx + 1) + I2(idx - 1) + I2(idx + nx) + I2(idx - nx));
        if (randoms[idx - 1] > exparr[de + 8]) continue;
        I2(idx) = -I2(idx);
    }
}

/* update_ising2d_gpu, src/ising2d_gpu_m.f90:133-147 (randoms already drawn) */
ORC_API void orc_ising2d_update(int64_t nx, int64_t ny, int32_t *s, const double *randoms,
                                const double exparr[17])
{
    ising2d_pass(nx, ny, s, randoms, exparr, 1);
    orc_ising2d_norishiro(nx, ny, s);
    ising2d_pass(nx, ny, s, randoms, exparr, 2);
    orc_ising2d_norishiro(nx, ny, s);
}

/* calc_energy_sum, src/ising2d_gpu_m.f90:198-212 */
ORC_API int64_t orc_ising2d_energy(int64_t nx, int64_t ny, const int32_t *s)
{
    const int64_t nall = nx * ny;
    int64_t res = 0;
#pragma omp parallel for reduction(+ : res) schedule(static)
    for (int64_t i = 1; i <= nall; ++i) res -= (int64_t)(I2(i) * (I2(i + 1) + I2(i + nx)));
    return res;
}

/* calc_magne_sum, src/ising2d_gpu_m.f90:214-228 */
ORC_API int64_t orc_ising2d_magne(int64_t nx, int64_t ny, const int32_t *s)
{
    const int64_t nall = nx * ny;
    int64_t res = 0;
#pragma omp parallel for reduction(+ : res) schedule(static)
    for (int64_t i = 1; i <= nall; ++i) res += I2(i);
    return res;
}
#undef I2

/* ==========================================================================
 * Ising 3D  (src/ising3d_gpu_m.f90)
 * storage: spins(1-nxy : nall+nxy), int32, values 0/1 (1 = up).
 * ========================================================================== */
#define I3(idx) s[(idx) + nxy - 1]

/* update_ws_ising3d_gpu, src/ising3d_gpu_m.f90:138-172.
 * energy_table(0:3,0:1) -> et[s1 + 4*sp]; ws(0:6,0:1) -> ws[S + 7*sp].
 * The reference's loop nest is kept (it overwrites entries repeatedly with
 * identical values; order of the floating-point expression is what matters). */
ORC_API void orc_ising3d_tables(double beta, int64_t et[8], double ws[14])
{
    static const int32_t spin_map[2] = {-1, 1};
    for (int i1 = 0; i1 <= 1; ++i1)
        for (int i2 = 0; i2 <= 1; ++i2)
            for (int i3 = 0; i3 <= 1; ++i3) {
                int s1 = i1 + i2 + i3;
                int32_t sum = spin_map[i1] + spin_map[i2] + spin_map[i3];
                et[s1 + 4 * 0] = -spin_map[0] * sum;
                et[s1 + 4 * 1] = -spin_map[1] * sum;
            }
    for (int i1 = 0; i1 <= 1; ++i1)
        for (int i2 = 0; i2 <= 1; ++i2)
            for (int i3 = 0; i3 <= 1; ++i3) {
                int s1 = i1 + i2 + i3;
                for (int i4 = 0; i4 <= 1; ++i4)
                    for (int i5 = 0; i5 <= 1; ++i5)
                        for (int i6 = 0; i6 <= 1; ++i6) {
                            int s2 = i4 + i5 + i6;
                            int64_t e1 = et[s1 + 4 * 0] + et[s2 + 4 * 0];
                            int64_t e2 = et[s1 + 4 * 1] + et[s2 + 4 * 1];
                            ws[s1 + s2 + 7 * 0] = fmin(1.0, exp(-beta * (double)(e2 - e1)));
                            ws[s1 + s2 + 7 * 1] = fmin(1.0, exp(-beta * (double)(e1 - e2)));
                        }
            }
}

/* update_norishiro_sub, src/ising3d_gpu_m.f90:111-122 */
ORC_API void orc_ising3d_norishiro(int64_t nx, int64_t ny, int64_t nz, int32_t *s)
{
    const int64_t nxy = nx * ny, nall = nxy * nz;
    for (int64_t idx = 1; idx <= nxy; ++idx) {
        I3(nall + idx) = I3(idx);
        I3(idx - nxy) = I3(nall - nxy + idx);
    }
}

/* set_allup_spin, src/ising3d_gpu_m.f90:79-82 */
ORC_API void orc_ising3d_set_allup(int64_t nx, int64_t ny, int64_t nz, int32_t *s)
{
    const int64_t nxy = nx * ny;
    for (int64_t j = 0; j < nxy * nz + 2 * nxy; ++j) s[j] = 1;
}

/* set_random_spin + sub, src/ising3d_gpu_m.f90:84-100 */
ORC_API void orc_ising3d_set_random(int64_t nx, int64_t ny, int64_t nz, int32_t *s,
                                    const double *randoms)
{
    const int64_t nxy = nx * ny, nall = nxy * nz;
    for (int64_t idx = 1; idx <= nall; ++idx) I3(idx) = (randoms[idx - 1] < 0.5) ? 1 : 0;
    orc_ising3d_norishiro(nx, ny, nz, s);
}

/* update_sub, src/ising3d_gpu_m.f90:189-206 */
static void ising3d_pass(int64_t nx, int64_t ny, int64_t nz, int32_t *s, const double *randoms,
                         const double ws[14], int offset)
{
    const int64_t nxy = nx * ny, nall = nxy * nz;
#pragma omp parallel for schedule(static)
    for (int64_t idx = offset; idx <= nall; idx += 2) {
        int32_t sum_spin = I3(idx - 1) + I3(idx + 1) + I3(idx - nx) + I3(idx + nx) +
                           I3(idx - nxy) + I3(idx + nxy);
        if (randoms[idx - 1] > ws[sum_spin + 7 * I3(idx)]) continue;
        I3(idx) = 1 - I3(idx);
    }
}

/* update_ising3d_gpu, src/ising3d_gpu_m.f90:174-188 */
ORC_API void orc_ising3d_update(int64_t nx, int64_t ny, int64_t nz, int32_t *s,
                                const double *randoms, const double ws[14])
{
    ising3d_pass(nx, ny, nz, s, randoms, ws, 1);
    orc_ising3d_norishiro(nx, ny, nz, s);
    ising3d_pass(nx, ny, nz, s, randoms, ws, 2);
    orc_ising3d_norishiro(nx, ny, nz, s);
}

/* calc_energy_sum, src/ising3d_gpu_m.f90:239-257 */
ORC_API int64_t orc_ising3d_energy(int64_t nx, int64_t ny, int64_t nz, const int32_t *s,
                                   const int64_t et[8])
{
    const int64_t nxy = nx * ny, nall = nxy * nz;
    int64_t res = 0;
#pragma omp parallel for reduction(+ : res) schedule(static)
    for (int64_t i = 1; i <= nall; ++i)
        res += et[(I3(i + 1) + I3(i + nx) + I3(i + nxy)) + 4 * I3(i)];
    return res;
}

/* calc_magne_sum, src/ising3d_gpu_m.f90:259-276 */
ORC_API int64_t orc_ising3d_magne(int64_t nx, int64_t ny, int64_t nz, const int32_t *s)
{
    const int64_t nxy = nx * ny, nall = nxy * nz;
    int64_t res = 0;
#pragma omp parallel for reduction(+ : res) schedule(static)
    for (int64_t i = 1; i <= nall; ++i) res += I3(i);
    return 2 * res - nall;
}
#undef I3

/* ==========================================================================
 * The drivers' statistics accumulator `variance_covariance_kahan`
 * (app/ising3d_gpu_relaxation.f90:4,17,36,46-55).  It lives in the external,
 * un-vendored dependency osada-yum/Numerical_utilities (fpm.toml:14, no
 * revision pinned), so it is restated from its use: add_data(v1, v2) keeps
 * Kahan-compensated running sums of v1, v2, v1^2, v2^2, v1 v2; num_sample,
 * mean1/2 = sum / n, square_mean1/2 = sum of squares / n,
 * var1/2 = n/(n-1) (square_mean - mean^2), cov = n/(n-1) (mean_v1v2 - mean1 mean2)
 * (unbiased estimators; 0 for n = 1).  PARITY UNPINNED for the estimator
 * convention (biased vs unbiased) -- the sums themselves are unambiguous.
 * st[0..4] = sums, st[5..9] = compensations.
 * ========================================================================== */
ORC_API void orc_kahan_add_data(double st[10], double v1, double v2)
{
    const double x[5] = {v1, v2, v1 * v1, v2 * v2, v1 * v2};
    for (int k = 0; k < 5; ++k) {
        const double y = x[k] - st[5 + k];
        const double t = st[k] + y;
        st[5 + k] = (t - st[k]) - y;
        st[k] = t;
    }
}

/* out: num_sample, mean1, mean2, square_mean1, square_mean2, var1, var2, cov */
ORC_API void orc_kahan_results(const double st[10], int64_t n_sample, double out[8])
{
    const double n = (double)n_sample;
    const double mean1 = st[0] / n, mean2 = st[1] / n, sq1 = st[2] / n, sq2 = st[3] / n, m12 = st[4] / n;
    const double f = n_sample > 1 ? n / (n - 1.0) : 0.0;
    out[0] = n; out[1] = mean1; out[2] = mean2; out[3] = sq1; out[4] = sq2;
    out[5] = f * (sq1 - mean1 * mean1); out[6] = f * (sq2 - mean2 * mean2); out[7] = f * (m12 - mean1 * mean2);
}

/* ==========================================================================
 * Heat-bath for Ising 2D / 3D.  NOT IN THE REFERENCE: named
 * by the north star only.  Defined here with the reference's conventions:
 *   p_up(S) = 1 / (1 + exp(-2*beta*h)),  h = 2S - z   (z = 4 or 6, S = number
 *   of up neighbours);  new spin = up iff u <= p_up(S), independent of the
 *   current spin.  Same storage, colouring, halo and uniform array as the
 *   Metropolis passes above.  PARITY UNPINNED (no reference symbol).
 * ========================================================================== */
ORC_API void orc_heatbath_table(double beta, int z, double *pup /* z+1 */)
{
    for (int S = 0; S <= z; ++S) pup[S] = 1.0 / (1.0 + exp(-2.0 * beta * (double)(2 * S - z)));
}

ORC_API void orc_ising2d_update_heatbath(int64_t nx, int64_t ny, int32_t *s,
                                         const double *randoms, const double pup[5])
{
    const int64_t nall = nx * ny;
#define I2(idx) s[(idx) + nx - 1]
    for (int offset = 1; offset <= 2; ++offset) {
#pragma omp parallel for schedule(static)
        for (int64_t idx = offset; idx <= nall; idx += 2) {
            int32_t sum = I2(idx + 1) + I2(idx - 1) + I2(idx + nx) + I2(idx - nx); /* -4..4 */
            int S = (sum + 4) / 2;
            I2(idx) = (randoms[idx - 1] <= pup[S]) ? 1 : -1;
        }
        orc_ising2d_norishiro(nx, ny, s);
    }
#undef I2
}

ORC_API void orc_ising3d_update_heatbath(int64_t nx, int64_t ny, int64_t nz, int32_t *s,
                                         const double *randoms, const double pup[7])
{
    const int64_t nxy = nx * ny, nall = nxy * nz;
#define I3(idx) s[(idx) + nxy - 1]
    for (int offset = 1; offset <= 2; ++offset) {
#pragma omp parallel for schedule(static)
        for (int64_t idx = offset; idx <= nall; idx += 2) {
            int S = I3(idx - 1) + I3(idx + 1) + I3(idx - nx) + I3(idx + nx) + I3(idx - nxy) +
                    I3(idx + nxy);
            I3(idx) = (randoms[idx - 1] <= pup[S]) ? 1 : 0;
        }
        orc_ising3d_norishiro(nx, ny, nz, s);
    }
#undef I3
}

/* ==========================================================================
 * Periodic Ising 2D / 3D (torus).  NOT IN THE REFERENCE, whose modules are helical only and valid for odd nx
 * (SURVEY.md Q1): this is the "1024^3 periodic" / "1024^2 periodic" / "65536^2 periodic" input of BASELINE.md
 * section 2 and SURVEY.md 8(d), defined with the reference's own update rule, tables, value conventions and
 * observables (src/ising3d_gpu_m.f90:189-206,239-276; src/ising2d_gpu_m.f90:148-162,191-228) and true periodic
 * neighbours; colour = (x + y + z) & 1, colour 0 first (the reference's odd 1-based indices).  All extents even.
 * storage: s[x + nx (y + ny z)], no halo; 3D values 0/1, 2D values -1/+1.  PARITY UNPINNED (no reference symbol).
 * method: 0 = Metropolis (w = ws(0:6,0:1) in 3D, exparr(-8:8) in 2D), 1 = heat-bath (w = p_up(0:z)).
 * ========================================================================== */
ORC_API void orc_isingp_update(int ndim, int64_t nx, int64_t ny, int64_t nz, int32_t *s, const double *randoms,
                               const double *w, int method)
{
    if (ndim == 2) nz = 1;
    const int64_t nxy = nx * ny;
    for (int colour = 0; colour < 2; ++colour) {
#pragma omp parallel for schedule(static)
        for (int64_t zy = 0; zy < nz * ny; ++zy) {
            const int64_t z = zy / ny, y = zy % ny;
            const int64_t yp = (y + 1) % ny, ym = (y + ny - 1) % ny, zp = (z + 1) % nz, zm = (z + nz - 1) % nz;
            for (int64_t x = (y + z + colour) & 1; x < nx; x += 2) {
                const int64_t xp = (x + 1) % nx, xm = (x + nx - 1) % nx;
                const int64_t i = x + nx * y + nxy * z;
                int32_t sum = s[xm + nx * y + nxy * z] + s[xp + nx * y + nxy * z] + s[x + nx * ym + nxy * z] +
                              s[x + nx * yp + nxy * z];
                if (ndim == 3) {
                    sum += s[x + nx * y + nxy * zm] + s[x + nx * y + nxy * zp];
                    if (method == 0) {
                        if (randoms[i] > w[sum + 7 * s[i]]) continue;   /* src/ising3d_gpu_m.f90:203-204 */
                        s[i] = 1 - s[i];
                    } else {
                        s[i] = (randoms[i] <= w[sum]) ? 1 : 0;
                    }
                } else {
                    if (method == 0) {
                        const int32_t de = 2 * s[i] * sum;                  /* src/ising2d_gpu_m.f90:195 */
                        if (randoms[i] <= w[de + 8]) s[i] = -s[i];          /* :159-160 */
                    } else {
                        s[i] = (randoms[i] <= w[(sum + 4) / 2]) ? 1 : -1;
                    }
                }
            }
        }
    }
}

/* E as the reference sums it: 3D  sum_i energy_table(s(+x) + s(+y) + s(+z), s(i)), 2D  -sum_i s(i) (s(+x) + s(+y));
 * M = 2 sum(s) - nall in 3D, sum(s) in 2D */
ORC_API void orc_isingp_energy_magne(int ndim, int64_t nx, int64_t ny, int64_t nz, const int32_t *s, int64_t *e_out,
                                     int64_t *m_out)
{
    if (ndim == 2) nz = 1;
    const int64_t nxy = nx * ny, nall = nxy * nz;
    int64_t et[8];
    double ws[14];
    orc_ising3d_tables(1.0, et, ws);
    int64_t e = 0, m = 0;
#pragma omp parallel for reduction(+ : e, m) schedule(static)
    for (int64_t i = 0; i < nall; ++i) {
        const int64_t x = i % nx, y = (i / nx) % ny, z = i / nxy;
        const int64_t ixp = (x + 1) % nx + nx * y + nxy * z, iyp = x + nx * ((y + 1) % ny) + nxy * z;
        if (ndim == 3) {
            const int64_t izp = x + nx * y + nxy * ((z + 1) % nz);
            e += et[(s[ixp] + s[iyp] + s[izp]) + 4 * s[i]];
        } else {
            e -= (int64_t)s[i] * (s[ixp] + s[iyp]);
        }
        m += s[i];
    }
    *e_out = e;
    *m_out = ndim == 3 ? 2 * m - nall : m;
}

/* ==========================================================================
 * q-state clock, helical  (src/clock_gpu_m.f90, src/clock_gpu_multi_m.f90)
 * storage: spins(1-nx : nall+nx) int32 in 0..q-1  (one such array per replica)
 * ========================================================================== */
#define CK(idx) s[(idx) + nx - 1]

/* init (spin_magne) src/clock_gpu_m.f90:60,66-72 and update_ws :105-146.
 *   magne[k]           = cos(pi_state_inv * k)
 *   etab[i + q*(j + q*c)] = energy_table(i, j, c)
 *   ws[i + q*(j + q*(k + q*(l + q*(cb + q*ca))))] = ws(i,j,k,l,cb,ca)       */
ORC_API void orc_clock_tables(int32_t q, double beta, double *magne, double *etab, double *ws)
{
    const double pi = 4 * atan(1.0);
    const double pi_state_inv = 2 * pi / q;
    for (int i = 0; i < q; ++i) magne[i] = cos(pi_state_inv * i);
    for (int c = 0; c < q; ++c)
        for (int j = 0; j < q; ++j)
            for (int i = 0; i < q; ++i) {
                /* calc_local_energy :142-145: - sum(cos(pi_state_inv * [i-c, j-c])) */
                double a = cos(pi_state_inv * (i - c));
                double b = cos(pi_state_inv * (j - c));
                etab[i + q * (j + q * c)] = -(a + b);
            }
    if (!ws) return;
#define ET(i, j, c) etab[(i) + q * ((j) + q * (c))]
    for (int ca = 0; ca < q; ++ca)
        for (int cb = 0; cb < q; ++cb)
            for (int l = 0; l < q; ++l)
                for (int k = 0; k < q; ++k)
                    for (int j = 0; j < q; ++j)
                        for (int i = 0; i < q; ++i) {
                            double de = (ET(i, j, ca) + ET(k, l, ca)) - (ET(i, j, cb) + ET(k, l, cb));
                            size_t at = (size_t)i +
                                        (size_t)q * (j + (size_t)q * (k + (size_t)q * (l + (size_t)q * (cb + (size_t)q * ca))));
                            ws[at] = (de <= 0.0) ? 1.0 : exp(-beta * de);
                        }
#undef ET
}

/* update_norishiro_sub, src/clock_gpu_m.f90:157-168 */
ORC_API void orc_clock_norishiro(int64_t nx, int64_t ny, int32_t *s)
{
    const int64_t nall = nx * ny;
    for (int64_t idx = 1; idx <= nx; ++idx) {
        CK(nall + idx) = CK(idx);
        CK(idx - nx) = CK(nall - nx + idx);
    }
}

/* set_random_spin_sub, src/clock_gpu_m.f90:94-104: spins = floor(r * q).
 * clamp: r == 1.0 would give q (out of range in the reference, quirk Q4);
 * the restatement clamps to q-1 and the product does the same. */
ORC_API void orc_clock_set_random(int64_t nx, int64_t ny, int32_t q, int32_t *s,
                                  const double *randoms)
{
    const int64_t nall = nx * ny;
    for (int64_t idx = 1; idx <= nall; ++idx) {
        int32_t v = (int32_t)floor(randoms[idx - 1] * q);
        if (v >= q) v = q - 1;
        CK(idx) = v;
    }
    orc_clock_norishiro(nx, ny, s);
}

/* update_clock_gpu + update_sub, src/clock_gpu_m.f90:183-216;
 * strict != 0 selects the comparator of clock_gpu_multi_m.f90:230-235
 * (reject when r >= w) instead of clock_gpu_m.f90:212 (reject when r > w). */
ORC_API void orc_clock_update(int64_t nx, int64_t ny, int32_t q, int32_t *s, const double *randoms,
                              const double *next_states, const double *ws, int strict)
{
    const int64_t nall = nx * ny;
    for (int offset = 1; offset <= 2; ++offset) {
#pragma omp parallel for schedule(static)
        for (int64_t idx = offset; idx <= nall; idx += 2) {
            int32_t nxt = (int32_t)floor(next_states[idx - 1] * q);
            if (nxt >= q) nxt = q - 1; /* quirk Q4 clamp */
            size_t at = (size_t)CK(idx + nx) +
                        (size_t)q * (CK(idx - nx) +
                                     (size_t)q * (CK(idx - 1) +
                                                  (size_t)q * (CK(idx + 1) +
                                                               (size_t)q * (CK(idx) + (size_t)q * nxt))));
            double w = ws[at];
            double r = randoms[idx - 1];
            if (strict ? (r >= w) : (r > w)) continue;
            CK(idx) = nxt;
        }
        orc_clock_norishiro(nx, ny, s);
    }
}

/* calc_energy_sum, src/clock_gpu_m.f90:245-262 (serial order i = 1..nall) */
ORC_API double orc_clock_energy(int64_t nx, int64_t ny, int32_t q, const int32_t *s,
                                const double *etab)
{
    const int64_t nall = nx * ny;
    double res = 0.0;
    for (int64_t i = 1; i <= nall; ++i)
        res += etab[CK(i - nx) + q * (CK(i - 1) + q * CK(i))];
    return res;
}

/* calc_magne_sum, src/clock_gpu_m.f90:264-280 */
ORC_API double orc_clock_magne(int64_t nx, int64_t ny, int32_t q, const int32_t *s,
                               const double *magne)
{
    (void)q;
    const int64_t nall = nx * ny;
    double res = 0.0;
    for (int64_t i = 1; i <= nall; ++i) res += magne[CK(i)];
    return res;
}

/* integer form of the two observables (what the product accumulates exactly):
 * hist[c] = #{i : s(i) = c};  pair[a + q*b] = #{bonds (i, i-1) and (i, i-nx)
 * with neighbour state a and centre state b}.  E = sum pair * (-cos(a-b)),
 * M = sum hist * cos. */
ORC_API void orc_clock_histograms(int64_t nx, int64_t ny, int32_t q, const int32_t *s,
                                  int64_t *hist, int64_t *pair)
{
    const int64_t nall = nx * ny;
    memset(hist, 0, sizeof(int64_t) * q);
    memset(pair, 0, sizeof(int64_t) * q * q);
    for (int64_t i = 1; i <= nall; ++i) {
        hist[CK(i)]++;
        pair[CK(i - nx) + q * CK(i)]++;
        pair[CK(i - 1) + q * CK(i)]++;
    }
}
#undef CK

/* ==========================================================================
 * 6-state clock, periodic "tableall" (src/clock/clock_tableall_gpu_m.f90) and
 * its dual-lattice twin (src/clock/clock_dual_lattice_tableall_m.f90).
 * storage: sixclock(nx, ny) int32, Fortran column-major: c[(x-1) + nx*(y-1)].
 * rnds(2, nx, ny): rnds[(j-1) + 2*((x-1) + nx*(y-1))].
 * q is a parameter here (mstate = 6 in the reference).
 * ========================================================================== */
/* state_to_magne :26, state_center_right_up_to_energy :27-33.
 * e3[c + q*(a + q*b)] = table(c, a, b)  with table(c, r, u) as the reference
 * indexes it.  The reshape at :28-33 runs global_c fastest, then global_u,
 * then global_r, so element (c, a, b) holds the value computed with
 * global_u = a, global_r = b:  -cos((a-c)*psi) - cos((b-c)*psi).           */
ORC_API void orc_tableall_tables(int32_t q, double beta, double *magne, double *e3, double *prob)
{
    const double pi = 4 * atan(1.0);
    const double psi = 2 * pi / q;
    for (int c = 0; c < q; ++c) magne[c] = cos(c * psi);
    for (int b = 0; b < q; ++b)          /* global_r */
        for (int a = 0; a < q; ++a)      /* global_u */
            for (int c = 0; c < q; ++c)  /* global_c */
                e3[c + q * (a + q * b)] = -cos((a - c) * psi) - cos((b - c) * psi);
    if (!prob) return;
#define E3(c, r, u) e3[(c) + q * ((r) + q * (u))]
    /* init_sixclock :66-86;  prob[c + q*(n + q*(r + q*(u + q*(l + q*d))))] */
    for (int d = 0; d < q; ++d)
        for (int l = 0; l < q; ++l)
            for (int u = 0; u < q; ++u)
                for (int r = 0; r < q; ++r)
                    for (int n = 0; n < q; ++n)
                        for (int c = 0; c < q; ++c) {
                            double de = E3(n, r, u) - E3(c, r, u) + E3(n, l, d) - E3(c, l, d);
                            size_t at = (size_t)c +
                                        (size_t)q * (n + (size_t)q * (r + (size_t)q * (u + (size_t)q * (l + (size_t)q * d))));
                            prob[at] = (de <= 0.0) ? 1.0 : exp(-beta * de);
                        }
#undef E3
}

/* clock_simple_gpu_m: no tables; delta_e accumulated over nearest_spins(1:4) = right, left, up, down
 * (src/clock/clock_simple_gpu_m.f90:83-113).  Tabulated here in the states_to_prob index order so that
 * orc_tableall_update can consume it: prob[c + q*(n + q*(r + q*(u + q*(l + q*d))))]. */
ORC_API void orc_clock_simple_prob(int32_t q, double beta, double *prob)
{
    const double pi = 4 * atan(1.0);
    const double psi = 2 * pi / q;
    for (int d = 0; d < q; ++d)
        for (int l = 0; l < q; ++l)
            for (int u = 0; u < q; ++u)
                for (int r = 0; r < q; ++r)
                    for (int n = 0; n < q; ++n)
                        for (int c = 0; c < q; ++c) {
                            const int nb[4] = {r, l, u, d};
                            double de = 0.0;
                            for (int i = 0; i < 4; ++i)
                                de = de + (-cos((nb[i] - n) * psi) + cos((nb[i] - c) * psi));
                            size_t at = (size_t)c +
                                        (size_t)q * (n + (size_t)q * (r + (size_t)q * (u + (size_t)q * (l + (size_t)q * d))));
                            prob[at] = (de > 0) ? exp(-beta * de) : 1.0;
                        }
}

#define TA(x, y) c[((x)-1) + nx * ((y)-1)]
#define RN(j, x, y) rnds[((j)-1) + 2 * (((x)-1) + nx * ((y)-1))]

/* update_metropolis + update_sub, src/clock/clock_tableall_gpu_m.f90:94-152 */
ORC_API void orc_tableall_update(int64_t nx, int64_t ny, int32_t q, int32_t *c, const double *rnds,
                                 const double *prob)
{
    for (int parity = 0; parity <= 1; ++parity) {
#pragma omp parallel for schedule(static)
        for (int64_t y = 1; y <= ny; ++y)
            for (int64_t x = 1; x <= nx; ++x) {
                if (((x + y) & 1) != parity) continue;
                int64_t rx = x + 1; if (rx > nx) rx = 1;
                int64_t lx = x - 1; if (lx < 1) lx = nx;
                int64_t uy = y + 1; if (uy > ny) uy = 1;
                int64_t dy = y - 1; if (dy < 1) dy = ny;
                int32_t n1 = TA(rx, y), n2 = TA(lx, y), n3 = TA(x, uy), n4 = TA(x, dy);
                int32_t ns = TA(x, y) + (int32_t)ceil(RN(1, x, y) * (q - 1));
                if (ns >= q) ns -= q;
                size_t at = (size_t)TA(x, y) +
                            (size_t)q * (ns + (size_t)q * (n1 + (size_t)q * (n3 + (size_t)q * (n2 + (size_t)q * n4))));
                if (RN(2, x, y) <= prob[at]) TA(x, y) = ns;
            }
    }
}

/* calc_magne :155-165 (returns per-site value, like the reference) */
ORC_API double orc_tableall_magne(int64_t nx, int64_t ny, int32_t q, const int32_t *c,
                                  const double *magne)
{
    (void)q;
    double res = 0.0;
    for (int64_t y = 1; y <= ny; ++y)
        for (int64_t x = 1; x <= nx; ++x) res += magne[TA(x, y)];
    return res * (1.0 / (double)(nx * ny));
}

/* calc_energy :167-181 */
ORC_API double orc_tableall_energy(int64_t nx, int64_t ny, int32_t q, const int32_t *c,
                                   const double *e3)
{
    double res = 0.0;
    for (int64_t y = 1; y <= ny; ++y)
        for (int64_t x = 1; x <= nx; ++x) {
            int64_t rx = x + 1; if (rx > nx) rx = 1;
            int64_t uy = y + 1; if (uy > ny) uy = 1;
            res += e3[TA(x, y) + q * (TA(rx, y) + q * TA(x, uy))];
        }
    return res * (1.0 / (double)(nx * ny));
}

/* integer observables for the periodic lattice: hist[c], pair[a + q*b] over
 * bonds (x,y)-(x+1,y) and (x,y)-(x,y+1) with centre state a, neighbour b. */
ORC_API void orc_tableall_histograms(int64_t nx, int64_t ny, int32_t q, const int32_t *c,
                                     int64_t *hist, int64_t *pair)
{
    memset(hist, 0, sizeof(int64_t) * q);
    memset(pair, 0, sizeof(int64_t) * q * q);
    for (int64_t y = 1; y <= ny; ++y)
        for (int64_t x = 1; x <= nx; ++x) {
            int64_t rx = x + 1; if (rx > nx) rx = 1;
            int64_t uy = y + 1; if (uy > ny) uy = 1;
            hist[TA(x, y)]++;
            pair[TA(x, y) + q * TA(rx, y)]++;
            pair[TA(x, y) + q * TA(x, uy)]++;
        }
}

/* dual lattice: update_sub, src/clock/clock_dual_lattice_tableall_m.f90:110-155.
 * even(nx/2, ny), odd(nx/2, ny): e[(X-1) + nh*(y-1)].  The host loop
 * (:96-104) updates `even` with parity_bit 0 then `odd` with parity_bit 1. */
static void dual_pass(int64_t nx, int64_t ny, int32_t q, int32_t *upd, const int32_t *nea,
                      const double *rnds, const double *prob, int parity_bit)
{
    const int64_t nh = nx / 2;
#define UP(X, y) upd[((X)-1) + nh * ((y)-1)]
#define NE(X, y) nea[((X)-1) + nh * ((y)-1)]
#pragma omp parallel for schedule(static)
    for (int64_t y = 1; y <= ny; ++y)
        for (int64_t x = 1; x <= nh; ++x) {
            int64_t rx = x + ((y + 1 + parity_bit) & 1); if (rx > nh) rx = 1;
            int64_t lx = x - ((y + parity_bit) & 1);     if (lx < 1) lx = nh;
            int64_t uy = y + 1; if (uy > ny) uy = 1;
            int64_t dy = y - 1; if (dy < 1) dy = ny;
            int32_t n1 = NE(rx, y), n2 = NE(lx, y), n3 = NE(x, uy), n4 = NE(x, dy);
            int64_t ax = 2 * x - ((y + parity_bit) & 1);
            int32_t ns = UP(x, y) + (int32_t)ceil(RN(1, ax, y) * (q - 1));
            if (ns >= q) ns -= q;
            size_t at = (size_t)UP(x, y) +
                        (size_t)q * (ns + (size_t)q * (n1 + (size_t)q * (n3 + (size_t)q * (n2 + (size_t)q * n4))));
            if (RN(2, ax, y) <= prob[at]) UP(x, y) = ns;
        }
#undef UP
#undef NE
}

ORC_API void orc_dual_update(int64_t nx, int64_t ny, int32_t q, int32_t *even, int32_t *odd,
                             const double *rnds, const double *prob)
{
    dual_pass(nx, ny, q, even, odd, rnds, prob, 0);
    dual_pass(nx, ny, q, odd, even, rnds, prob, 1);
}

/* calc_energy of the dual lattice, :175-201 (per-site) */
ORC_API double orc_dual_energy(int64_t nx, int64_t ny, int32_t q, const int32_t *even,
                               const int32_t *odd, const double *e3)
{
    const int64_t nh = nx / 2;
    double res = 0.0;
#define EV(X, y) even[((X)-1) + nh * ((y)-1)]
#define OD(X, y) odd[((X)-1) + nh * ((y)-1)]
    for (int64_t y = 1; y <= ny; ++y) {
        int64_t uy = y + 1; if (uy > ny) uy = 1;
        int64_t add_even = (y + 1) & 1, add_odd = y & 1;
        for (int64_t x = 1; x <= nh; ++x) {
            int64_t rx = x + add_even; if (rx > nh) rx = 1;
            res += e3[EV(x, y) + q * (OD(rx, y) + q * OD(x, uy))];
            rx = x + add_odd; if (rx > nh) rx = 1;
            res += e3[OD(x, y) + q * (EV(rx, y) + q * EV(x, uy))];
        }
    }
#undef EV
#undef OD
    return res * (1.0 / (double)(nx * ny));
}
#undef TA
#undef RN

/* ==========================================================================
 * XY 2D periodic  (src/xy2d_periodic_gpu_m.f90)
 * storage: spins(0:nx+1, 0:ny+1, 1:2) real64, column-major:
 *   S(x, y, k) = sp[x + (nx+2)*(y + (ny+2)*(k-1))]
 * randoms(nx, ny), candidates(nx, ny): r[(x-1) + nx*(y-1)]
 * ========================================================================== */
#define XS(x, y, k) sp[(x) + (nx + 2) * ((y) + (ny + 2) * ((k)-1))]
#define XR(a, x, y) a[((x)-1) + nx * ((y)-1)]

/* update_norishiro_updown_sub / leftright_sub, :304-326 (corners untouched) */
ORC_API void orc_xy_norishiro(int64_t nx, int64_t ny, double *sp)
{
    for (int64_t x = 1; x <= nx; ++x)
        for (int k = 1; k <= 2; ++k) {
            XS(x, ny + 1, k) = XS(x, 1, k);
            XS(x, 0, k) = XS(x, ny, k);
        }
    for (int64_t y = 1; y <= ny; ++y)
        for (int k = 1; k <= 2; ++k) {
            XS(0, y, k) = XS(nx, y, k);
            XS(nx + 1, y, k) = XS(1, y, k);
        }
}

/* set_allup_spin, :86-101 */
ORC_API void orc_xy_set_allup(int64_t nx, int64_t ny, double *sp)
{
    for (int64_t y = 1; y <= ny; ++y)
        for (int64_t x = 1; x <= nx; ++x) {
            XS(x, y, 1) = 1.0;
            XS(x, y, 2) = 0.0;
        }
    orc_xy_norishiro(nx, ny, sp);
}

/* set_random_spin + sub, :105-122.  The reference does NOT refresh the halo
 * here (quirk Q6); refresh != 0 adds it (what the product's halo-free storage
 * is equivalent to). */
ORC_API void orc_xy_set_random(int64_t nx, int64_t ny, double *sp, const double *randoms,
                               int refresh)
{
    const double pi = 4 * atan(1.0);
    for (int64_t y = 1; y <= ny; ++y)
        for (int64_t x = 1; x <= nx; ++x) {
            XS(x, y, 1) = cos(2 * pi * XR(randoms, x, y));
            XS(x, y, 2) = sin(2 * pi * XR(randoms, x, y));
        }
    if (refresh) orc_xy_norishiro(nx, ny, sp);
}

/* update_sub + calc_delta_energy, :368-397; one colour (offset 0 or 1) */
static void xy_pass(int64_t nx, int64_t ny, double *sp, double beta, const double *randoms,
                    const double *candidates, int offset)
{
    const double pi = 4 * atan(1.0);
    const int64_t nall = nx * ny;
#pragma omp parallel for schedule(static)
    for (int64_t t = 1; t <= nall / 2; ++t) {
        int64_t idx = 2 * t - 1;
        int64_t y = (idx - 1) / nx + 1;
        int64_t x = idx - (y - 1) * nx + ((((y + offset) & 1) == 1) ? 0 : 1);
        double c1 = cos(2 * pi * XR(candidates, x, y));
        double c2 = sin(2 * pi * XR(candidates, x, y));
        double d1 = c1 - XS(x, y, 1), d2 = c2 - XS(x, y, 2);
        double n1 = XS(x + 1, y, 1) + XS(x - 1, y, 1) + XS(x, y + 1, 1) + XS(x, y - 1, 1);
        double n2 = XS(x + 1, y, 2) + XS(x - 1, y, 2) + XS(x, y + 1, 2) + XS(x, y - 1, 2);
        double de = -(d1 * n1 + d2 * n2);
        if (XR(randoms, x, y) > exp(-beta * de)) continue;
        XS(x, y, 1) = c1;
        XS(x, y, 2) = c2;
    }
}

/* update_xy2d_gpu, :353-367 */
ORC_API void orc_xy_update(int64_t nx, int64_t ny, double *sp, double beta, const double *randoms,
                           const double *candidates)
{
    xy_pass(nx, ny, sp, beta, randoms, candidates, 0);
    orc_xy_norishiro(nx, ny, sp);
    xy_pass(nx, ny, sp, beta, randoms, candidates, 1);
    orc_xy_norishiro(nx, ny, sp);
}

/* over_relaxation_sub, :418-439 */
static void xy_or_pass(int64_t nx, int64_t ny, double *sp, int offset)
{
    const int64_t nall = nx * ny;
#pragma omp parallel for schedule(static)
    for (int64_t t = 1; t <= nall / 2; ++t) {
        int64_t idx = 2 * t - 1;
        int64_t y = (idx - 1) / nx + 1;
        int64_t x = idx - (y - 1) * nx + ((((y + offset) & 1) == 1) ? 0 : 1);
        double h1 = XS(x - 1, y, 1) + XS(x + 1, y, 1) + XS(x, y - 1, 1) + XS(x, y + 1, 1);
        double h2 = XS(x - 1, y, 2) + XS(x + 1, y, 2) + XS(x, y - 1, 2) + XS(x, y + 1, 2);
        double inv = 1 / hypot(h1, h2);
        h1 = h1 * inv;
        h2 = h2 * inv;
        double dot2 = 2 * (h1 * XS(x, y, 1) + h2 * XS(x, y, 2));
        double s1 = dot2 * h1 - XS(x, y, 1);
        double s2 = dot2 * h2 - XS(x, y, 2);
        double rabs = hypot(s1, s2);
        XS(x, y, 1) = s1 / rabs;
        XS(x, y, 2) = s2 / rabs;
    }
}

/* update_over_relaxation_xy2d_gpu, :400-416 */
ORC_API void orc_xy_over_relaxation(int64_t nx, int64_t ny, double *sp, int32_t n_steps)
{
    for (int i = 0; i < n_steps; ++i) {
        xy_or_pass(nx, ny, sp, 0);
        orc_xy_norishiro(nx, ny, sp);
        xy_or_pass(nx, ny, sp, 1);
        orc_xy_norishiro(nx, ny, sp);
    }
}

/* calc_energy_sum_xy2d_gpu_sub, :496-508 (serial order idx = 1..nall) */
ORC_API double orc_xy_energy(int64_t nx, int64_t ny, const double *sp)
{
    double res = 0.0;
    for (int64_t y = 1; y <= ny; ++y)
        for (int64_t x = 1; x <= nx; ++x) {
            res -= XS(x, y, 1) * (XS(x + 1, y, 1) + XS(x, y + 1, 1));
            res -= XS(x, y, 2) * (XS(x + 1, y, 2) + XS(x, y + 1, 2));
        }
    return res;
}

/* calc_magne_sum (:510-521) for comp = 1, calc_magne_y_sum (:523-534) for 2 */
ORC_API double orc_xy_magne(int64_t nx, int64_t ny, const double *sp, int comp)
{
    double res = 0.0;
    for (int64_t y = 1; y <= ny; ++y)
        for (int64_t x = 1; x <= nx; ++x) res += XS(x, y, comp);
    return res;
}

/* calc_autocorrelation_sum_xy2d_gpu_sub, :536-549 */
ORC_API double orc_xy_autocorrelation(int64_t nx, int64_t ny, const double *sp, const double *sp0)
{
    double res = 0.0;
    for (int64_t y = 1; y <= ny; ++y)
        for (int64_t x = 1; x <= nx; ++x)
            for (int k = 1; k <= 2; ++k)
                res += XS(x, y, k) * sp0[(x) + (nx + 2) * ((y) + (ny + 2) * ((k)-1))];
    return res;
}

/* calc_correlation_sum_xy2d_gpu_sub, :551-567 */
ORC_API double orc_xy_correlation(int64_t nx, int64_t ny, const double *sp)
{
    double res = 0.0;
    for (int64_t y = 1; y <= ny; ++y)
        for (int64_t x = 1; x <= nx; ++x) {
            int64_t nyy = y + (ny / 2 - 1); if (nyy > ny) nyy -= ny;
            int64_t nxx = x + (nx / 2 - 1); if (nxx > nx) nxx -= nx;
            res += XS(x, y, 1) * XS(nxx, nyy, 1);
            res += XS(x, y, 2) * XS(nxx, nyy, 2);
        }
    return res;
}

/* rotate_whole_spin_theta_sub, :281-293 */
ORC_API void orc_xy_rotate(int64_t nx, int64_t ny, double *sp, double theta)
{
    for (int64_t y = 1; y <= ny; ++y)
        for (int64_t x = 1; x <= nx; ++x) {
            double cur = atan2(XS(x, y, 2), XS(x, y, 1));
            XS(x, y, 1) = cos(cur + theta);
            XS(x, y, 2) = sin(cur + theta);
        }
}

/* metropolis_by_field_sub, :198-216 (note the reference's acceptance test
 * `randoms > 1 - exp(delta_energy)` is restated literally) */
ORC_API void orc_xy_metropolis_by_field(int64_t nx, int64_t ny, double *sp, const double *randoms,
                                        const double *candidates, double hx, double hy)
{
    const double pi = 4 * atan(1.0);
    for (int64_t y = 1; y <= ny; ++y)
        for (int64_t x = 1; x <= nx; ++x) {
            double c1 = cos(2 * pi * XR(candidates, x, y));
            double c2 = sin(2 * pi * XR(candidates, x, y));
            double de = -(hx * (c1 - XS(x, y, 1)) + hy * (c2 - XS(x, y, 2)));
            if (XR(randoms, x, y) > 1 - exp(de)) continue;
            XS(x, y, 1) = c1;
            XS(x, y, 2) = c2;
        }
}
#undef XS
#undef XR

/* ==========================================================================
 * XY 2D helical  (src/xy2d_gpu_m.f90, module xy2d_gpu_m): the XY analogue of ising2d_gpu_m.
 * storage: spins(1-nx : nall+nx, 1:2) real64 (cos, sin), column-major:
 *   S(i, k) = sp[(i - 1 + nx) + len*(k-1)],  len = nall + 2 nx;  colour = parity of the linear index i.
 * randoms(1:nall), candidates(1:nall): r[i-1].
 * ========================================================================== */
#define HS(i, k) sp[((i)-1 + nx) + len * ((k)-1)]

/* update_norishiro_sub, :114-125 */
ORC_API void orc_xyh_norishiro(int64_t nx, int64_t ny, double *sp)
{
    const int64_t nall = nx * ny, len = nall + 2 * nx;
    for (int k = 1; k <= 2; ++k)
        for (int64_t idx = 1; idx <= nx; ++idx) {
            HS(nall + idx, k) = HS(idx, k);
            HS(idx - nx, k) = HS(nall - nx + idx, k);
        }
}

/* set_allup_spin_sub, :79-88 (halo cells included) */
ORC_API void orc_xyh_set_allup(int64_t nx, int64_t ny, double *sp)
{
    const int64_t len = nx * ny + 2 * nx;
    for (int64_t j = 0; j < len; ++j) { sp[j] = 1.0; sp[len + j] = 0.0; }
}

/* set_random_spin_xy2d_gpu, :90-105 */
ORC_API void orc_xyh_set_random(int64_t nx, int64_t ny, double *sp, const double *randoms)
{
    const int64_t nall = nx * ny, len = nall + 2 * nx;
    const double pi = 4 * atan(1.0);
    for (int64_t idx = 1; idx <= nall; ++idx) {
        HS(idx, 1) = cos(2 * pi * randoms[idx - 1]);
        HS(idx, 2) = sin(2 * pi * randoms[idx - 1]);
    }
    orc_xyh_norishiro(nx, ny, sp);
}

/* update_sub :157-174 with calc_delta_energy :243-252; offset 1 = odd linear indices, 2 = even */
static void xyh_pass(int64_t nx, int64_t ny, double *sp, double beta, const double *randoms,
                     const double *candidates, int offset)
{
    const int64_t nall = nx * ny, len = nall + 2 * nx;
    const double pi = 4 * atan(1.0);
#pragma omp parallel for schedule(static)
    for (int64_t idx = offset; idx <= nall; idx += 2) {
        double c1 = cos(2 * pi * candidates[idx - 1]);
        double c2 = sin(2 * pi * candidates[idx - 1]);
        double d1 = c1 - HS(idx, 1), d2 = c2 - HS(idx, 2);
        double n1 = HS(idx - 1, 1) + HS(idx + 1, 1) + HS(idx + nx, 1) + HS(idx - nx, 1);
        double n2 = HS(idx - 1, 2) + HS(idx + 1, 2) + HS(idx + nx, 2) + HS(idx - nx, 2);
        double de = -(d1 * n1 + d2 * n2);
        if (randoms[idx - 1] > exp(-beta * de)) continue;
        HS(idx, 1) = c1;
        HS(idx, 2) = c2;
    }
}

/* update_xy2d_gpu, :138-156 */
ORC_API void orc_xyh_update(int64_t nx, int64_t ny, double *sp, double beta, const double *randoms,
                            const double *candidates)
{
    xyh_pass(nx, ny, sp, beta, randoms, candidates, 1);
    orc_xyh_norishiro(nx, ny, sp);
    xyh_pass(nx, ny, sp, beta, randoms, candidates, 2);
    orc_xyh_norishiro(nx, ny, sp);
}

/* over_relaxation_sub, :198-213 (no renormalisation in this module) */
static void xyh_or_pass(int64_t nx, int64_t ny, double *sp, int offset)
{
    const int64_t nall = nx * ny, len = nall + 2 * nx;
#pragma omp parallel for schedule(static)
    for (int64_t idx = offset; idx <= nall; idx += 2) {
        double h1 = HS(idx - 1, 1) + HS(idx + 1, 1) + HS(idx + nx, 1) + HS(idx - nx, 1);
        double h2 = HS(idx - 1, 2) + HS(idx + 1, 2) + HS(idx + nx, 2) + HS(idx - nx, 2);
        double inv = 1 / hypot(h1, h2);
        h1 = h1 * inv;
        h2 = h2 * inv;
        double dot2 = 2 * (h1 * HS(idx, 1) + h2 * HS(idx, 2));
        double s1 = dot2 * h1 - HS(idx, 1), s2 = dot2 * h2 - HS(idx, 2);
        HS(idx, 1) = s1;
        HS(idx, 2) = s2;
    }
}

/* update_over_relaxation_xy2d_gpu, :176-196 */
ORC_API void orc_xyh_over_relaxation(int64_t nx, int64_t ny, double *sp, int32_t n_steps)
{
    for (int i = 0; i < n_steps; ++i) {
        xyh_or_pass(nx, ny, sp, 1);
        orc_xyh_norishiro(nx, ny, sp);
        xyh_or_pass(nx, ny, sp, 2);
        orc_xyh_norishiro(nx, ny, sp);
    }
}

/* calc_energy_sum :259-276, calc_magne_sum :278-291 */
ORC_API double orc_xyh_energy(int64_t nx, int64_t ny, const double *sp)
{
    const int64_t nall = nx * ny, len = nall + 2 * nx;
    double res = 0.0;
    for (int k = 1; k <= 2; ++k)
        for (int64_t i = 1; i <= nall; ++i) res -= HS(i, k) * (HS(i + 1, k) + HS(i + nx, k));
    return res;
}
ORC_API double orc_xyh_magne(int64_t nx, int64_t ny, const double *sp)
{
    const int64_t nall = nx * ny, len = nall + 2 * nx;
    double res = 0.0;
    for (int64_t i = 1; i <= nall; ++i) res += HS(i, 1);
    return res;
}
#undef HS
