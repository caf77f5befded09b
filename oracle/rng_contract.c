/* ==========================================================================
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT PATH.
 *
 * The RNG CONTRACT of the B200 build, restated on the CPU.
 *
 * The reference fills a device array with cuRAND XORWOW uniforms once per
 * sweep and the kernels index it by site (src/ising3d_gpu_m.f90:179,203;
 * src/ising2d_gpu_m.f90:138,159; src/clock_gpu_m.f90:188-189,211-212;
 * src/xy2d_periodic_gpu_m.f90:355-356,382-384).  The B200 build never
 * materialises that array: each uniform is a pure function
 *      u = f(seed, draw, site[, stream])
 * evaluated in registers with Philox4x32-10.  The functions below evaluate
 * the SAME function on the CPU and write the array the reference would have
 * read, in the reference's own index order, so that oracle.c (which consumes
 * arrays exactly like the reference) and the CUDA kernels see one stream.
 *
 * Written independently of the device code (csrc/rng.cuh): agreement of the
 * two is what the GPU parity tests establish.
 * ========================================================================== */
#include <math.h>
#include <stdint.h>
#include "philox.h"

#define ORC_API __attribute__((visibility("default")))

/* key[1] domain tags */
#define TAG_ISING 0x49534E47u /* "ISNG": accept uniforms of Ising 2D/3D     */
#define TAG_INIT  0x494E4954u /* "INIT": set_random_spin of every model     */
#define TAG_CLOCK 0x434C4F4Bu /* "CLOK": clock (helical) accept + proposal  */
#define TAG_TORUS 0x544F5253u /* "TORS": periodic clock (tableall)          */
#define TAG_XY    0x58593244u /* "XY2D": XY accept + candidate              */

/* counter layout (all models):
 *   c0 = block index, low 32      c1 = block index, high 32
 *   c2 = draw index, low 32       c3 = draw[32..47] | colour << 16 | sub << 24
 * "draw" counts generate calls (one per sweep / per set_random_spin);
 * skip_curand(n) advances it. */
static inline void mk_ctr(uint32_t c[4], uint64_t blk, uint64_t draw, uint32_t colour, uint32_t sub)
{
    c[0] = (uint32_t)blk;
    c[1] = (uint32_t)(blk >> 32);
    c[2] = (uint32_t)draw;
    c[3] = (uint32_t)((draw >> 32) & 0xFFFFu) | (colour << 16) | (sub << 24);
}

/* --------------------------------------------------------------------------
 * Ring models (Ising 2D, Ising 3D): the helical lattice is a ring of N sites
 * with 0-based linear index i = idx - 1, colour = i & 1, colour-site index
 * k = i >> 1 (Nc = N/2 per colour).  The build folds each colour ring into
 * 16 byte-lanes of L = ceil(Nc/16) positions: lane = k / L, p = k % L, so one
 * 128-bit vector (index p) holds 16 sites that are L apart.
 *
 * Accept uniform, 32-bit resolution, evaluated lazily on the GPU in two stages:
 *   stage 1  R  = philox(ctr(p, draw, colour, 0), (seed, TAG_ISING)); 16 bytes
 *            lane -> byte position m = BYTEPOS[lane];  b7 = byte_m(R) & 0x7F
 *   stage 2  R2 = philox(ctr(p, draw, colour, 1 + (m >> 2)), same key)
 *            low25 = R2[m & 3] & 0x1FFFFFF
 *   U = b7 << 25 | low25 ;  u = (U + 1) * 2^-32  in (0, 1]
 * (the GPU only evaluates stage 2 when b7 equals the top 7 bits of the
 * acceptance threshold; the value is the same either way).
 * -------------------------------------------------------------------------- */
static const int BYTEPOS[16] = {0, 2, 4, 6, 1, 3, 5, 7, 8, 10, 12, 14, 9, 11, 13, 15};
/* inverse view: byte position m of the Philox output serves lane
 *   {0,4,1,5, 2,6,3,7, 8,12,9,13, 10,14,11,15}[m]                         */

ORC_API int64_t orc_ring_fold_len(int64_t n_sites)
{
    int64_t nc = n_sites / 2;
    return (nc + 15) / 16;
}

static inline uint32_t ring_U_rep(uint32_t seed, uint32_t tag, uint64_t draw, int64_t L, int64_t i, uint32_t rep);
static inline uint32_t ring_U(uint32_t seed, uint32_t tag, uint64_t draw, int64_t L, int64_t i)
{
    return ring_U_rep(seed, tag, draw, L, i, 0);
}
/* sample `rep` of a batch (b200mc_ising*_create_multi): the sample index is the high word of the block counter */
static inline uint32_t ring_U_rep(uint32_t seed, uint32_t tag, uint64_t draw, int64_t L, int64_t i, uint32_t rep)
{
    uint32_t colour = (uint32_t)(i & 1);
    int64_t k = i >> 1;
    int lane = (int)(k / L);
    uint64_t p = (uint64_t)(k % L) | ((uint64_t)rep << 32);
    int m = BYTEPOS[lane];
    uint32_t key[2] = {seed, tag}, c[4], r[4], r2[4];
    mk_ctr(c, p, draw, colour, 0);
    orc_philox4x32_10(c, key, r);
    uint32_t b7 = (r[m >> 2] >> (8 * (m & 3))) & 0x7Fu;
    mk_ctr(c, p, draw, colour, 1u + (uint32_t)(m >> 2));
    orc_philox4x32_10(c, key, r2);
    uint32_t low25 = r2[m & 3] & 0x1FFFFFFu;
    return (b7 << 25) | low25;
}

/* out[i] = u(site i), i = 0..N-1 (== randoms(idx), idx = i+1, of the reference) */
ORC_API void orc_ising_uniforms(uint32_t seed, uint64_t draw, int64_t n_sites, double *out)
{
    const int64_t L = orc_ring_fold_len(n_sites);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_sites; ++i)
        out[i] = ((double)ring_U(seed, TAG_ISING, draw, L, i) + 1.0) * 0x1p-32;
}

ORC_API void orc_ising_uniforms_rep(uint32_t seed, uint64_t draw, uint32_t rep, int64_t n_sites, double *out)
{
    const int64_t L = orc_ring_fold_len(n_sites);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_sites; ++i)
        out[i] = ((double)ring_U_rep(seed, TAG_ISING, draw, L, i, rep) + 1.0) * 0x1p-32;
}

ORC_API void orc_ring_init_uniforms_rep(uint32_t seed, uint64_t draw, uint32_t rep, int64_t n_sites, double *out)
{
    const int64_t L = orc_ring_fold_len(n_sites);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_sites; ++i) {
        uint32_t colour = (uint32_t)(i & 1);
        int64_t k = i >> 1;
        int lane = (int)(k / L);
        uint64_t p = (uint64_t)(k % L) | ((uint64_t)rep << 32);
        uint32_t key[2] = {seed, TAG_INIT}, c[4], r[4];
        mk_ctr(c, p, draw, colour, (uint32_t)(lane >> 2));
        orc_philox4x32_10(c, key, r);
        out[i] = ((double)r[lane & 3] + 1.0) * 0x1p-32;
    }
}

/* set_random_spin uniforms (one 32-bit uniform per site, ring models):
 *   R = philox(ctr(p, draw, colour, lane >> 2), (seed, TAG_INIT)); U = R[lane & 3] */
ORC_API void orc_ring_init_uniforms(uint32_t seed, uint64_t draw, int64_t n_sites, double *out)
{
    const int64_t L = orc_ring_fold_len(n_sites);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_sites; ++i) {
        uint32_t colour = (uint32_t)(i & 1);
        int64_t k = i >> 1;
        int lane = (int)(k / L);
        uint64_t p = (uint64_t)(k % L);
        uint32_t key[2] = {seed, TAG_INIT}, c[4], r[4];
        mk_ctr(c, p, draw, colour, (uint32_t)(lane >> 2));
        orc_philox4x32_10(c, key, r);
        out[i] = ((double)r[lane & 3] + 1.0) * 0x1p-32;
    }
}

/* --------------------------------------------------------------------------
 * Periodic Ising (torus, csrc/ising_periodic.cu): colour = (x + y + z) & 1, the sites of a colour stored
 * row-major with xi = x >> 1 running fastest, 16 consecutive xi per 128-bit vector:
 *   k = (z ny + y) (nx / 2) + xi,  vector v = k / 16,  lane = k % 16.
 * The uniforms are the ring models' with (position, lane) = (v, lane): the same two-stage accept uniform
 * (TAG_ISING) and the same one-word set_random_spin uniform (TAG_INIT).  out[i], i = x + nx (y + ny z).
 * -------------------------------------------------------------------------- */
static inline void torus_site(int64_t i, int64_t nx, int64_t ny, uint32_t *colour, uint64_t *v, int *lane)
{
    const int64_t x = i % nx, y = (i / nx) % ny, z = i / (nx * ny);
    const int64_t k = (z * ny + y) * (nx / 2) + (x >> 1);
    *colour = (uint32_t)((x + y + z) & 1);
    *v = (uint64_t)(k >> 4);
    *lane = (int)(k & 15);
}

ORC_API void orc_isingp_uniforms(uint32_t seed, uint64_t draw, int64_t nx, int64_t ny, int64_t nz, double *out)
{
    const int64_t n = nx * ny * nz;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        uint32_t colour; uint64_t v; int lane;
        torus_site(i, nx, ny, &colour, &v, &lane);
        const int m = BYTEPOS[lane];
        uint32_t key[2] = {seed, TAG_ISING}, c[4], r[4], r2[4];
        mk_ctr(c, v, draw, colour, 0);
        orc_philox4x32_10(c, key, r);
        const uint32_t b7 = (r[m >> 2] >> (8 * (m & 3))) & 0x7Fu;
        mk_ctr(c, v, draw, colour, 1u + (uint32_t)(m >> 2));
        orc_philox4x32_10(c, key, r2);
        const uint32_t U = (b7 << 25) | (r2[m & 3] & 0x1FFFFFFu);
        out[i] = ((double)U + 1.0) * 0x1p-32;
    }
}

ORC_API void orc_isingp_init_uniforms(uint32_t seed, uint64_t draw, int64_t nx, int64_t ny, int64_t nz, double *out)
{
    const int64_t n = nx * ny * nz;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        uint32_t colour; uint64_t v; int lane;
        torus_site(i, nx, ny, &colour, &v, &lane);
        uint32_t key[2] = {seed, TAG_INIT}, c[4], r[4];
        mk_ctr(c, v, draw, colour, (uint32_t)(lane >> 2));
        orc_philox4x32_10(c, key, r);
        out[i] = ((double)r[lane & 3] + 1.0) * 0x1p-32;
    }
}

/* --------------------------------------------------------------------------
 * Clock models, contract v3 (cuda_fortran_mc_simulation_spin_b200/csrc/clock_word.cuh): the two 32-bit
 * uniforms (accept, proposal) of the 16 sites of a vector = words w = 0..3 of sites e = 0..3.
 * X[0..11] = the words of the Philox blocks with sub-counters 0, 1, 2; Y[0..11] = sub-counters
 * 4, 5, 6 (the GPU only evaluates Y when the first look at an accept uniform -- 15 bits --
 * does not decide; the value is the same):
 *   half(x, hs) = hs ? x >> 16 : x & 0xFFFF
 *   accept    a16 = half(X[3w + (e & 1)], e >> 1),  a16' = half(Y[3w + (e & 1)], e >> 1)
 *             U_a = (a16 & 0x7FFF) << 17 | (a16 >> 15) << 16 | a16'
 *   proposal  W_0 = X[3w + 2],  W_e = low32(W_(e-1) pm):  the proposals of the four sites are
 *             the leading base-pm digits d_e = floor(W_e pm / 2^32) of ONE uniform (pm = the
 *             number of proposal cells: q - 1 for new = c + ceiling(u (q-1)),
 *             src/clock/clock_tableall_gpu_m.f90:142; q for floor(u q), src/clock_gpu_m.f90:211).
 *             The array entry is U_p = W_e, nudged to W_e - 1 in the ~pm 2^-32 of all cases
 *             where the "+ 1" of u = (U + 1) 2^-32 would carry u pm across the cell boundary,
 *             so that the reference's own rounding of the real64 u reproduces d_e.
 * -------------------------------------------------------------------------- */
static inline uint32_t clk_half(uint32_t w, int hs) { return hs ? w >> 16 : w & 0xFFFFu; }
static inline uint32_t clk_cell(uint32_t U, uint32_t pm, int periodic)
{
    if (periodic) return (uint32_t)((((uint64_t)U + 1u) * pm + 0xFFFFFFFFull) >> 32) - 1u;   /* ceiling(u pm) - 1 */
    uint32_t k = (uint32_t)((((uint64_t)U + 1u) * pm) >> 32);                                 /* min(floor(u pm), pm - 1) */
    return k < pm - 1 ? k : pm - 1;
}
static void clk_vector_uniforms(uint64_t blk, uint64_t draw, uint32_t colour, const uint32_t key[2], uint32_t pm, int periodic,
                                uint32_t Ua[16], uint32_t Up[16])
{
    uint32_t X[12], Y[12], c[4];
    for (uint32_t i = 0; i < 3; ++i) {
        mk_ctr(c, blk, draw, colour, i);
        orc_philox4x32_10(c, key, X + 4 * i);
        mk_ctr(c, blk, draw, colour, 4u + i);
        orc_philox4x32_10(c, key, Y + 4 * i);
    }
    for (int w = 0; w < 4; ++w) {
        uint32_t W = X[3 * w + 2];
        for (int e = 0; e < 4; ++e) {
            const int hs = e >> 1;
            const uint32_t a16 = clk_half(X[3 * w + (e & 1)], hs), a16b = clk_half(Y[3 * w + (e & 1)], hs);
            Ua[4 * w + e] = ((a16 & 0x7FFFu) << 17) | ((a16 >> 15) << 16) | a16b;
            const uint64_t prod = (uint64_t)W * pm;
            const uint32_t d = (uint32_t)(prod >> 32);
            uint32_t U = W;
            if (clk_cell(U, pm, periodic) != d) U = W - 1u;
            Up[4 * w + e] = U;
            W = (uint32_t)prod;
        }
    }
}

/* --------------------------------------------------------------------------
 * Clock, helical ring (clock_gpu_m / clock_gpu_multi_m): same fold as Ising.  Vector p,
 * colour c, replica j, byte lane `lane` = site 4w + e of the vector:
 *   blocks philox(ctr(p, draw, c, sub), (seed, TAG_CLOCK + j)), pm = q
 *   accept U_a -> randoms(idx), proposal U_p -> next_states(idx)   (clk_vector_uniforms)
 * -------------------------------------------------------------------------- */
ORC_API void orc_clock_uniforms(uint32_t seed, uint64_t draw, int32_t replica, int64_t n_sites, int32_t q,
                                double *randoms, double *next_states)
{
    const int64_t L = orc_ring_fold_len(n_sites), nc = n_sites / 2;
    const uint32_t key[2] = {seed, TAG_CLOCK + (uint32_t)replica};
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < L; ++p)
        for (uint32_t colour = 0; colour < 2; ++colour) {
            uint32_t Ua[16], Up[16];
            clk_vector_uniforms((uint64_t)p, draw, colour, key, (uint32_t)q, 0, Ua, Up);
            for (int lane = 0; lane < 16; ++lane) {
                const int64_t k = (int64_t)lane * L + p;
                if (k >= nc) continue;
                const int64_t i = 2 * k + colour;
                randoms[i] = ((double)Ua[lane] + 1.0) * 0x1p-32;
                next_states[i] = ((double)Up[lane] + 1.0) * 0x1p-32;
            }
        }
}

/* block-wise evaluation of orc_ising_uniforms (same values, 5 Philox calls per
 * 16 sites instead of 2 per site): used where generation speed matters
 * (bench.py's CPU baseline, which like the reference draws a fresh array of
 * nall uniforms every sweep, src/ising3d_gpu_m.f90:179). */
ORC_API void orc_ising_uniforms_fast(uint32_t seed, uint64_t draw, int64_t n_sites, double *out)
{
    const int64_t L = orc_ring_fold_len(n_sites);
    const int64_t nc = n_sites / 2;
    const uint32_t key[2] = {seed, TAG_ISING};
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < L; ++p) {
        for (uint32_t colour = 0; colour < 2; ++colour) {
            uint32_t c[4], r[4], r2[4][4];
            mk_ctr(c, (uint64_t)p, draw, colour, 0);
            orc_philox4x32_10(c, key, r);
            for (uint32_t sub = 0; sub < 4; ++sub) {
                mk_ctr(c, (uint64_t)p, draw, colour, 1u + sub);
                orc_philox4x32_10(c, key, r2[sub]);
            }
            for (int lane = 0; lane < 16; ++lane) {
                const int64_t k = (int64_t)lane * L + p;
                if (k >= nc) continue;
                const int m = BYTEPOS[lane];
                const uint32_t b7 = (r[m >> 2] >> (8 * (m & 3))) & 0x7Fu;
                const uint32_t low25 = r2[m >> 2][m & 3] & 0x1FFFFFFu;
                out[2 * k + colour] = ((double)((b7 << 25) | low25) + 1.0) * 0x1p-32;
            }
        }
    }
}

/* --------------------------------------------------------------------------
 * XY periodic (xy2d_periodic_gpu_m): true torus, colour = (x0 + y0) & 1 with
 * 0-based x0, y0 (== the reference's (x + y) parity).  Colour-compact index
 * xi = x0 >> 1; a "group" is 4 consecutive xi of one row: blk = y0 * gpr + (xi >> 2),
 * gpr = ceil((nx/2) / 4).  Contract v2: 23 + 23 bits per site and sweep (both uniforms are used in
 * fp32; 23 bits are built into a float without an integer conversion: as_float(0x3F800000 | U) - 1 + 2^-23,
 * exact), three Philox blocks per 8 sites:
 *   R = philox(ctr(blk, draw, colour, 0), (seed, TAG_XY))                    one block per group and row
 *   C = philox(ctr(blk of the even row of the pair (y0 & ~1), draw, colour, 1), same key)
 *   candidate U_c = R[xi & 3] & 0x7FFFFF                                          -> candidates(x, y)
 *   accept    U_r = (R[xi & 3] >> 24 & 0x7F) << 16 | half(C[xi & 3], y0 & 1)      -> randoms(x, y)
 * u = (U + 1) 2^-23 in (0, 1]  (half(w, 0) = w & 0xFFFF, half(w, 1) = w >> 16).
 * set_random_spin: R = philox(ctr(blk, draw, colour, 0), (seed, TAG_INIT)), U = R[xi & 3].
 * Arrays are written in the reference's order randoms(nx, ny): out[x0 + nx * y0].
 * -------------------------------------------------------------------------- */
ORC_API void orc_xy_uniforms(uint32_t seed, uint64_t draw, int64_t nx, int64_t ny, double *randoms,
                             double *candidates)
{
    const int64_t nxh = nx / 2, gpr = (nxh + 3) / 4;
    const uint32_t key[2] = {seed, TAG_XY};
#pragma omp parallel for schedule(static)
    for (int64_t y0 = 0; y0 < ny; ++y0)
        for (int64_t x0 = 0; x0 < nx; ++x0) {
            const uint32_t colour = (uint32_t)((x0 + y0) & 1);
            const int64_t xi = x0 >> 1;
            uint32_t c[4], r[4], cp[4];
            mk_ctr(c, (uint64_t)(y0 * gpr + (xi >> 2)), draw, colour, 0u);
            orc_philox4x32_10(c, key, r);
            mk_ctr(c, (uint64_t)((y0 & ~(int64_t)1) * gpr + (xi >> 2)), draw, colour, 1u);
            orc_philox4x32_10(c, key, cp);
            const uint32_t W = r[xi & 3], ch = (y0 & 1) ? cp[xi & 3] >> 16 : cp[xi & 3] & 0xFFFFu;
            randoms[x0 + nx * y0] = ((double)((((W >> 24) & 0x7Fu) << 16) | ch) + 1.0) * 0x1p-23;
            candidates[x0 + nx * y0] = ((double)(W & 0x7FFFFFu) + 1.0) * 0x1p-23;
        }
}

ORC_API void orc_xy_init_uniforms(uint32_t seed, uint64_t draw, int64_t nx, int64_t ny, double *out)
{
    const int64_t nxh = nx / 2, gpr = (nxh + 3) / 4;
    const uint32_t key[2] = {seed, TAG_INIT};
#pragma omp parallel for schedule(static)
    for (int64_t y0 = 0; y0 < ny; ++y0)
        for (int64_t x0 = 0; x0 < nx; ++x0) {
            const uint32_t colour = (uint32_t)((x0 + y0) & 1);
            const int64_t xi = x0 >> 1;
            uint32_t c[4], r[4];
            mk_ctr(c, (uint64_t)(y0 * gpr + (xi >> 2)), draw, colour, 0);
            orc_philox4x32_10(c, key, r);
            out[x0 + nx * y0] = ((double)r[xi & 3] + 1.0) * 0x1p-32;
        }
}

/* --------------------------------------------------------------------------
 * Periodic clock (clock_tableall_gpu_m / clock_dual_lattice_tableall_gpu_m): true torus,
 * colour = (x0 + y0) & 1, colour-compact index xi = x0 >> 1; rows are cut into vectors of 16
 * compact sites: v = xi >> 4, j = xi & 15, nvr = ceil((nx/2) / 16), block = y0 * nvr + v.
 * Two 32-bit uniforms per site and sweep (contract v3, clk_vector_uniforms above), pm = q - 1:
 *   blocks philox(ctr(replica << 32 | block, draw, colour, sub), (seed, TAG_TORUS))
 *   accept   U_a -> rnds(2, x, y)      proposal U_p -> rnds(1, x, y)
 * (the sample index of a batch is the high word of the block counter, as for the Ising batches)
 * u = (U + 1) 2^-32.  Written in the reference's order rnds(2, nx, ny)
 * (src/clock/clock_tableall_gpu_m.f90:95): out[(j-1) + 2*(x0 + nx*y0)].
 * -------------------------------------------------------------------------- */
ORC_API void orc_torus_uniforms(uint32_t seed, uint64_t draw, int32_t replica, int64_t nx, int64_t ny, int32_t q,
                                double *rnds)
{
    const int64_t nxh = nx / 2, nvr = (nxh + 15) / 16;
    const uint32_t key[2] = {seed, TAG_TORUS};
#pragma omp parallel for schedule(static)
    for (int64_t y0 = 0; y0 < ny; ++y0)
        for (int64_t v = 0; v < nvr; ++v)
            for (uint32_t colour = 0; colour < 2; ++colour) {
                uint32_t Ua[16], Up[16];
                const uint64_t blk = (uint64_t)(y0 * nvr + v) | ((uint64_t)(uint32_t)replica << 32);
                clk_vector_uniforms(blk, draw, colour, key, (uint32_t)(q - 1), 1, Ua, Up);
                for (int j = 0; j < 16; ++j) {
                    const int64_t xi = 16 * v + j;
                    if (xi >= nxh) break;
                    const int64_t x0 = 2 * xi + (int64_t)((y0 + colour) & 1);
                    rnds[0 + 2 * (x0 + nx * y0)] = ((double)Up[j] + 1.0) * 0x1p-32;
                    rnds[1 + 2 * (x0 + nx * y0)] = ((double)Ua[j] + 1.0) * 0x1p-32;
                }
            }
}

/* --------------------------------------------------------------------------
 * XY helical (xy2d_gpu_m): ring of N sites, colour = i & 1 (i = idx - 1), colour-site index k = i >> 1;
 * a group is 4 consecutive k: blk = k >> 2.  Two 32-bit uniforms per site and sweep:
 *   R = philox(ctr(blk, draw, colour, (k & 3) >> 1), (seed, TAG_XYH))
 *   accept U_r = R[2 * (k & 1)] -> randoms(idx) ;  candidate U_c = R[2 * (k & 1) + 1] -> candidates(idx)
 * set_random_spin: R = philox(ctr(blk, draw, colour, 0), (seed, TAG_INIT)), U = R[k & 3].
 * -------------------------------------------------------------------------- */
#define TAG_XYH 0x5859484Cu /* "XYHL" */
ORC_API void orc_xyh_uniforms(uint32_t seed, uint64_t draw, int64_t n_sites, double *randoms, double *candidates)
{
    const uint32_t key[2] = {seed, TAG_XYH};
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_sites; ++i) {
        const uint32_t colour = (uint32_t)(i & 1);
        const int64_t k = i >> 1;
        uint32_t c[4], r[4];
        mk_ctr(c, (uint64_t)(k >> 2), draw, colour, (uint32_t)((k & 3) >> 1));
        orc_philox4x32_10(c, key, r);
        randoms[i] = ((double)r[2 * (k & 1)] + 1.0) * 0x1p-32;
        candidates[i] = ((double)r[2 * (k & 1) + 1] + 1.0) * 0x1p-32;
    }
}
ORC_API void orc_xyh_init_uniforms(uint32_t seed, uint64_t draw, int64_t n_sites, double *out)
{
    const uint32_t key[2] = {seed, TAG_INIT};
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_sites; ++i) {
        const uint32_t colour = (uint32_t)(i & 1);
        const int64_t k = i >> 1;
        uint32_t c[4], r[4];
        mk_ctr(c, (uint64_t)(k >> 2), draw, colour, 0);
        orc_philox4x32_10(c, key, r);
        out[i] = ((double)r[k & 3] + 1.0) * 0x1p-32;
    }
}

/* --------------------------------------------------------------------------
 * Bit-packed Ising 2D / 3D (cuda_fortran_mc_simulation_spin_b200/csrc/ising_bits.cu): each colour
 * ring of Nc = N / 2 sites folded into 128 bit-lanes of L = Nc / 128 positions (128 | Nc): site
 * k -> lane k / L, position p = k % L = bit (lane & 31) of word (lane >> 5) of vector p.
 * Accept uniform: the 8 leading bits from bit planes (plane j of the 128 uniforms of a vector = one
 * Philox block), the low 24 bits from a per-site Philox word:
 *   R_j = philox(ctr(p, draw, colour, sub = j), (seed, TAG_ISNB)), j = 0..7
 *   S   = philox(ctr(p, draw, colour, sub = 32 + (lane >> 2)), same key)
 *   U = sum_j ((R_j[lane >> 5] >> (lane & 31)) & 1) << (31 - j)  |  S[lane & 3] & 0xFFFFFF;   u = (U + 1) 2^-32
 * (the GPU compares plane by plane with the threshold's bits and evaluates S only for a site whose 8
 * leading bits equal its threshold's; the value is the same).
 * set_random_spin: R = philox(ctr(p, draw, colour, 0), (seed, TAG_INIB)),
 *   U = ((R[lane >> 5] >> (lane & 31)) & 1) << 31   (the reference only tests u < 0.5,
 *   src/ising3d_gpu_m.f90:99).
 * out[i] = u(site i), i = 0..N-1.
 * -------------------------------------------------------------------------- */
#define TAG_ISNB 0x49534E42u
#define TAG_INIB 0x494E4942u
ORC_API int orc_isingbits_uniforms(uint32_t seed, uint64_t draw, int64_t n_sites, int init, double *out)
{
    const int64_t nc = n_sites / 2;
    if (nc % 128) return 1;
    const int64_t L = nc / 128;
    const uint32_t key[2] = {seed, init ? TAG_INIB : TAG_ISNB};
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < L; ++p)
        for (uint32_t colour = 0; colour < 2; ++colour) {
            uint32_t U[128] = {0}, c[4], r[4];
            const int planes = init ? 1 : 8;
            for (int j = 0; j < planes; ++j) {
                mk_ctr(c, (uint64_t)p, draw, colour, (uint32_t)j);
                orc_philox4x32_10(c, key, r);
                for (int lane = 0; lane < 128; ++lane) U[lane] |= ((r[lane >> 5] >> (lane & 31)) & 1u) << (31 - j);
            }
            if (!init)
                for (int q4 = 0; q4 < 32; ++q4) {
                    mk_ctr(c, (uint64_t)p, draw, colour, 32u + (uint32_t)q4);
                    orc_philox4x32_10(c, key, r);
                    for (int t = 0; t < 4; ++t) U[4 * q4 + t] |= r[t] & 0xFFFFFFu;
                }
            for (int lane = 0; lane < 128; ++lane) {
                const int64_t k = (int64_t)lane * L + p;
                out[2 * k + colour] = ((double)U[lane] + 1.0) * 0x1p-32;
            }
        }
    return 0;
}
