// XY 2D periodic (xy2d_periodic_gpu_m): kernels, host-side handle and C ABI.
// Reference: type(xy2d_gpu) src/xy2d_periodic_gpu_m.f90:14-59.
//
// Layout: the reference stores (cos, sin) as real64 SoA with a one-cell halo frame,
// spins(0:nx+1, 0:ny+1, 1:2) (16 B/site + two real64 uniform arrays).  Here a spin is ONE
// fp32 angle in TURNS (theta / 2 pi), colours split into compact arrays [ny][nx/2]
// (4 B/site); the periodic wrap is index arithmetic, so there is no halo to refresh on one GPU.
// A thread owns 4 consecutive colour-compact sites of one row (one aligned 128-bit load) and
// reads 3 aligned float4 + 1 scalar of the other colour.  cos/sin come from the SFU
// (MUFU.SIN/COS after reduction to [-1/2, 1/2) turns); uniforms from Philox in registers.
#include <math.h>
#include <stdlib.h>
#include <new>
#include "../../include/b200mc.h"
#include "common.cuh"
#include "ring.cuh"   // dist_*: the NCCL communicator of slab mode

namespace {

#define TWO_PI_F 6.283185307179586f
#define INV_TWO_PI_F 0.15915494309189535f

struct XYArgs {
    float* own;
    const float* oth;
    int nxh, ny, gpr;  // sites per row per colour, rows, float4 groups per row
    int pitch;         // floats per row of a colour array: nxh rounded up to 4; the pitch - nxh padding floats of a row mirror its
                       // first sites (xi = 0, 1, ..: the periodic continuation to the right), kept current by every kernel that
                       // writes xi < pitch - nxh.  With them the last, partial group of a row (nx/2 not a multiple of 4; the
                       // reference only needs nx even, src/xy2d_periodic_gpu_m.f90:377-380) reads its right-hand neighbour like
                       // every other group; its own padding lanes are computed but neither stored nor summed.
    int colour;
    float beta;
    uint64_t draw;
    uint32_t rk0[10];
    double* acc;   // MEASURE: acc[0] += E, acc[1] += sum cos, acc[2] += sum sin
    // rows as a slab (the step towards slabs along y): halo = 1 -> the other colour's rows -1 and ny are
    // halo rows kept current by the host side after every colour pass (no periodic wrap of the row index in the
    // kernels); yoff = global index of local row 0 (even), used for the RNG counter
    int halo, yoff;
};

__device__ __forceinline__ void sincos_turns(float t, float& s, float& c)
{
    // t - rint(t) in [-1/2, 1/2] with the 1.5 * 2^23 trick (|t| < 2^22): two FADD instead of FRND, which shares the
    // XU pipe with MUFU -- the pipe that bounds these kernels
#ifdef XY_FRND
    const float r = t - rintf(t);
#else
    const float r = t - __fsub_rn(__fadd_rn(t, 12582912.0f), 12582912.0f);
#endif
    __sincosf(turns_to_mufu_arg_centred(r), &s, &c);        // MUFU.SIN / MUFU.COS, |x| <= pi
}

template <uint32_t TAG>
__device__ __forceinline__ uint4 philox_tag(uint4 c, const uint32_t (&rk0)[10])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t lo0, hi0, lo1, hi1;
        mulwide(PHILOX_M0, c.x, lo0, hi0);
        mulwide(PHILOX_M1, c.z, lo1, hi1);
        uint4 n;
        n.x = hi1 ^ c.y ^ rk0[r];
        n.y = lo1;
        n.z = hi0 ^ c.w ^ (TAG + (uint32_t)r * PHILOX_W1);
        n.w = lo0;
        c = n;
    }
    return c;
}

// Column-strip traversal.  A thread owns the 4 colour-compact sites of group g and walks XY_ROWS consecutive
// rows.  The cos / sin of the other colour's row y (4 values) serve three updates: the same-row neighbours of
// row y, the "up" neighbours of row y - 1 and the "down" neighbours of row y + 1 -- so they are computed once
// and kept in a rolling three-row register window.  Per 4 sites and row that is 13 sincos (4 new neighbour
// values, the fifth same-row value, 4 own, 4 candidates) instead of 21: the kernel is bound by the SFU queue.
#ifndef XY_ROWS
#define XY_ROWS 32
#endif

struct XYRow { float c[4], s[4]; };

__device__ __forceinline__ void xy_load_row(const XYArgs& a, int y, int g, XYRow& r)
{
    const float4 raw = *reinterpret_cast<const float4*>(a.oth + (ptrdiff_t)y * a.pitch + 4 * g);   // y = -1 / ny: halo rows
    // (the same function as for the rows loaded inside the strip loop: a site's cos / sin must not depend on where the
    // row falls in a strip, or the trajectory would depend on the decomposition into strips and slabs)
    sincos_unit(raw.x, r.s[0], r.c[0]);
    sincos_unit(raw.y, r.s[1], r.c[1]);
    sincos_unit(raw.z, r.s[2], r.c[2]);
    sincos_unit(raw.w, r.s[3], r.c[3]);
}

// store the 4 new values of group xi0 of a row (row = start of the row in the colour array): the last group of a row
// may be partial, and the first group also refreshes the row's mirror padding
__device__ __forceinline__ void xy_store_group(float* row, int xi0, int nxh, int pitch, const float (&ov)[4])
{
    if (xi0 + 4 <= nxh) *reinterpret_cast<float4*>(row + xi0) = make_float4(ov[0], ov[1], ov[2], ov[3]);
    else {
#pragma unroll
        for (int j = 0; j < 3; ++j) if (xi0 + j < nxh) row[xi0 + j] = ov[j];
    }
    if (xi0 == 0 && pitch != nxh) {
#pragma unroll
        for (int j = 0; j < 3; ++j) if (nxh + j < pitch) row[nxh + j] = ov[j];
    }
}

// OVERRELAX = false: update_sub + calc_delta_energy, src/xy2d_periodic_gpu_m.f90:368-397
// OVERRELAX = true : over_relaxation_sub, :418-439: reflect the spin about the local field.  In angles:
//                    theta' = 2 phi - theta with phi = atan2(h_y, h_x) (the reference's renormalisation is the identity here)
// MEASURE (second colour pass of a sweep when the caller measures every MCS): every bond has exactly one end in
// the colour being updated, so E = -sum over these sites of s_new . h; sum cos / sum sin: the new spins plus the
// other colour's row `mid` (each of its sites belongs to exactly one thread-row).
// (sincos_unit, atan2_turns, frac_turns: common.cuh -- shared with the helical module)

#ifndef XY_MINB_M
#define XY_MINB_M 4   // Metropolis: 64 registers
#endif
#ifndef XY_MINB_O
#define XY_MINB_O 3   // over-relaxation and the variants with fused sums: 80 registers
#endif
static_assert(XY_ROWS % 2 == 0, "the strip loop is unrolled by two rows (the neighbour pattern alternates with the row parity)");

// One row of a strip.  P = (y + colour) & 1 at compile time: the x position of colour-compact site xi is 2 xi + P, its
// same-row neighbours are the other colour's xi - 1 + P and xi + P.
// RNG contract (v2; CPU restatement: oracle/rng_contract.c, orc_xy_uniforms).  A group = 4 colour-compact sites of one row,
// blk = y gpr + g.  Per site 23 + 23 bits, turned into fp32 WITHOUT an integer-to-float conversion (I2F shares the XU pipe
// with MUFU, the pipe that bounds the Metropolis pass): as_float(0x3F800000 | U) is 1 + U 2^-23 in [1, 2), one FADD later
// (U + 1) 2^-23 in (0, 1], exact:
//   R = philox(ctr(blk, draw, colour, 0), (seed, TAG_XY))               one block per group and row
//   C = philox(ctr(blk of the EVEN row of the pair (y & ~1), draw, colour, 1), same key)    one block per group and row PAIR
//   candidate U_c = R[j] & 0x7FFFFF;   accept U_r = (R[j] >> 24 & 0x7F) << 16 | half(C[j], y & 1);   u = (U + 1) 2^-23 in (0, 1]
// (the accept mantissa is byte-aligned: one PRMT gathers the two bytes of C and the top byte of R, one LOP3 sets the exponent)
// Three Philox blocks per 8 sites instead of four (round 1: two full 32-bit words per site, rounded to fp32).
__device__ __forceinline__ uint4 xy_pair_block(const XYArgs& a, uint64_t blk_even)
{
    return philox_tag<TAG_XY>(mk_ctr(blk_even, a.draw, (uint32_t)a.colour, 1u), a.rk0);
}
__device__ __forceinline__ void xy_group_uniforms(const XYArgs& a, uint64_t blk, const uint4& C, int odd, float (&r)[4], float (&ct)[4])
{
    const uint4 R = philox_tag<TAG_XY>(mk_ctr(blk, a.draw, (uint32_t)a.colour, 0u), a.rk0);
    const uint32_t W[4] = {R.x, R.y, R.z, R.w}, cw[4] = {C.x, C.y, C.z, C.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        // bytes (C half lo, C half hi, R byte 3, -) -> mantissa; bit 23 and the exponent are forced by the mask / or
        const uint32_t g = prmt(cw[j], W[j], odd ? 0x0732u : 0x0710u);
        ct[j] = __uint_as_float((W[j] & 0x007FFFFFu) | 0x3F800000u) + (0x1p-23f - 1.0f);      // candidate angle in turns, (0, 1]
        r[j] = __uint_as_float((g & 0x007FFFFFu) | 0x3F800000u) + (0x1p-23f - 1.0f);          // accept uniform, (0, 1]
    }
}

// RAGGED = false (nx/2 a multiple of 4, e.g. every benchmark shape): no partial group, no mirror padding -- the strip code
// of round 1, 64 registers without spills; RAGGED = true adds the per-lane validity tests and the mirror stores.
template <bool OVERRELAX, bool MEASURE, int P, bool RAGGED, int ODD>
__device__ __forceinline__ void xy_strip_row(const XYArgs& a, int y, int idx, const uint4& C, float nbl2e, const float4 r_up, const float r_edge, const float4 o,
                                             const XYRow& dn, const XYRow& mid, XYRow& up, float* prow, int xi0, float& es, float& mx, float& my)
{
    const int nvalid = RAGGED ? a.nxh - xi0 : 4;   // >= 4 except in the last group of a row whose nx/2 is not a multiple of 4
    sincos_unit(r_up.x, up.s[0], up.c[0]);
    sincos_unit(r_up.y, up.s[1], up.c[1]);
    sincos_unit(r_up.z, up.s[2], up.c[2]);
    sincos_unit(r_up.w, up.s[3], up.c[3]);
    float se, ce;
    sincos_unit(r_edge, se, ce);
    // ordered left to right: P = 0: e, m0, m1, m2, m3 ; P = 1: m0, m1, m2, m3, e
    // (same summation order as calc_delta_energy, src/xy2d_periodic_gpu_m.f90:395: x+1, x-1, y+1, y-1)
    float hx[4], hy[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float lc = P ? mid.c[j] : (j ? mid.c[j - 1] : ce), ls = P ? mid.s[j] : (j ? mid.s[j - 1] : se);
        const float rc = P ? (j < 3 ? mid.c[j + 1] : ce) : mid.c[j], rs = P ? (j < 3 ? mid.s[j + 1] : se) : mid.s[j];
        hx[j] = rc + lc + up.c[j] + dn.c[j];
        hy[j] = rs + ls + up.s[j] + dn.s[j];
    }
    float ov[4] = {o.x, o.y, o.z, o.w};
    if (OVERRELAX) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float t = 2.0f * atan2_turns(hy[j], hx[j]) - ov[j];                    // (-2, 1]
            ov[j] = frac_turns(t);                                                        // [0, 1]
            if (MEASURE && j < nvalid) {
                float sn, cn;
                sincos_unit(ov[j], sn, cn);
                es -= cn * hx[j] + sn * hy[j];
                mx += cn; my += sn;
            }
        }
    } else {
        float rr[4], cand[4];
        xy_group_uniforms(a, (uint64_t)idx, C, ODD, rr, cand);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float r = rr[j], ct = cand[j];
            float cs, cc, ss, sc;
            sincos_unit(ct, cs, cc);
            sincos_unit(ov[j], ss, sc);
            const float de = (cc - sc) * hx[j] + (cs - ss) * hy[j];   // -dE
            float w;                                                   // exp(-beta dE) = 2^(beta log2(e) (-dE))
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w) : "f"(de * nbl2e));
            const bool acc = !(r > w);                                 // accept iff r <= exp(-beta dE), :384
            if (acc) ov[j] = ct;
            if (MEASURE && j < nvalid) {
                const float cn = acc ? cc : sc, sn = acc ? cs : ss;
                es -= cn * hx[j] + sn * hy[j];
                mx += cn; my += sn;
            }
        }
    }
    if (MEASURE) {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (j < nvalid) { mx += mid.c[j]; my += mid.s[j]; }
    }
    if constexpr (RAGGED) xy_store_group(prow, xi0, a.nxh, a.pitch, ov);
    else *reinterpret_cast<float4*>(prow + xi0) = make_float4(ov[0], ov[1], ov[2], ov[3]);
}

template <bool OVERRELAX, bool MEASURE, int COLOUR, bool RAGGED>
__global__ void __launch_bounds__(256, (OVERRELAX || MEASURE) ? XY_MINB_O : XY_MINB_M)
xy_strip_kernel(const __grid_constant__ XYArgs a)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nblk = (a.ny + XY_ROWS - 1) / XY_ROWS;
    const bool active = tid < nblk * a.gpr;
    float es = 0.f, mx = 0.f, my = 0.f;
    if (active) {
        const int rb = tid / a.gpr, g = tid - rb * a.gpr;
        const int y0 = rb * XY_ROWS, y1 = min(y0 + XY_ROWS, a.ny);   // both even (ny is even)
        const int nxh = a.nxh, xi0 = 4 * g, pitch = RAGGED ? a.pitch : a.nxh;
        const float nbl2e = a.beta * 1.4426950408889634f;
        XYRow dn, mid, up;
        xy_load_row(a, (y0 == 0 && !a.halo) ? a.ny - 1 : y0 - 1, g, dn);
        xy_load_row(a, y0, g, mid);
        // the fifth same-row value: left of the group on rows with P = 0, right of it on rows with P = 1
        const int xe0 = xi0 == 0 ? nxh - 1 : xi0 - 1, xe1 = xi0 + 4 >= nxh ? 0 : xi0 + 4;   // (a partial last group never uses xe1: its right-hand neighbour is the mirror padding)
        // the raw values of a row (the other colour's row y + 1, the fifth same-row value, the own row) are
        // loaded one row ahead, unconditionally (clamped row index: straight-line code that the scheduler issues
        // at the top): the stores to `own` would otherwise pin every load behind them (no restrict on the two
        // colour arrays) and each row would pay a full DRAM round trip
        auto ld_up = [&](int y) { return __ldg(reinterpret_cast<const float4*>(a.oth + (ptrdiff_t)((y + 1 >= a.ny && !a.halo) ? y + 1 - a.ny : y + 1) * pitch + xi0)); };
        auto ld_own = [&](int y) { return *reinterpret_cast<const float4*>(a.own + (size_t)y * pitch + xi0); };
        float4 u0 = ld_up(y0), o0 = ld_own(y0);
        float e0 = __ldg(a.oth + (size_t)y0 * pitch + (COLOUR ? xe1 : xe0));
        if constexpr (OVERRELAX) {
            // over-relaxation (fewer instructions per row, 80 registers): TWO rows ahead -- the raw values of rows y + 2 and
            // y + 3 are requested before rows y and y + 1 are processed (463 -> 512 flips/ns; no effect on Metropolis,
            // which is bound by issue slots and the XU pipe)
            float4 u1 = ld_up(y0 + 1), o1 = ld_own(y0 + 1);
            float e1 = __ldg(a.oth + (size_t)(y0 + 1) * pitch + (COLOUR ? xe0 : xe1));
            for (int y = y0; y < y1; y += 2) {
                const int yn = min(y + 2, y1 - 2);
                const float4 nu0 = ld_up(yn), no0 = ld_own(yn), nu1 = ld_up(yn + 1), no1 = ld_own(yn + 1);
                const float ne0 = __ldg(a.oth + (size_t)yn * pitch + (COLOUR ? xe1 : xe0));
                const float ne1 = __ldg(a.oth + (size_t)(yn + 1) * pitch + (COLOUR ? xe0 : xe1));
                const uint4 C = make_uint4(0u, 0u, 0u, 0u);   // (over-relaxation draws no random numbers)
                xy_strip_row<OVERRELAX, MEASURE, COLOUR, RAGGED, 0>(a, y, (y + a.yoff) * a.gpr + g, C, nbl2e, u0, e0, o0, dn, mid, up,
                                                         a.own + (size_t)y * pitch, xi0, es, mx, my);
                xy_strip_row<OVERRELAX, MEASURE, COLOUR ^ 1, RAGGED, 1>(a, y + 1, (y + 1 + a.yoff) * a.gpr + g, C, nbl2e, u1, e1, o1, mid, up, dn,
                                                             a.own + (size_t)(y + 1) * pitch, xi0, es, mx, my);
                u0 = nu0; o0 = no0; e0 = ne0; u1 = nu1; o1 = no1; e1 = ne1;
                // rows rotate by two: (dn, mid, up) <- (up of the first row = mid of the second, up of the second)
                const XYRow t = mid; mid = dn; dn = up; (void)t;
            }
        } else {
            for (int y = y0; y < y1; y += 2) {
                const uint4 C = xy_pair_block(a, (uint64_t)((y + a.yoff) * a.gpr + g));   // the accept uniforms' low halves of rows y, y + 1
                const float4 u1 = ld_up(y + 1), o1 = ld_own(y + 1);
                const float e1 = __ldg(a.oth + (size_t)(y + 1) * pitch + (COLOUR ? xe0 : xe1));
                xy_strip_row<OVERRELAX, MEASURE, COLOUR, RAGGED, 0>(a, y, (y + a.yoff) * a.gpr + g, C, nbl2e, u0, e0, o0, dn, mid, up,
                                                         a.own + (size_t)y * pitch, xi0, es, mx, my);
                const int yn = min(y + 2, y1 - 2);
                u0 = ld_up(yn); o0 = ld_own(yn);
                e0 = __ldg(a.oth + (size_t)yn * pitch + (COLOUR ? xe1 : xe0));
                xy_strip_row<OVERRELAX, MEASURE, COLOUR ^ 1, RAGGED, 1>(a, y + 1, (y + 1 + a.yoff) * a.gpr + g, C, nbl2e, u1, e1, o1, mid, up, dn,
                                                             a.own + (size_t)(y + 1) * pitch, xi0, es, mx, my);
                const XYRow t = mid; mid = dn; dn = up; (void)t;
            }
        }
    }
    if (MEASURE) {
        double part[3] = {(double)es, (double)mx, (double)my};
        block_atomic_add_f64<3>(a.acc, part);
    }
}

// Reference-stream Metropolis pass: the accept uniforms r and the candidate uniforms c come from device arrays in the
// reference's own layout (randoms_(nx, ny), candidates_(nx, ny), column-major: element (x, y) at (y-1) nx + (x-1);
// src/xy2d_periodic_gpu_m.f90:355-356, read at :382-384), e.g. filled by cuRAND like the reference does.  Same fp32
// arithmetic as xy_strip_row (cos / sin by sincos_unit, same summation order), one thread per site.  Not the fast path.
__global__ void __launch_bounds__(256)
xy_pass_randoms_kernel(const __grid_constant__ XYArgs a, const double* __restrict__ randoms, const double* __restrict__ cands, int nx)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)a.ny * a.nxh) return;
    const int y = (int)(i / a.nxh), xi = (int)(i - (long long)y * a.nxh);
    const int P = (y + a.colour) & 1;            // lattice x of this site: 2 xi + P
    const int x0 = 2 * xi + P;
    const int xl = P ? xi : (xi == 0 ? a.nxh - 1 : xi - 1), xr = P ? (xi + 1 == a.nxh ? 0 : xi + 1) : xi;
    const int yu = (y + 1 >= a.ny && !a.halo) ? 0 : y + 1, yd = (y == 0 && !a.halo) ? a.ny - 1 : y - 1;
    float sr, cr, sl, cl, su, cu, sd, cd;
    sincos_unit(a.oth[(ptrdiff_t)y * a.pitch + xr], sr, cr);
    sincos_unit(a.oth[(ptrdiff_t)y * a.pitch + xl], sl, cl);
    sincos_unit(a.oth[(ptrdiff_t)yu * a.pitch + xi], su, cu);
    sincos_unit(a.oth[(ptrdiff_t)yd * a.pitch + xi], sd, cd);
    const float hx = cr + cl + cu + cd, hy = sr + sl + su + sd;
    const size_t ridx = (size_t)(y + a.yoff) * (size_t)nx + (size_t)x0;
    const float ct = (float)cands[ridx];
    const float ov = a.own[(size_t)y * a.pitch + xi];
    float cs, cc, ss, sc;
    sincos_unit(ct, cs, cc);
    sincos_unit(ov, ss, sc);
    const float de = (cc - sc) * hx + (cs - ss) * hy;   // -dE
    float w;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w) : "f"(de * (a.beta * 1.4426950408889634f)));
    if (!(randoms[ridx] > (double)w)) {
        a.own[(size_t)y * a.pitch + xi] = ct;   // (0, 1] turns, like the strip kernel
        if (xi < a.pitch - a.nxh) a.own[(size_t)y * a.pitch + a.nxh + xi] = ct;   // the row's mirror padding
    }
}

// metropolis_by_field_sub, :198-216 (initial-state preparation): every site, no coupling.
// candidate (cos 2 pi c, sin 2 pi c); dE = -(h . (cand - s)); accepted iff r <= 1 - exp(dE)
// (the reference's test is `randoms > 1 - exp(delta_energy) -> return`, :213).
// Uniforms: the Metropolis contract (same counters as the Metropolis kernel, both colours).
__global__ void __launch_bounds__(256)
xy_field_kernel(const __grid_constant__ XYArgs a, float hx, float hy)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.ny * a.gpr) return;
    const int y = idx / a.gpr, g = idx - y * a.gpr;
    float* prow = a.own + (size_t)y * a.pitch;
    const float4 o = *reinterpret_cast<const float4*>(prow + 4 * g);
    float ov[4] = {o.x, o.y, o.z, o.w};
    {
        const int yg = y + a.yoff;
        const uint4 C = xy_pair_block(a, (uint64_t)((yg & ~1) * a.gpr + g));
        float rr[4], cand[4];
        xy_group_uniforms(a, (uint64_t)(yg * a.gpr + g), C, yg & 1, rr, cand);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float cs, cc, ss, sc;
            sincos_turns(cand[j], cs, cc);
            sincos_turns(ov[j], ss, sc);
            const float de = -(hx * (cc - sc) + hy * (cs - ss));
            if (!(rr[j] > 1.0f - __expf(de))) ov[j] = cand[j];
        }
    }
    xy_store_group(prow, 4 * g, a.nxh, a.pitch, ov);
}

// fused E, Mx, My (three OpenACC reductions in the reference, :496-534), real64 accumulation.
// acc[0] += -sum s . (s_{x+1} + s_{y+1}), acc[1] += sum cos, acc[2] += sum sin
// Column strip like the update: a thread owns 8 consecutive lattice sites of a row (4 of each colour), walks
// XY_ROWS rows and keeps the previous row's cos/sin, so every site costs one sincos (+ 1/8 for the group edge)
// instead of three.  Bonds: (x, x+1) inside the row, (y-1, y) against the previous row.
__global__ void __launch_bounds__(256)
xy_measure_kernel(const float* __restrict__ c0, const float* __restrict__ c1, int nxh, int pitch, int ny, int gpr, double* acc, int halo)
{
    double part[3] = {0.0, 0.0, 0.0};
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nblk = (ny + XY_ROWS - 1) / XY_ROWS;
    if (tid < nblk * gpr) {
        const int rb = tid / gpr, g = tid - rb * gpr;
        const int y0 = rb * XY_ROWS, y1 = min(y0 + XY_ROWS, ny);
        // lattice sites of this group that exist: 8, or 2 (nxh - 4 g) in the last group of a row whose nx/2 is not a
        // multiple of 4; the site right of the last valid one is then the row's mirror padding (= x0 = 0)
        const int nv = min(8, 2 * (nxh - 4 * g));
        float pc[8], ps[8], qc[8], qs[8];
        float es = 0.f, mx = 0.f, my = 0.f;
        // cos / sin of the 8 sites x0 = 8 g .. 8 g + 7 of row y, in lattice order: even x0 belong to colour (y & 1)
        auto load_row = [&](int y, float (&c)[8], float (&sn)[8]) {
            const float4 a = *reinterpret_cast<const float4*>(((y & 1) ? c1 : c0) + (ptrdiff_t)y * pitch + 4 * g);
            const float4 b = *reinterpret_cast<const float4*>(((y & 1) ? c0 : c1) + (ptrdiff_t)y * pitch + 4 * g);
            sincos_turns(a.x, sn[0], c[0]); sincos_turns(b.x, sn[1], c[1]);
            sincos_turns(a.y, sn[2], c[2]); sincos_turns(b.y, sn[3], c[3]);
            sincos_turns(a.z, sn[4], c[4]); sincos_turns(b.z, sn[5], c[5]);
            sincos_turns(a.w, sn[6], c[6]); sincos_turns(b.w, sn[7], c[7]);
        };
        load_row((y0 == 0 && !halo) ? ny - 1 : y0 - 1, pc, ps);   // (halo: row -1 of both colours)
        for (int y = y0; y < y1; ++y) {
            load_row(y, qc, qs);
            // the site right of the group: x0 = 8 g + 8 (periodic), an even x0 -> colour (y & 1), xi = 4 g + 4
            const int xe = (4 * g + 4 >= nxh) ? 0 : 4 * g + 4;
            float ec, esn;
            sincos_turns(((y & 1) ? c1 : c0)[(size_t)y * pitch + xe], esn, ec);
            float e = nv == 8 ? qc[7] * ec + qs[7] * esn : 0.f;
#pragma unroll
            for (int k = 0; k < 7; ++k) if (k < nv) e += qc[k] * qc[k + 1] + qs[k] * qs[k + 1];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (k < nv) {
                    e += pc[k] * qc[k] + ps[k] * qs[k];
                    mx += qc[k];
                    my += qs[k];
                }
                pc[k] = qc[k];
                ps[k] = qs[k];
            }
            es -= e;
        }
        part[0] = (double)es; part[1] = (double)mx; part[2] = (double)my;
    }
    block_atomic_add_f64<3>(acc, part);
}

// autocorrelation with the stored snapshot (:536-549) and the fixed-distance correlation (:551-567)
// acc[0] += sum s . s0 ; acc[1] += sum s(x, y) . s(x + nx/2 - 1, y + ny/2 - 1)
__global__ void __launch_bounds__(256)
xy_corr_kernel(const float* __restrict__ c0, const float* __restrict__ c1, const float* __restrict__ z0,
               const float* __restrict__ z1, int nx, int ny, int pitch, double* acc)
{
    const int nxh = pitch;   // row stride of the colour arrays
    double part[2] = {0.0, 0.0};
    const long long total = (long long)nx * ny;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int y0 = (int)(i / nx), x0 = (int)(i - (long long)y0 * nx);
        const int col = (x0 + y0) & 1;
        const float t = (col ? c1 : c0)[(size_t)y0 * nxh + (x0 >> 1)];
        float s, c;
        sincos_turns(t, s, c);
        if (z0) {
            const float t0 = (col ? z1 : z0)[(size_t)y0 * nxh + (x0 >> 1)];
            float s0, cc0;
            sincos_turns(t0, s0, cc0);
            part[0] += (double)(c * cc0 + s * s0);
        }
        int xn = x0 + (nx / 2 - 1); if (xn >= nx) xn -= nx;
        int yn = y0 + (ny / 2 - 1); if (yn >= ny) yn -= ny;
        const int coln = (xn + yn) & 1;
        const float tn = (coln ? c1 : c0)[(size_t)yn * nxh + (xn >> 1)];
        float sn, cn;
        sincos_turns(tn, sn, cn);
        part[1] += (double)(c * cn + s * sn);
    }
    block_atomic_add_f64<2>(acc, part);
}

// set_random_spin_sub (:112-122): theta = 2 pi u  ->  turns = u
__global__ void __launch_bounds__(256)
xy_random_kernel(float* own, int nxh, int pitch, int ny, int gpr, int colour, uint32_t seed, uint64_t draw, int yoff)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ny * gpr) return;
    const int y = idx / gpr, g = idx - y * gpr;
    const uint4 R = philox4x32_10(mk_ctr((uint64_t)(idx + yoff * gpr), draw, (uint32_t)colour, 0u), make_uint2(seed, TAG_INIT));
    const float ov[4] = {((float)R.x + 1.0f) * 0x1p-32f, ((float)R.y + 1.0f) * 0x1p-32f, ((float)R.z + 1.0f) * 0x1p-32f, ((float)R.w + 1.0f) * 0x1p-32f};
    xy_store_group(own + (size_t)y * pitch, 4 * g, nxh, pitch, ov);
}

__global__ void xy_fill_kernel(float* a, float* b, size_t n, float v)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { a[i] = v; b[i] = v; }
}

// rotate_whole_spin_theta_sub (:281-293): add a constant angle (turns)
__global__ void xy_rotate_kernel(float* a, float* b, size_t n, float dt)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        float t = a[i] + dt; a[i] = t - floorf(t);
        t = b[i] + dt; b[i] = t - floorf(t);
    }
}

// spins() in the reference layout spins(0:nx+1, 0:ny+1, 1:2), real64, halo frame refreshed, corners 0
__global__ void xy_export_kernel(const float* c0, const float* c1, int nx, int ny, int pitch, double* out)
{
    const long long W = nx + 2, Ht = ny + 2;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W * Ht) return;
    const int y = (int)(i / W), x = (int)(i - (long long)y * W);
    const bool xh = (x == 0 || x == nx + 1), yh = (y == 0 || y == ny + 1);
    double c = 0.0, s = 0.0;
    if (!(xh && yh)) {
        const int x0 = x == 0 ? nx - 1 : (x == nx + 1 ? 0 : x - 1);
        const int y0 = y == 0 ? ny - 1 : (y == ny + 1 ? 0 : y - 1);
        const float t = (((x0 + y0) & 1) ? c1 : c0)[(size_t)y0 * pitch + (x0 >> 1)];
        sincospi(2.0 * (double)t, &s, &c);
    }
    out[i] = c;
    out[W * Ht + i] = s;
}
__global__ void xy_export_turns_kernel(const float* c0, const float* c1, int nx, int ny, int pitch, float* out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)nx * ny) return;
    const int y0 = (int)(i / nx), x0 = (int)(i - (long long)y0 * nx);
    out[i] = (((x0 + y0) & 1) ? c1 : c0)[(size_t)y0 * pitch + (x0 >> 1)];
}
__global__ void xy_import_turns_kernel(float* c0, float* c1, int nx, int ny, int pitch, const float* in)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)nx * ny) return;
    const int y0 = (int)(i / nx), x0 = (int)(i - (long long)y0 * nx);
    const float t = in[i];
    float* row = (((x0 + y0) & 1) ? c1 : c0) + (size_t)y0 * pitch;
    const float v = t - floorf(t);   // stored angles live in [0, 1] turns (sincos_unit)
    row[x0 >> 1] = v;
    if ((x0 >> 1) < pitch - nx / 2) row[nx / 2 + (x0 >> 1)] = v;   // the row's mirror padding
}

struct XY {
    int64_t nx, ny;
    int nxh, gpr, pitch;   // pitch: floats per row (nxh rounded up to 4)
    float* c[2];   // colour arrays: row 0 of the allocation below (one spare row before and after = the halo rows of slab mode)
    float* base[2];
    // slabs along y: this rank holds rows [yoff, yoff + ny) of ny_glob; ny is LOCAL everywhere below
    int rank, nranks;
    int64_t ny_glob, yoff;
    void* comm;
    int halo;      // 1: kernels read halo rows instead of wrapping the row index (B200MC_XY_HALO=1: single-GPU self-neighbour experiment)
    float* z[2];   // autocorrelation snapshot (allocated on first use)
    float* stage;  // nx*ny floats / export staging
    double* d_acc;
    cudaStream_t stream;
    double beta;
    uint32_t seed;
    uint64_t draw;
    bool obs_valid;
    double obs[3];
    int sms;
    // fused measurement (see the Ising handle): after a measured sweep the last colour pass accumulates E, Mx, My itself
    bool want_fused, fused_pending;
};

void fill_args(XY* m, int colour, XYArgs* a)
{
    a->own = m->c[colour]; a->oth = m->c[colour ^ 1];
    a->nxh = m->nxh; a->pitch = m->pitch; a->ny = (int)m->ny; a->gpr = m->gpr; a->colour = colour;
    a->beta = (float)m->beta; a->draw = m->draw;
    for (int r = 0; r < 10; ++r) a->rk0[r] = m->seed + (uint32_t)r * PHILOX_W0;
    a->acc = m->d_acc;
    a->halo = m->halo; a->yoff = (int)m->yoff;
}

// slab mode: rows -1 and ny of the colour just written.  Single GPU (self-neighbour): row ny - 1 and row 0 of the same
// array; between ranks these two copies become the send / receive of one row each way.
int halo_rows(XY* m, int colour)
{
    if (!m->halo) return B200MC_OK;
    float* c = m->c[colour];
    const size_t row = (size_t)m->pitch * sizeof(float);
    if (m->nranks > 1)   // my first row is the row above rank-1's last one, my last row the row below rank+1's first
        return dist_exchange_ring(m->comm, m->rank, m->nranks, c, c + (size_t)(m->ny - 1) * m->pitch, c - m->pitch, c + (size_t)m->ny * m->pitch, row, m->stream);
    CK(cudaMemcpyAsync(c - m->pitch, c + (size_t)(m->ny - 1) * m->pitch, row, cudaMemcpyDeviceToDevice, m->stream));
    CK(cudaMemcpyAsync(c + (size_t)m->ny * m->pitch, c, row, cudaMemcpyDeviceToDevice, m->stream));
    return B200MC_OK;
}

// (the ragged instantiation only for rows with a partial last group / mirror padding)
#define XY_STRIP_LAUNCH(OVR, MEAS, COL) do { \
        if (m->pitch != m->nxh) xy_strip_kernel<OVR, MEAS, COL, true><<<(strips + 255) / 256, 256, 0, m->stream>>>(a); \
        else xy_strip_kernel<OVR, MEAS, COL, false><<<(strips + 255) / 256, 256, 0, m->stream>>>(a); } while (0)

int sweep(XY* m)
{
    if (m->fused_pending) m->want_fused = false;   // the sums of the previous pass were never asked for
    m->obs_valid = false; m->fused_pending = false;
    const int strips = (int)((m->ny + XY_ROWS - 1) / XY_ROWS) * m->gpr;
    for (int colour = 0; colour < 2; ++colour) {
        XYArgs a;
        fill_args(m, colour, &a);
        const bool fuse = colour == 1 && m->want_fused;
        if (fuse) CK(cudaMemsetAsync(m->d_acc, 0, 3 * sizeof(double), m->stream));
        COUNT_LAUNCH();
        if (colour) {
            if (fuse) XY_STRIP_LAUNCH(false, true, 1);
            else XY_STRIP_LAUNCH(false, false, 1);
        } else XY_STRIP_LAUNCH(false, false, 0);   // (sums are fused into colour 1 only)
        CK(cudaGetLastError());
        { int rch = halo_rows(m, colour); if (rch) return rch; }
        if (fuse) m->fused_pending = true;
    }
    m->draw += 1;
    return B200MC_OK;
}

// One Metropolis MCS with the caller's uniforms instead of the built-in generator (parity / cuRAND-stream mode):
// randoms, candidates = host arrays of nall real64 in the reference's (nx, ny) order, :355-356
int sweep_with_randoms(XY* m, const double* randoms, const double* cands)
{
    if (!randoms || !cands) ARG_FAIL("null randoms / candidates");
    if (m->nranks > 1) { snprintf(g_b200mc_err, sizeof(g_b200mc_err), "xy2d update_with_randoms: not available in slab mode"); return B200MC_ERR_UNSUPPORTED; }
    const size_t n = (size_t)m->nx * m->ny;
    double* d = nullptr;
    CK(cudaMalloc(&d, 2 * n * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(d, randoms, n * sizeof(double), cudaMemcpyHostToDevice, m->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d + n, cands, n * sizeof(double), cudaMemcpyHostToDevice, m->stream);
    if (e != cudaSuccess) { cudaFree(d); CK(e); }
    m->obs_valid = false; m->fused_pending = false;
    const long long total = (long long)m->ny * m->nxh;
    int rc = B200MC_OK;
    for (int colour = 0; colour < 2 && !rc; ++colour) {
        XYArgs a;
        fill_args(m, colour, &a);
        COUNT_LAUNCH();
        xy_pass_randoms_kernel<<<(unsigned)((total + 255) / 256), 256, 0, m->stream>>>(a, d, d + n, (int)m->nx);
        if (cudaGetLastError() != cudaSuccess) { rc = B200MC_ERR_CUDA; snprintf(g_b200mc_err, sizeof(g_b200mc_err), "xy_pass_randoms_kernel launch failed"); break; }
        rc = halo_rows(m, colour);
    }
    cudaStreamSynchronize(m->stream);
    cudaFree(d);
    return rc;
}

int over_relax(XY* m, int n_steps)
{
    if (n_steps <= 0) return B200MC_OK;
    // (sums fused into a preceding Metropolis pass are simply superseded: update -> over-relaxation -> measure is the
    // drivers' order, app/xy2d_periodic_gpu_over_relaxation.f90:43-47)
    m->obs_valid = false; m->fused_pending = false;
    const int strips = (int)((m->ny + XY_ROWS - 1) / XY_ROWS) * m->gpr;
    for (int i = 0; i < n_steps; ++i)
        for (int colour = 0; colour < 2; ++colour) {
            XYArgs a;
            fill_args(m, colour, &a);
            const bool fuse = colour == 1 && i == n_steps - 1 && m->want_fused;
            if (fuse) CK(cudaMemsetAsync(m->d_acc, 0, 3 * sizeof(double), m->stream));
            COUNT_LAUNCH();
            if (colour) {
                if (fuse) XY_STRIP_LAUNCH(true, true, 1);
                else XY_STRIP_LAUNCH(true, false, 1);
            } else XY_STRIP_LAUNCH(true, false, 0);
            CK(cudaGetLastError());
            { int rch = halo_rows(m, colour); if (rch) return rch; }
            if (fuse) m->fused_pending = true;
        }
    return B200MC_OK;
}

int by_field(XY* m, double hx, double hy)
{
    m->obs_valid = false; m->fused_pending = false;
    const int total = (int)m->ny * m->gpr;
    for (int colour = 0; colour < 2; ++colour) {
        XYArgs a;
        fill_args(m, colour, &a);
        COUNT_LAUNCH();
        xy_field_kernel<<<(total + 255) / 256, 256, 0, m->stream>>>(a, (float)hx, (float)hy);
        CK(cudaGetLastError());
        { int rch = halo_rows(m, colour); if (rch) return rch; }
    }
    m->draw += 1;
    return B200MC_OK;
}

int measure(XY* m)
{
    if (m->obs_valid) return B200MC_OK;
    if (!m->fused_pending) {
        CK(cudaMemsetAsync(m->d_acc, 0, 3 * sizeof(double), m->stream));
        COUNT_LAUNCH();
        const int strips = (int)((m->ny + XY_ROWS - 1) / XY_ROWS) * m->gpr;
        xy_measure_kernel<<<(strips + 255) / 256, 256, 0, m->stream>>>(m->c[0], m->c[1], m->nxh, m->pitch, (int)m->ny, m->gpr, m->d_acc, m->halo);
        CK(cudaGetLastError());
    }
    m->fused_pending = false;
    m->want_fused = true;
    if (m->nranks > 1) { int rcr = dist_allreduce_f64(m->comm, m->d_acc, 3, m->stream); if (rcr) return rcr; }   // every rank gets the global sums
    CK(cudaMemcpyAsync(m->obs, m->d_acc, 3 * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    m->obs_valid = true;
    return B200MC_OK;
}

int corr(XY* m, bool autoc, double* out)
{
    if (autoc && !m->z[0]) ARG_FAIL("calc_autocorrelation_sum before set_initial_magne_autocorrelation_state");
    if (m->nranks > 1) { snprintf(g_b200mc_err, sizeof(g_b200mc_err), "xy2d: the correlation sums are not available in slab mode"); return B200MC_ERR_UNSUPPORTED; }
    CK(cudaMemsetAsync(m->d_acc, 0, 3 * sizeof(double), m->stream));
    COUNT_LAUNCH();
    xy_corr_kernel<<<m->sms * 8, 256, 0, m->stream>>>(m->c[0], m->c[1], autoc ? m->z[0] : nullptr, autoc ? m->z[1] : nullptr, (int)m->nx, (int)m->ny, m->pitch, m->d_acc);
    CK(cudaGetLastError());
    double r[2];
    CK(cudaMemcpyAsync(r, m->d_acc, 2 * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    m->obs_valid = false; m->fused_pending = false;  // d_acc reused
    *out = autoc ? r[0] : r[1];
    return B200MC_OK;
}

void destroy(XY* m)
{
    cudaStreamSynchronize(m->stream);
    if (m->comm) dist_comm_destroy(m->comm);
    cudaFree(m->base[0]); cudaFree(m->base[1]); cudaFree(m->z[0]); cudaFree(m->z[1]); cudaFree(m->stage); cudaFree(m->d_acc);
    delete m;
}

int fill(XY* m, float v)
{
    const size_t n = (size_t)m->pitch * m->ny;
    m->obs_valid = false; m->fused_pending = false;
    COUNT_LAUNCH();
    xy_fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, m->stream>>>(m->c[0], m->c[1], n, v);
    CK(cudaGetLastError());
    { int rch = halo_rows(m, 0); if (!rch) rch = halo_rows(m, 1); if (rch) return rch; }
    return B200MC_OK;
}

int ensure_stage(XY* m)
{
    if (!m->stage) CK(cudaMalloc(&m->stage, (size_t)2 * (m->nx + 2) * (m->ny + 2) * sizeof(double)));
    return B200MC_OK;
}

}  // namespace

#define HX(h) (reinterpret_cast<XY*>(h))
#define CHECK_X(h) do { if (!(h)) ARG_FAIL("invalid handle"); } while (0)

extern "C" {

int b200mc_xy2d_create(void** out, int64_t nx, int64_t ny, double kbt, int32_t iseed)
{
    return b200mc_xy2d_create_slab(out, nx, ny, kbt, iseed, 0, 1, nullptr);
}
int b200mc_xy2d_create_slab(void** out, int64_t nx, int64_t ny_global, double kbt, int32_t iseed, int32_t rank, int32_t nranks, const char* nccl_id)
{
    if (!out) ARG_FAIL("null handle pointer");
    *out = nullptr;
    if (nranks < 1 || rank < 0 || rank >= nranks) ARG_FAIL("bad rank %d / %d", rank, nranks);
    if (nranks > 1 && !nccl_id) ARG_FAIL("slab mode needs the NCCL unique id of the job (b200mc_dist_unique_id on rank 0, broadcast by the caller)");
    if (ny_global % nranks || ((ny_global / nranks) & 1)) ARG_FAIL("xy2d slabs: ny (%lld) must split into an even number of rows per rank (%d ranks)", (long long)ny_global, nranks);
    const int64_t ny = ny_global / nranks;
    if (!(kbt > 0.0)) ARG_FAIL("kbt must be > 0");
    // the reference needs nx, ny even (colouring, src/xy2d_periodic_gpu_m.f90:377-380); so does this build (rows are padded to whole float4 groups)
    if (nx < 8 || (nx & 1) || ny < 2 || (ny & 1)) ARG_FAIL("xy2d: need nx >= 8 even, ny >= 2 even (got %lld x %lld)", (long long)nx, (long long)ny);
    if ((nx / 2 + 3) / 4 * ny >= (int64_t)0x7FFFFFFF) ARG_FAIL("xy2d: lattice too large");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        snprintf(g_b200mc_err, sizeof(g_b200mc_err), "no CUDA device: this library has no CPU fallback");
        return B200MC_ERR_CUDA;
    }
    XY* m = new (std::nothrow) XY();
    if (!m) ARG_FAIL("out of host memory");
    m->nx = nx; m->ny = ny; m->nxh = (int)(nx / 2); m->gpr = (m->nxh + 3) / 4; m->pitch = 4 * m->gpr; m->stream = 0;
    m->beta = 1 / kbt; m->seed = (uint32_t)iseed; m->draw = 0; m->obs_valid = false; m->fused_pending = false; m->want_fused = false;
    m->c[0] = m->c[1] = m->base[0] = m->base[1] = m->z[0] = m->z[1] = m->stage = nullptr; m->d_acc = nullptr; m->halo = 0; m->comm = nullptr; m->nranks = 1; m->rank = 0;
    int dev = 0; m->sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&m->sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t n = (size_t)m->pitch * ny;
    { const char* t = getenv("B200MC_XY_HALO"); m->halo = ((t && atoi(t)) || nranks > 1) ? 1 : 0; }
    m->rank = rank; m->nranks = nranks; m->ny_glob = ny_global; m->yoff = (int64_t)rank * ny; m->comm = nullptr;
    if (ny_global * ((nx / 2 + 3) / 4) >= (int64_t)0x7FFFFFFF) { delete m; ARG_FAIL("xy2d: lattice too large for the 32-bit RNG block index"); }
    if (nranks > 1) {
        int rcc = dist_comm_init(&m->comm, rank, nranks, nccl_id);
        if (rcc) { delete m; return rcc; }
    }
    const size_t nalloc = n + 2 * (size_t)m->pitch;
    if (cudaMalloc(&m->base[0], nalloc * sizeof(float)) != cudaSuccess || cudaMalloc(&m->base[1], nalloc * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&m->d_acc, 3 * sizeof(double)) != cudaSuccess) {
        snprintf(g_b200mc_err, sizeof(g_b200mc_err), "cudaMalloc failed");
        destroy(m); return B200MC_ERR_CUDA;
    }
    m->c[0] = m->base[0] + m->pitch; m->c[1] = m->base[1] + m->pitch;
    cudaMemsetAsync(m->base[0], 0, nalloc * sizeof(float), m->stream);
    cudaMemsetAsync(m->base[1], 0, nalloc * sizeof(float), m->stream);
    int rc = fill(m, 0.0f);  // set_allup_spin: all along +x
    if (rc) { destroy(m); return rc; }
    *out = m;
    return B200MC_OK;
}
int b200mc_xy2d_destroy(void* h) { if (h) destroy(HX(h)); return B200MC_OK; }
int b200mc_xy2d_set_stream(void* h, void* s) { CHECK_X(h); HX(h)->stream = (cudaStream_t)s; return B200MC_OK; }
int b200mc_xy2d_skip_curand(void* h, int64_t n)
{
    CHECK_X(h);
    if (n < 0) ARG_FAIL("n_skip < 0");
    const int64_t per = HX(h)->nx * HX(h)->ny_glob;  // one generate call draws nall uniforms
    HX(h)->draw += (uint64_t)((n + per - 1) / per);
    return B200MC_OK;
}
int b200mc_xy2d_set_allup_spin(void* h) { CHECK_X(h); return fill(HX(h), 0.0f); }
int b200mc_xy2d_set_random_spin(void* h)
{
    CHECK_X(h);
    XY* m = HX(h);
    m->obs_valid = false; m->fused_pending = false;
    const int total = (int)m->ny * m->gpr;
    for (int c = 0; c < 2; ++c) {
        COUNT_LAUNCH();
        xy_random_kernel<<<(total + 255) / 256, 256, 0, m->stream>>>(m->c[c], m->nxh, m->pitch, (int)m->ny, m->gpr, c, m->seed, m->draw, (int)m->yoff);
        CK(cudaGetLastError());
        { int rch = halo_rows(m, c); if (rch) return rch; }
    }
    m->draw += 1;
    return B200MC_OK;
}
int b200mc_xy2d_set_kbt(void* h, double kbt) { CHECK_X(h); if (!(kbt > 0.0)) ARG_FAIL("kbt must be > 0"); HX(h)->beta = 1 / kbt; return B200MC_OK; }
int b200mc_xy2d_set_beta(void* h, double beta) { CHECK_X(h); HX(h)->beta = beta; return B200MC_OK; }
int b200mc_xy2d_update(void* h) { CHECK_X(h); return sweep(HX(h)); }
int b200mc_xy2d_update_with_randoms(void* h, const double* randoms, const double* candidates) { CHECK_X(h); return sweep_with_randoms(HX(h), randoms, candidates); }
int b200mc_xy2d_update_n(void* h, int32_t n) { CHECK_X(h); for (int i = 0; i < n; ++i) { int rc = sweep(HX(h)); if (rc) return rc; } return B200MC_OK; }
int b200mc_xy2d_update_over_relaxation(void* h, int32_t n_steps) { CHECK_X(h); return over_relax(HX(h), n_steps); }
int b200mc_xy2d_calc_energy_sum(void* h, double* e) { CHECK_X(h); int rc = measure(HX(h)); if (rc) return rc; *e = HX(h)->obs[0]; return B200MC_OK; }
int b200mc_xy2d_calc_magne_sum(void* h, double* m) { CHECK_X(h); int rc = measure(HX(h)); if (rc) return rc; *m = HX(h)->obs[1]; return B200MC_OK; }
int b200mc_xy2d_calc_magne_y_sum(void* h, double* m) { CHECK_X(h); int rc = measure(HX(h)); if (rc) return rc; *m = HX(h)->obs[2]; return B200MC_OK; }
int b200mc_xy2d_measure(void* h, double* e, double* mx, double* my)
{
    CHECK_X(h);
    int rc = measure(HX(h));
    if (rc) return rc;
    if (e) *e = HX(h)->obs[0];
    if (mx) *mx = HX(h)->obs[1];
    if (my) *my = HX(h)->obs[2];
    return B200MC_OK;
}
int b200mc_xy2d_set_initial_magne_autocorrelation_state(void* h)
{
    CHECK_X(h);
    XY* m = HX(h);
    const size_t n = (size_t)m->pitch * m->ny;
    if (!m->z[0]) { CK(cudaMalloc(&m->z[0], n * sizeof(float))); CK(cudaMalloc(&m->z[1], n * sizeof(float))); }
    CK(cudaMemcpyAsync(m->z[0], m->c[0], n * sizeof(float), cudaMemcpyDeviceToDevice, m->stream));
    CK(cudaMemcpyAsync(m->z[1], m->c[1], n * sizeof(float), cudaMemcpyDeviceToDevice, m->stream));
    return B200MC_OK;
}
int b200mc_xy2d_calc_autocorrelation_sum(void* h, double* r) { CHECK_X(h); return corr(HX(h), true, r); }
int b200mc_xy2d_calc_correlation_sum(void* h, double* r) { CHECK_X(h); return corr(HX(h), false, r); }
// rotate_summation_magne_toward_xaxis (:219-232): theta = atan2(My, Mx); rotate every spin by -theta
int b200mc_xy2d_rotate_summation_magne_toward_xaxis(void* h, int32_t with_autocorrelation)
{
    CHECK_X(h);
    XY* m = HX(h);
    int rc = measure(m);
    if (rc) return rc;
    const double theta = atan2(m->obs[2], m->obs[1]);
    const float dt = (float)(-theta / (2 * 3.14159265358979323846));
    const size_t n = (size_t)m->pitch * m->ny;
    m->obs_valid = false; m->fused_pending = false;
    COUNT_LAUNCH();
    xy_rotate_kernel<<<(unsigned)((n + 255) / 256), 256, 0, m->stream>>>(m->c[0], m->c[1], n, dt);
    if (with_autocorrelation && m->z[0]) {
        COUNT_LAUNCH();
        xy_rotate_kernel<<<(unsigned)((n + 255) / 256), 256, 0, m->stream>>>(m->z[0], m->z[1], n, dt);
    }
    CK(cudaGetLastError());
    { int rch = halo_rows(m, 0); if (!rch) rch = halo_rows(m, 1); if (rch) return rch; }
    return B200MC_OK;
}
int b200mc_xy2d_metropolis_by_field(void* h, double hx, double hy) { CHECK_X(h); return by_field(HX(h), hx, hy); }

namespace {
enum { PREP_FINITE = 0, PREP_SMALL = 1, PREP_NEAR = 2 };
// the three initial-state loops (:126-196): random start, then field sweeps until |m| meets the
// criterion, then rotate the total magnetisation onto the x axis
int prepare(XY* m, int kind, double target, double pct)
{
    int rc = b200mc_xy2d_set_random_spin(m);
    if (rc) return rc;
    double field_x = 1.0;
    for (int it = 0;; ++it) {
        if ((rc = measure(m))) return rc;
        const double n = (double)(m->nx * m->ny_glob);
        const double mx = m->obs[1] / n, my = m->obs[2] / n, mabs = hypot(mx, my);
        if (kind == PREP_FINITE) {
            if (fabs(mabs - target) / target < 1e-2) break;   // epsilon, :130
            if (mabs > target) field_x = -field_x / 2; else field_x = field_x * 2;
        } else if (kind == PREP_SMALL) {
            if (mabs < target) break;
        } else {
            if (fabs(mabs - target) / target <= pct) break;
        }
        if (it >= 4096) {   // the reference loops for ever when its heuristic cycles (it does for most targets of set_finite_magne_spin)
            snprintf(g_b200mc_err, sizeof(g_b200mc_err), "xy2d: initial-state loop did not reach |m| = %g after %d field sweeps (last %g)", target, it, mabs);
            return B200MC_ERR_STATE;
        }
        rc = kind == PREP_FINITE ? by_field(m, field_x, 0.0) : by_field(m, -mx, -my);
        if (rc) return rc;
    }
    return b200mc_xy2d_rotate_summation_magne_toward_xaxis(m, 0);
}
}  // namespace
int b200mc_xy2d_set_finite_magne_spin(void* h, double init_magne) { CHECK_X(h); if (!(init_magne > 0.0)) ARG_FAIL("init_magne must be > 0"); return prepare(HX(h), PREP_FINITE, init_magne, 0.0); }
int b200mc_xy2d_set_random_small_spin(void* h, double near_magne) { CHECK_X(h); if (!(near_magne > 0.0)) ARG_FAIL("near_magne must be > 0"); return prepare(HX(h), PREP_SMALL, near_magne, 0.0); }
int b200mc_xy2d_set_random_near_spin(void* h, double near_magne, double diff_parcent) { CHECK_X(h); if (!(near_magne > 0.0)) ARG_FAIL("near_magne must be > 0"); return prepare(HX(h), PREP_NEAR, near_magne, diff_parcent); }

int b200mc_xy2d_get_spins(void* h, double* out)
{
    CHECK_X(h);
    if (!out) ARG_FAIL("null output");
    XY* m = HX(h);
    int rc = ensure_stage(m);
    if (rc) return rc;
    const long long n = (long long)(m->nx + 2) * (m->ny + 2);
    COUNT_LAUNCH();
    xy_export_kernel<<<(unsigned)((n + 255) / 256), 256, 0, m->stream>>>(m->c[0], m->c[1], (int)m->nx, (int)m->ny, m->pitch, reinterpret_cast<double*>(m->stage));
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, m->stage, (size_t)2 * n * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    return B200MC_OK;
}
// angles in turns (theta / 2 pi), fp32, row-major [ny][nx]: the build's native state (exact round trip)
int b200mc_xy2d_get_angles(void* h, float* out)
{
    CHECK_X(h);
    if (!out) ARG_FAIL("null output");
    XY* m = HX(h);
    int rc = ensure_stage(m);
    if (rc) return rc;
    const long long n = (long long)m->nx * m->ny;
    COUNT_LAUNCH();
    xy_export_turns_kernel<<<(unsigned)((n + 255) / 256), 256, 0, m->stream>>>(m->c[0], m->c[1], (int)m->nx, (int)m->ny, m->pitch, m->stage);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, m->stage, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    return B200MC_OK;
}
int b200mc_xy2d_set_angles(void* h, const float* in)
{
    CHECK_X(h);
    if (!in) ARG_FAIL("null input");
    XY* m = HX(h);
    int rc = ensure_stage(m);
    if (rc) return rc;
    const long long n = (long long)m->nx * m->ny;
    m->obs_valid = false; m->fused_pending = false;
    CK(cudaMemcpyAsync(m->stage, in, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, m->stream));
    COUNT_LAUNCH();
    xy_import_turns_kernel<<<(unsigned)((n + 255) / 256), 256, 0, m->stream>>>(m->c[0], m->c[1], (int)m->nx, (int)m->ny, m->pitch, m->stage);
    CK(cudaGetLastError());
    { int rch = halo_rows(m, 0); if (!rch) rch = halo_rows(m, 1); if (rch) return rch; }
    CK(cudaStreamSynchronize(m->stream));
    return B200MC_OK;
}
int64_t b200mc_xy2d_nx(void* h) { return h ? HX(h)->nx : -1; }
int64_t b200mc_xy2d_ny(void* h) { return h ? HX(h)->ny : -1; }
int64_t b200mc_xy2d_nall(void* h) { return h ? HX(h)->nx * HX(h)->ny_glob : -1; }   // (slab mode: ny() is the local row count, nall() the whole lattice)
double b200mc_xy2d_kbt(void* h) { return h ? 1 / HX(h)->beta : 0.0; }
double b200mc_xy2d_beta(void* h) { return h ? HX(h)->beta : 0.0; }
int b200mc_xy2d_sync(void* h) { CHECK_X(h); CK(cudaStreamSynchronize(HX(h)->stream)); return B200MC_OK; }

}  // extern "C"
