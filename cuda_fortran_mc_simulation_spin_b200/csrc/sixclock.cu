// Periodic q-state clock with the full q^6 acceptance table ("tableall") and its compact
// two-colour storage ("dual lattice"): kernels, host-side handle and C ABI.
// Reference: module clock_tableall_gpu_m, src/clock/clock_tableall_gpu_m.f90:1-182, and module
// clock_dual_lattice_tableall_gpu_m, src/clock/clock_dual_lattice_tableall_m.f90:1-202 (the
// y-compacted twin clock_dual_lattice_yhalf_tableall_m.f90 differs in index arithmetic only).
//
// Layout.  The reference stores sixclock(nx, ny) int32 (tableall) or two int32 colour arrays
// sixclock_even/odd(nx/2, ny) (dual lattice) plus rnds(2, nx, ny) real64 = 20 B/site.  Here a
// state is one byte, colours split as in the dual-lattice module: colour c = (x0 + y0) & 1,
// compact index xi = x0 >> 1 (x0 = 2 xi + ((y0 + c) & 1)), rows padded to a multiple of 16
// bytes, replicas ("multi-sample batch") stacked: [replica][y0][pitch].  A thread owns one
// aligned 16-byte vector of one row; its neighbours are three aligned 16-byte vectors of the
// other colour (same row, row above, row below) plus one byte of the adjacent vector; the
// periodic wrap is index arithmetic (no halo).  Pad bytes (xi >= nx/2) always hold valid states
// and are masked out of the observables.
//
// RNG contract (v3; CPU restatement: oracle/rng_contract.c, orc_torus_uniforms; bit assignment:
// clock_word.cuh).  Vector v = xi >> 4 of row y0: Philox block counter = sample << 32 | (y0 nvr + v),
// key (seed, TAG_TORUS), sub-counters 0..2 (first look) and 4..6 (low halves of the accept uniforms);
// site j = xi & 15 = word j >> 2, e = j & 3.  The accept uniform is looked at in two stages (15
// bits, then all 32: the second stage only for a tie), the four proposals of a word are the
// leading base-(q-1) digits of one 32-bit uniform:
//   new = c + ceiling(rnds1 (q-1))  (:142)   <=>  k - 1 = floor(W_e (q-1) / 2^32)
//   rnds2 <= prob                  (:146)   <=>  U_a < thr = floor(prob 2^32)
#include <math.h>
#include <stdlib.h>
#include <map>
#include <new>
#include <vector>
#include "../../include/b200mc.h"
#include "common.cuh"
#include "clock_word.cuh"
#include "ising_kernels.cuh"   // philox_rk

namespace {

#define SIX_MAX_CLASSES 256

struct SixArgs {
    uint8_t* own;            // colour being updated, all replicas
    const uint8_t* oth;
    int nxh, ny, nvr;        // sites per row per colour, rows, 16-byte vectors per row
    int nrows;               // n_multi * ny
    int colour;
    uint32_t sample0;        // global index of this handle's first sample (a batch split across GPUs)
    uint32_t q;
    const uint8_t* cls;      // q^6 class ids: c + q(new + q(r + q(u + q(l + q d))))   (states_to_prob index order, :72-80)
    const uint32_t* thi;     // per class: thr >> 16  (0 .. 65536)
    const uint32_t* tlo;     // per class: thr & 0xFFFF
    const uint16_t* thr16;   // direct lookup (q <= 6): T[k][F] = thr >> 17, k = 0 .. q-2 (new = c + 1 + k mod q), F = r + q u + q^2 l + q^3 d + q^4 c
    uint32_t tab_bytes, q5;
    int cls_in_smem;
    int prefetch;            // L2 prefetch of the next iteration's vectors (B200MC_SIX_PREFETCH, A/B)
    uint64_t draw;
    uint32_t rk0[10];        // Philox round keys seed + r W0
};

__device__ __forceinline__ uint32_t ldg_u8(const uint8_t* p)
{
    uint32_t v;
    asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t word_of(const uint4& v, int w) { return w == 0 ? v.x : w == 1 ? v.y : w == 2 ? v.z : v.w; }

// the other colour's same-row values one compact position to the left (p == 0: x0 - 1) or to
// the right (p == 1: x0 + 1) of the 16 sites of vector v, periodic in x.  pv = address of vector v of the row.
__device__ __forceinline__ uint4 six_shifted(const uint8_t* pv, const uint4& b, int v, int nvr, int nxh, int p)
{
    uint4 s;
    if (p == 0) {
        const uint32_t e = ldg_u8(pv + (v == 0 ? nxh - 1 : -1));
        s.x = (b.x << 8) | e;
        s.y = __funnelshift_l(b.x, b.y, 8);
        s.z = __funnelshift_l(b.y, b.z, 8);
        s.w = __funnelshift_l(b.z, b.w, 8);
    } else {
        const bool last = v == nvr - 1;
        const uint32_t e = ldg_u8(pv + (last ? -16 * v : 16));
        s.x = __funnelshift_r(b.x, b.y, 8);
        s.y = __funnelshift_r(b.y, b.z, 8);
        s.z = __funnelshift_r(b.z, b.w, 8);
        s.w = (b.w >> 8) | (e << 24);
        if (last && (nxh & 15)) {  // the row's last site sits at local position lp < 15 (padded row)
            const int lp = nxh - 1 - 16 * v, wi = lp >> 2, sh = 8 * (lp & 3);
            const uint32_t keep = ~(0xFFu << sh), ins = e << sh;
            if (wi == 0) s.x = (s.x & keep) | ins;
            if (wi == 1) s.y = (s.y & keep) | ins;
            if (wi == 2) s.z = (s.z & keep) | ins;
            if (wi == 3) s.w = (s.w & keep) | ins;
        }
    }
    return s;
}

// update_sub, src/clock/clock_tableall_gpu_m.f90:107-152 (dual lattice: :110-155), one colour
#ifndef SIX_MINB
#define SIX_MINB 3
#endif
// Acceptance lookup.  Class path (any q <= 12): q^6 one-byte class ids (shared or global memory) + per-class thresholds,
// every site with its full 32-bit uniforms.  Direct path (q <= 6, default): clock_word.cuh.
struct SixSmem {
    uint32_t cls_addr;        // shared-window address of the class table
    const uint32_t* sthi;     // per class thr >> 16 (shared memory; class path)
    const uint8_t* gcls;
};

struct SixRows { uint4 o, rt, up, lf, dn; uint4* po; uint64_t blk; };   // blk: Philox block counter = (sample << 32) | (y nvr + v)

// loads of one vector = 16 sites of row (rep, y) at compact position 16 v .. 16 v + 15; idx = (rep ny + y) nvr + v is its
// linear index in both colour arrays (rows are nvr vectors long).  P = (y + colour) & 1: the x position of compact site
// xi is 2 xi + P; right = x0 + 1, left = x0 - 1.
template <int P>
__device__ __forceinline__ void six_load(const SixArgs& a, int idx, int y, int rep, int v, SixRows& n)
{
    const int nvr = a.nvr, ny = a.ny;
    const int wrap = (ny - 1) * nvr;
    const int du = (y + 1 == ny) ? -wrap : nvr, dd = (y == 0) ? wrap : -nvr;
    const uint4* pv = reinterpret_cast<const uint4*>(a.oth) + idx;
    n.po = reinterpret_cast<uint4*>(a.own) + idx;
    n.o = *n.po;
    const uint4 b = ld_other(pv);
    n.up = ld_other(pv + du);
    n.dn = ld_other(pv + dd);
    const uint4 s = six_shifted(reinterpret_cast<const uint8_t*>(pv), b, v, nvr, a.nxh, P);
    n.rt = P ? s : b; n.lf = P ? b : s;
    n.blk = ((uint64_t)(a.sample0 + (uint32_t)rep) << 32) | (uint32_t)(y * nvr + v);
}

// exact evaluation of the 16 sites of a vector (RNG contract v3, clock_word.cuh): both stages of the accept uniforms, class
// table, full 33-bit threshold.  MODE 0: class ids and thresholds from global memory (the direct path's tie redo), 1: class
// ids from shared memory, 2: class ids from global memory, thresholds from shared memory.
template <int MODE>
__device__ __forceinline__ uint4 six_vector_exact_body(const SixArgs& a, const SixSmem& sm, const SixRows& n)
{
    const uint32_t q = a.q, qm1 = q - 1;
    uint32_t X[12], Y[12];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const uint4 R = philox_rk<TAG_TORUS>(mk_ctr(n.blk, a.draw, (uint32_t)a.colour, (uint32_t)i), a.rk0);
        const uint4 R2 = philox_rk<TAG_TORUS>(mk_ctr(n.blk, a.draw, (uint32_t)a.colour, 4u + (uint32_t)i), a.rk0);
        X[4 * i] = R.x; X[4 * i + 1] = R.y; X[4 * i + 2] = R.z; X[4 * i + 3] = R.w;
        Y[4 * i] = R2.x; Y[4 * i + 1] = R2.y; Y[4 * i + 2] = R2.z; Y[4 * i + 3] = R2.w;
    }
    uint32_t outw[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        const uint32_t ow = word_of(n.o, w), rt = word_of(n.rt, w), up = word_of(n.up, w), lf = word_of(n.lf, w), dn = word_of(n.dn, w);
        uint32_t W = X[3 * w + 2], res = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const uint32_t Ua = clk_accept32(X[3 * w + (e & 1)], Y[3 * w + (e & 1)], e >> 1);
            uint32_t k;
            mulwide(W, qm1, W, k);                                  // ceiling(rnds1 (q - 1)) - 1, :142
            const uint32_t c = (ow >> (8 * e)) & 0xFFu, r = (rt >> (8 * e)) & 0xFFu, u = (up >> (8 * e)) & 0xFFu,
                           l = (lf >> (8 * e)) & 0xFFu, d = (dn >> (8 * e)) & 0xFFu;
            uint32_t nw = c + 1u + k;
            if (nw >= q) nw -= q;
            const uint32_t idx = c + q * (nw + q * (r + q * (u + q * (l + q * d))));
            uint32_t cl;
            if (MODE == 1) asm("ld.shared.u8 %0, [%1];" : "=r"(cl) : "r"(sm.cls_addr + idx));
            else cl = a.cls[idx];
            const unsigned long long thr = ((unsigned long long)(MODE == 0 ? a.thi[cl] : sm.sthi[cl]) << 16) | a.tlo[cl];
            res |= (((unsigned long long)Ua < thr) ? nw : c) << (8 * e);
        }
        outw[w] = res;
    }
    return make_uint4(outw[0], outw[1], outw[2], outw[3]);
}
// (arguments and result by value: a struct passed by reference would live in local memory on the hot path)
__device__ __noinline__ uint4 six_vector_exact(const SixArgs& a, uint64_t blk, uint4 o, uint4 rt, uint4 up, uint4 lf, uint4 dn)
{
    SixSmem sm;
    sm.cls_addr = 0; sm.sthi = nullptr; sm.gcls = a.cls;
    SixRows n;
    n.o = o; n.rt = rt; n.up = up; n.lf = lf; n.dn = dn; n.po = nullptr; n.blk = blk;
    return six_vector_exact_body<0>(a, sm, n);
}

template <bool SMEM, int P>
__device__ __forceinline__ void six_vector(const SixArgs& a, const SixSmem& sm, int idx, int y, int rep, int v)
{
    SixRows n;
    six_load<P>(a, idx, y, rep, v, n);
    *n.po = six_vector_exact_body<SMEM ? 1 : 2>(a, sm, n);
}

// direct path: table T[k][F] of the thresholds' high 15 bits in shared memory, F = r + q u + q^2 l + q^3 d + q^4 c
// Q = q at compile time (0: run time)
template <int P, int Q>
__device__ __forceinline__ void six_vector_direct(const SixArgs& a, const uint8_t* tab, int idx, int y, int rep, int v)
{
    const uint32_t q = Q ? (uint32_t)Q : a.q, qm1 = q - 1, kstride = Q ? 2u * Q * Q * Q * Q * Q : 2u * a.q5;
    SixRows n;
    six_load<P>(a, idx, y, rep, v, n);
    uint32_t X[12];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const uint4 R = philox_rk<TAG_TORUS>(mk_ctr(n.blk, a.draw, (uint32_t)a.colour, (uint32_t)i), a.rk0);
        X[4 * i] = R.x; X[4 * i + 1] = R.y; X[4 * i + 2] = R.z; X[4 * i + 3] = R.w;
    }
    uint4 out;
    uint32_t outw[4], amin = 0x7FFF7FFFu;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        uint32_t Fe, Fo;
        clk_index_fields(word_of(n.rt, w), word_of(n.up, w), word_of(n.lf, w), word_of(n.dn, w), word_of(n.o, w), q, Fe, Fo);
        outw[w] = clk_word_fast<true>(word_of(n.o, w), Fe, Fo, X[3 * w], X[3 * w + 1], X[3 * w + 2], tab, kstride, qm1, q, amin);
    }
    out = make_uint4(outw[0], outw[1], outw[2], outw[3]);
    if (clk_accept_tie(amin)) out = six_vector_exact(a, n.blk, n.o, n.rt, n.up, n.lf, n.dn);   // rare (2^-15 per site): redo the vector exactly
    *n.po = out;
}

#define SIX_DIRECT_THREADS 768
template <bool SMEM, bool DIRECT = false, int Q = 0, int THREADS = SIX_DIRECT_THREADS>
__global__ void __launch_bounds__(DIRECT ? THREADS : 256, DIRECT ? 1 : SIX_MINB)
sixclock_pass_kernel(const __grid_constant__ SixArgs a)
{
    extern __shared__ __align__(16) uint8_t smraw[];
    // class path: SIX_MAX_CLASSES words of thresholds, then q^6 class bytes (SMEM); direct path: the u16 table
    // T[k][F] (2 q^5 (q - 1) bytes)
    uint32_t* sthi = reinterpret_cast<uint32_t*>(smraw);
    uint8_t* scls = smraw + (DIRECT ? 0 : SIX_MAX_CLASSES * sizeof(uint32_t));
    if (!DIRECT) for (int i = threadIdx.x; i < SIX_MAX_CLASSES; i += blockDim.x) sthi[i] = a.thi[i];
    if (SMEM || DIRECT) {
        const uint4* src = reinterpret_cast<const uint4*>(DIRECT ? reinterpret_cast<const uint8_t*>(a.thr16) : a.cls);
        uint4* dst = reinterpret_cast<uint4*>(scls);
        const uint32_t bytes = DIRECT ? 2u * a.q5 * (a.q - 1) : a.tab_bytes;
        for (uint32_t i = threadIdx.x; i < (bytes + 15) / 16; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    SixSmem sm;
    sm.cls_addr = (uint32_t)__cvta_generic_to_shared(scls);
    sm.sthi = sthi;
    sm.gcls = a.cls;
    const int nvr = a.nvr, ny = a.ny;
    // linear vector index idx = (replica ny + y) nvr + v of this thread; (y, v, replica) advanced incrementally: no
    // division in the loop
    const int stride = gridDim.x * blockDim.x;
    const int dY = stride / nvr, dv = stride - dY * nvr;
    const int total = a.nrows * nvr;
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int Y = idx / nvr, v = idx - Y * nvr;
    int rep = Y / ny, y = Y - rep * ny;
    while (idx < total) {
        if (DIRECT && a.prefetch && (threadIdx.x & 7) == 0 && idx + stride < total) {
            // the two streamed vectors of this thread's NEXT iteration into L2 (one lane per 128-byte line); the rows above
            // and below are the same lines other threads of the wave stream as their own row
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const uint4*>(a.own) + idx + stride));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const uint4*>(a.oth) + idx + stride));
        }
        if (DIRECT) {
            if ((y + a.colour) & 1) six_vector_direct<1, Q>(a, scls, idx, y, rep, v);
            else six_vector_direct<0, Q>(a, scls, idx, y, rep, v);
        } else {
            if ((y + a.colour) & 1) six_vector<SMEM, 1>(a, sm, idx, y, rep, v);
            else six_vector<SMEM, 0>(a, sm, idx, y, rep, v);
        }
        idx += stride;
        v += dv; y += dY;
        if (v >= nvr) { v -= nvr; ++y; }
        while (y >= ny) { y -= ny; ++rep; }
    }
}

// reference-stream pass: rnds(2, nx, ny) real64 from a device array, real64 compare against the
// real64 table, exactly like update_sub (:142-150)
__global__ void __launch_bounds__(256)
sixclock_pass_rnds_kernel(const __grid_constant__ SixArgs a, const double* __restrict__ rnds, const double* __restrict__ prob, int rep)
{
    const int nx = 2 * a.nxh;
    const long long total = (long long)a.nxh * a.ny;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int y = (int)(i / a.nxh), xi = (int)(i - (long long)y * a.nxh);
    const size_t pitch = (size_t)a.nvr * 16;
    const int Y = rep * a.ny + y;
    const int p = (y + a.colour) & 1;
    const int x0 = 2 * xi + p;
    const int yu = (y + 1 == a.ny) ? 0 : y + 1, yd = (y == 0) ? a.ny - 1 : y - 1;
    const uint8_t* base = a.oth + (size_t)rep * a.ny * pitch;
    const int xr = (x0 + 1 == nx) ? 0 : x0 + 1, xl = (x0 == 0) ? nx - 1 : x0 - 1;
    const uint32_t q = a.q;
    const uint32_t r = base[(size_t)y * pitch + (xr >> 1)], l = base[(size_t)y * pitch + (xl >> 1)];
    const uint32_t u = base[(size_t)yu * pitch + xi], d = base[(size_t)yd * pitch + xi];
    uint8_t* po = a.own + (size_t)Y * pitch + xi;
    const uint32_t c = *po;
    const size_t at = (size_t)x0 + (size_t)nx * y;
    int nw = (int)c + (int)ceil(rnds[2 * at] * (double)(q - 1));
    if (nw >= (int)q) nw -= (int)q;
    if (nw < 0 || nw >= (int)q) return;  // rnds outside (0, 1]
    const size_t ix = (size_t)c + q * ((size_t)nw + q * ((size_t)r + q * ((size_t)u + q * ((size_t)l + (size_t)q * d))));
    if (rnds[2 * at + 1] <= prob[ix]) *po = (uint8_t)nw;
}

// (a - b) mod q per byte, a, b in [0, q)
__device__ __forceinline__ uint32_t six_submod(uint32_t a, uint32_t b, uint32_t q)
{
    const uint32_t d = a + q * 0x01010101u - b;
    const uint32_t ge = ((d + (0x80u - q) * 0x01010101u) >> 7) & 0x01010101u;
    return d - ge * q;
}
__device__ __forceinline__ void six_onehot8(uint32_t w0, uint32_t w1, uint32_t keep0, uint32_t keep1, uint32_t& hA, uint32_t& hB)
{
    const uint32_t pk = w0 + (w1 << 4);
    hA = prmt(0x08040201u, 0x80402010u, pk) & prmt(keep0, keep1, 0x5140u);
    hB = prmt(0x08040201u, 0x80402010u, pk >> 16) & prmt(keep0, keep1, 0x7362u);
}

// Exact integer observables per replica (calc_magne :155-165 and calc_energy :167-181 only depend
// on them): acc[rep*192 + c] += #{state c}; acc[rep*192 + 64 + d] += #{(s(x+1, y) - s(x, y)) mod q = d};
// acc[rep*192 + 128 + d] += same for (x, y+1).   q <= 8: SWAR one-hot + POPC; else shared atomics.
__global__ void __launch_bounds__(256)
sixclock_measure_kernel(const uint8_t* __restrict__ c0, const uint8_t* __restrict__ c1, int nxh, int ny, int nvr, int rep,
                        uint32_t q, unsigned long long* acc)
{
    __shared__ unsigned long long sacc[3 * 64];
    for (int i = threadIdx.x; i < 3 * 64; i += blockDim.x) sacc[i] = 0;
    __syncthreads();
    uint32_t cnt[3][8];
#pragma unroll
    for (int h = 0; h < 3; ++h)
#pragma unroll
        for (int c = 0; c < 8; ++c) cnt[h][c] = 0;
    const size_t pitch = (size_t)nvr * 16;
    const int total = ny * nvr;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int y = idx / nvr, v = idx - y * nvr;
        const int Y = rep * ny + y, Yu = rep * ny + ((y + 1 == ny) ? 0 : y + 1);
        uint32_t keep[4];
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            keep[w] = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (16 * v + 4 * w + j < nxh) keep[w] |= 0xFFu << (8 * j);
        }
#pragma unroll
        for (int colour = 0; colour < 2; ++colour) {
            const uint8_t* own = colour ? c1 : c0;
            const uint8_t* oth = colour ? c0 : c1;
            const int p = (y + colour) & 1;
            const uint8_t* row = oth + (size_t)Y * pitch;
            const uint4 o = *(reinterpret_cast<const uint4*>(own + (size_t)Y * pitch) + v);
            const uint4 b = *(reinterpret_cast<const uint4*>(row) + v);
            const uint4 u = *(reinterpret_cast<const uint4*>(oth + (size_t)Yu * pitch) + v);
            const uint4 r = p ? six_shifted(row + 16 * (size_t)v, b, v, nvr, nxh, 1) : b;   // x0 + 1
            const uint32_t ow[4] = {o.x, o.y, o.z, o.w}, rw[4] = {r.x, r.y, r.z, r.w}, uw[4] = {u.x, u.y, u.z, u.w};
            if (q <= 8) {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    uint32_t hs[2], hr[2], hu[2];
                    six_onehot8(ow[2 * g], ow[2 * g + 1], keep[2 * g], keep[2 * g + 1], hs[0], hs[1]);
                    six_onehot8(six_submod(rw[2 * g], ow[2 * g], q), six_submod(rw[2 * g + 1], ow[2 * g + 1], q),
                                keep[2 * g], keep[2 * g + 1], hr[0], hr[1]);
                    six_onehot8(six_submod(uw[2 * g], ow[2 * g], q), six_submod(uw[2 * g + 1], ow[2 * g + 1], q),
                                keep[2 * g], keep[2 * g + 1], hu[0], hu[1]);
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const uint32_t m = 0x01010101u << c;
                        cnt[0][c] += __popc(hs[0] & m) + __popc(hs[1] & m);
                        cnt[1][c] += __popc(hr[0] & m) + __popc(hr[1] & m);
                        cnt[2][c] += __popc(hu[0] & m) + __popc(hu[1] & m);
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    if (!((keep[j >> 2] >> (8 * (j & 3))) & 1u)) continue;
                    const uint32_t s = (ow[j >> 2] >> (8 * (j & 3))) & 0xFFu;
                    const uint32_t rr = (rw[j >> 2] >> (8 * (j & 3))) & 0xFFu;
                    const uint32_t uu = (uw[j >> 2] >> (8 * (j & 3))) & 0xFFu;
                    uint32_t dr = rr + q - s; if (dr >= q) dr -= q;
                    uint32_t du = uu + q - s; if (du >= q) du -= q;
                    atomicAdd(&sacc[s], 1ull);
                    atomicAdd(&sacc[64 + dr], 1ull);
                    atomicAdd(&sacc[128 + du], 1ull);
                }
            }
        }
    }
    if (q <= 8) {
#pragma unroll
        for (int h = 0; h < 3; ++h)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                uint32_t t = cnt[h][c];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
                if ((threadIdx.x & 31) == 0 && t) atomicAdd(&sacc[h * 64 + c], (unsigned long long)t);
            }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * 64; i += blockDim.x)
        if (sacc[i]) atomicAdd(&acc[(size_t)rep * 192 + i], sacc[i]);
}

// sixclock(nx, ny) int32 (column-major: x fastest) <-> the two byte colour arrays of one replica
__global__ void sixclock_export_kernel(const uint8_t* c0, const uint8_t* c1, int nx, int ny, size_t pitch, int32_t* out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)nx * ny) return;
    const int y0 = (int)(i / nx), x0 = (int)(i - (long long)y0 * nx);
    out[i] = (((x0 + y0) & 1) ? c1 : c0)[(size_t)y0 * pitch + (x0 >> 1)];
}
__global__ void sixclock_import_kernel(uint8_t* c0, uint8_t* c1, int nx, int ny, size_t pitch, const int32_t* in)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)nx * ny) return;
    const int y0 = (int)(i / nx), x0 = (int)(i - (long long)y0 * nx);
    (((x0 + y0) & 1) ? c1 : c0)[(size_t)y0 * pitch + (x0 >> 1)] = (uint8_t)in[i];
}
// sixclock_even / sixclock_odd (nx/2, ny) int32 <-> one colour array of one replica
__global__ void sixclock_export_half_kernel(const uint8_t* c, int nxh, int ny, size_t pitch, int32_t* out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)nxh * ny) return;
    const int y0 = (int)(i / nxh), xi = (int)(i - (long long)y0 * nxh);
    out[i] = c[(size_t)y0 * pitch + xi];
}
__global__ void sixclock_import_half_kernel(uint8_t* c, int nxh, int ny, size_t pitch, const int32_t* in)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)nxh * ny) return;
    const int y0 = (int)(i / nxh), xi = (int)(i - (long long)y0 * nxh);
    c[(size_t)y0 * pitch + xi] = (uint8_t)in[i];
}

struct Six {
    int64_t nx, ny;
    int32_t q, n_multi;
    int32_t sample0;   // this handle holds samples sample0 .. sample0 + n_multi - 1 of the job
    int variant;                 // 0: tableall / table delta-E expression, 1: clock_simple's
    int nxh, nvr;
    size_t pitch, rep_bytes;     // bytes per row / per replica per colour
    uint8_t* c[2];
    cudaStream_t stream;
    double beta;
    uint32_t seed;
    uint64_t draw;
    std::vector<double> magne, e3, prob;  // host tables exactly as the reference builds them
    uint8_t* d_cls;
    uint32_t *d_thi, *d_tlo;
    uint16_t* d_thr16;   // direct lookup table (q^6 entries)
    int prefetch;
    int direct, threads; // direct: one-load lookup with SIX_DIRECT_THREADS-thread blocks, one per SM
    double* d_prob;
    double* d_rnds;
    int32_t* d_stage;
    unsigned long long* d_acc;
    int grid, smem_bytes, cls_in_smem, sms;
    bool obs_valid;
    std::vector<long long> obs;  // n_multi x 192
    // optional per-launch timing of the pass kernel
    bool timing;
    std::vector<cudaEvent_t> evs;
    size_t ev_used;
};

int build_tables(Six* m)
{
    const int q = m->q;
    const double pi = 4 * atan(1.0);
    const double psi = 2 * pi / q;  // pi_state_inv, :11
    const size_t q3 = (size_t)q * q * q, q6 = q3 * q3;
    m->magne.resize(q); m->e3.resize(q3); m->prob.resize(q6);
    for (int c = 0; c < q; ++c) m->magne[c] = cos(c * psi);  // state_to_magne, :26
    // state_center_right_up_to_energy, :27-33: the constructor runs global_c fastest, then global_u,
    // then global_r, so element (i1, i2, i3) was computed with global_c = i1, global_u = i2, global_r = i3
    for (int i3 = 0; i3 < q; ++i3)
        for (int i2 = 0; i2 < q; ++i2)
            for (int i1 = 0; i1 < q; ++i1)
                m->e3[i1 + q * (i2 + q * i3)] = -cos((i2 - i1) * psi) - cos((i3 - i1) * psi);
#define E3(i1, i2, i3) m->e3[(i1) + q * ((i2) + q * (i3))]
    std::map<uint64_t, int> classes;
    std::vector<uint8_t> cls(q6);
    std::vector<uint32_t> thi, tlo;
    // init_sixclock, :66-86 (same loop nest, same expression order)
    for (int d = 0; d < q; ++d)
        for (int l = 0; l < q; ++l)
            for (int u = 0; u < q; ++u)
                for (int r = 0; r < q; ++r)
                    for (int nc = 0; nc < q; ++nc)
                        for (int c = 0; c < q; ++c) {
                            double de;
                            if (m->variant == 0) {
                                de = E3(nc, r, u) - E3(c, r, u) + E3(nc, l, d) - E3(c, l, d);
                            } else {
                                // clock_simple_gpu_m update_sub, src/clock/clock_simple_gpu_m.f90:108-113: no tables, the
                                // four neighbours in the order right, left, up, down (nearest_spins(1:4), :83-99)
                                const int nbv[4] = {r, l, u, d};
                                de = 0.0;
                                for (int i = 0; i < 4; ++i) de = de + (-cos((nbv[i] - nc) * psi) + cos((nbv[i] - c) * psi));
                            }
                            // tableall: prob = 1 if dE <= 0 (:76-80); table / simple: accepted outright unless dE > 0
                            // (clock_table_gpu_m.f90:123-125) -- the same predicate
                            const double w = (de <= 0.0) ? 1.0 : exp(-m->beta * de);
                            const size_t at = (size_t)c + (size_t)q * (nc + (size_t)q * (r + (size_t)q * (u + (size_t)q * (l + (size_t)q * d))));
                            m->prob[at] = w;
                            uint64_t t = (uint64_t)floor(w * 4294967296.0);  // rnds2 <= w  <=>  U_a < floor(w 2^32)
                            if (t > 4294967296ull) t = 4294967296ull;
                            auto it = classes.find(t);
                            int id;
                            if (it == classes.end()) {
                                id = (int)thi.size();
                                if (id >= SIX_MAX_CLASSES) {
                                    snprintf(g_b200mc_err, sizeof(g_b200mc_err), "sixclock: more than %d distinct acceptance thresholds (mstate = %d)", SIX_MAX_CLASSES, q);
                                    return B200MC_ERR_UNSUPPORTED;
                                }
                                classes[t] = id;
                                thi.push_back((uint32_t)(t >> 16));
                                tlo.push_back((uint32_t)(t & 0xFFFFu));
                            } else id = it->second;
                            cls[at] = (uint8_t)id;
                        }
#undef E3
    thi.resize(SIX_MAX_CLASSES, 0); tlo.resize(SIX_MAX_CLASSES, 0);
    // direct table (q <= 6): T[k][F] = thr >> 17 (0 .. 32768), k = 0 .. q - 2 (new = c + 1 + k mod q),
    // F = r + q u + q^2 l + q^3 d + q^4 c (the last q^5 entries of the buffer are unused)
    std::vector<uint16_t> thr16(q6, 0);
    {
        const size_t q5 = q6 / q;
        for (int k = 0; k + 1 < q; ++k)
            for (int c = 0; c < q; ++c)
                for (int d = 0; d < q; ++d)
                    for (int l = 0; l < q; ++l)
                        for (int u = 0; u < q; ++u)
                            for (int r = 0; r < q; ++r) {
                                const int nc = (c + 1 + k) % q;
                                const size_t at = (size_t)c + (size_t)q * (nc + (size_t)q * (r + (size_t)q * (u + (size_t)q * (l + (size_t)q * d))));
                                const size_t F = (size_t)r + (size_t)q * (u + (size_t)q * (l + (size_t)q * (d + (size_t)q * c)));
                                const uint64_t t = ((uint64_t)thi[cls[at]] << 16) | tlo[cls[at]];
                                thr16[(size_t)k * q5 + F] = (uint16_t)(t >> 17);
                            }
    }
    CK(cudaMemcpyAsync(m->d_thr16, thr16.data(), q6 * sizeof(uint16_t), cudaMemcpyHostToDevice, m->stream));
    CK(cudaMemcpyAsync(m->d_cls, cls.data(), q6, cudaMemcpyHostToDevice, m->stream));
    CK(cudaMemcpyAsync(m->d_thi, thi.data(), SIX_MAX_CLASSES * sizeof(uint32_t), cudaMemcpyHostToDevice, m->stream));
    CK(cudaMemcpyAsync(m->d_tlo, tlo.data(), SIX_MAX_CLASSES * sizeof(uint32_t), cudaMemcpyHostToDevice, m->stream));
    if (m->d_prob) CK(cudaMemcpyAsync(m->d_prob, m->prob.data(), q6 * sizeof(double), cudaMemcpyHostToDevice, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    return B200MC_OK;
}

void fill_args(Six* m, int colour, SixArgs* a)
{
    a->own = m->c[colour]; a->oth = m->c[colour ^ 1];
    a->nxh = m->nxh; a->ny = (int)m->ny; a->nvr = m->nvr; a->nrows = (int)(m->n_multi * m->ny);
    a->colour = colour; a->q = (uint32_t)m->q; a->sample0 = (uint32_t)m->sample0;
    a->cls = m->d_cls; a->thi = m->d_thi; a->tlo = m->d_tlo; a->thr16 = m->d_thr16;
    a->tab_bytes = (uint32_t)m->prob.size(); a->q5 = a->tab_bytes / (uint32_t)m->q; a->cls_in_smem = m->cls_in_smem;
    a->draw = m->draw;
    a->prefetch = m->prefetch;
    for (int r = 0; r < 10; ++r) a->rk0[r] = m->seed + (uint32_t)r * PHILOX_W0;
}

int sweep(Six* m)
{
    m->obs_valid = false;
    for (int colour = 0; colour < 2; ++colour) {  // parity 0 = even sites first, :96-101
        SixArgs a;
        fill_args(m, colour, &a);
        if (m->timing) {
            while (m->evs.size() < m->ev_used + 2) { cudaEvent_t e; CK(cudaEventCreate(&e)); m->evs.push_back(e); }
            CK(cudaEventRecord(m->evs[m->ev_used], m->stream));
        }
        COUNT_LAUNCH();
        if (m->direct && m->q == 6 && m->threads == 1024) sixclock_pass_kernel<false, true, 6, 1024><<<m->grid, 1024, m->smem_bytes, m->stream>>>(a);
        else if (m->direct && m->q == 6) sixclock_pass_kernel<false, true, 6><<<m->grid, SIX_DIRECT_THREADS, m->smem_bytes, m->stream>>>(a);
        else if (m->direct) sixclock_pass_kernel<false, true><<<m->grid, SIX_DIRECT_THREADS, m->smem_bytes, m->stream>>>(a);
        else if (m->cls_in_smem) sixclock_pass_kernel<true><<<m->grid, 256, m->smem_bytes, m->stream>>>(a);
        else sixclock_pass_kernel<false><<<m->grid, 256, m->smem_bytes, m->stream>>>(a);
        CK(cudaGetLastError());
        if (m->timing) { CK(cudaEventRecord(m->evs[m->ev_used + 1], m->stream)); m->ev_used += 2; }
    }
    m->draw += 1;
    return B200MC_OK;
}

int update_with_rnds(Six* m, const double* rnds)
{
    if (!rnds) ARG_FAIL("null rnds");
    const size_t N = (size_t)m->nx * m->ny, q6 = m->prob.size();
    if (!m->d_prob) {
        CK(cudaMalloc(&m->d_prob, q6 * sizeof(double)));
        CK(cudaMemcpy(m->d_prob, m->prob.data(), q6 * sizeof(double), cudaMemcpyHostToDevice));
        CK(cudaMalloc(&m->d_rnds, 2 * N * sizeof(double)));
    }
    m->obs_valid = false;
    const unsigned grid = (unsigned)(((size_t)m->nxh * m->ny + 255) / 256);
    for (int rep = 0; rep < m->n_multi; ++rep) {
        CK(cudaMemcpyAsync(m->d_rnds, rnds + (size_t)rep * 2 * N, 2 * N * sizeof(double), cudaMemcpyHostToDevice, m->stream));
        for (int colour = 0; colour < 2; ++colour) {
            SixArgs a;
            fill_args(m, colour, &a);
            COUNT_LAUNCH();
            sixclock_pass_rnds_kernel<<<grid, 256, 0, m->stream>>>(a, m->d_rnds, m->d_prob, rep);
            CK(cudaGetLastError());
        }
        CK(cudaStreamSynchronize(m->stream));
    }
    return B200MC_OK;
}

int measure(Six* m)
{
    if (m->obs_valid) return B200MC_OK;
    CK(cudaMemsetAsync(m->d_acc, 0, (size_t)m->n_multi * 192 * sizeof(unsigned long long), m->stream));
    const int need = (int)(((size_t)m->ny * m->nvr + 255) / 256);
    const int grid = need < m->sms * 8 ? need : m->sms * 8;
    for (int rep = 0; rep < m->n_multi; ++rep) {
        COUNT_LAUNCH();
        sixclock_measure_kernel<<<grid, 256, 0, m->stream>>>(m->c[0], m->c[1], m->nxh, (int)m->ny, m->nvr, rep, (uint32_t)m->q, m->d_acc);
        CK(cudaGetLastError());
    }
    m->obs.resize((size_t)m->n_multi * 192);
    CK(cudaMemcpyAsync(m->obs.data(), m->d_acc, m->obs.size() * sizeof(long long), cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    m->obs_valid = true;
    return B200MC_OK;
}

void destroy(Six* m)
{
    cudaStreamSynchronize(m->stream);
    cudaFree(m->c[0]); cudaFree(m->c[1]); cudaFree(m->d_cls); cudaFree(m->d_thi); cudaFree(m->d_tlo); cudaFree(m->d_thr16);
    cudaFree(m->d_prob); cudaFree(m->d_rnds); cudaFree(m->d_stage); cudaFree(m->d_acc);
    for (cudaEvent_t e : m->evs) cudaEventDestroy(e);
    delete m;
}

int stage(Six* m)
{
    if (!m->d_stage) CK(cudaMalloc(&m->d_stage, (size_t)m->nx * m->ny * sizeof(int32_t)));
    return B200MC_OK;
}

}  // namespace

#define HS(h) (reinterpret_cast<Six*>(h))
#define CHECK_S(h) do { if (!(h)) ARG_FAIL("invalid handle"); } while (0)

extern "C" {

int b200mc_sixclock_create(void** out, int64_t nx, int64_t ny, double kbt, int32_t mstate, int32_t n_multi, int32_t iseed)
{
    return b200mc_sixclock_create_variant(out, nx, ny, kbt, mstate, n_multi, iseed, 0);
}
int b200mc_sixclock_create_variant(void** out, int64_t nx, int64_t ny, double kbt, int32_t mstate, int32_t n_multi, int32_t iseed,
                                   int32_t variant)
{
    if (variant != 0 && variant != 1) ARG_FAIL("sixclock: unknown variant %d", variant);
    if (!out) ARG_FAIL("null handle pointer");
    *out = nullptr;
    if (!(kbt > 0.0)) ARG_FAIL("kbt must be > 0");
    if (mstate < 2) ARG_FAIL("mstate must be >= 2");
    if (mstate > 12) { snprintf(g_b200mc_err, sizeof(g_b200mc_err), "sixclock: mstate = %d not supported (q^6 table; max 12)", mstate); return B200MC_ERR_UNSUPPORTED; }
    if (n_multi < 1) ARG_FAIL("n_multi must be >= 1");
    // (x + y) parity colouring on a torus needs both extents even (the reference races otherwise)
    if (nx < 2 || ny < 2 || (nx & 1) || (ny & 1)) ARG_FAIL("sixclock: nx and ny must be even and >= 2 (got %lld x %lld)", (long long)nx, (long long)ny);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        snprintf(g_b200mc_err, sizeof(g_b200mc_err), "no CUDA device: this library has no CPU fallback");
        return B200MC_ERR_CUDA;
    }
    const int64_t nxh = nx / 2, nvr = (nxh + 15) / 16;
    if ((double)n_multi * (double)ny * (double)nvr >= 2147483000.0) ARG_FAIL("sixclock: lattice x batch too large for 32-bit vector indices");
    Six* m = new (std::nothrow) Six();
    if (!m) ARG_FAIL("out of host memory");
    m->sample0 = 0;
    m->nx = nx; m->ny = ny; m->q = mstate; m->n_multi = n_multi; m->nxh = (int)nxh; m->nvr = (int)nvr; m->variant = variant;
    m->pitch = (size_t)nvr * 16; m->rep_bytes = m->pitch * (size_t)ny;
    m->stream = 0; m->seed = (uint32_t)iseed; m->draw = 0; m->beta = 1 / kbt; m->obs_valid = false;
    m->c[0] = m->c[1] = nullptr; m->d_cls = nullptr; m->d_thi = m->d_tlo = nullptr; m->d_thr16 = nullptr; m->d_prob = nullptr; m->d_rnds = nullptr;
    m->d_stage = nullptr; m->d_acc = nullptr; m->timing = false; m->ev_used = 0;
    const size_t q6 = (size_t)mstate * mstate * mstate * mstate * mstate * mstate;
    const size_t bytes = m->rep_bytes * (size_t)n_multi;
    if (cudaMalloc(&m->c[0], bytes) != cudaSuccess || cudaMalloc(&m->c[1], bytes) != cudaSuccess ||
        cudaMalloc(&m->d_cls, (q6 + 15) / 16 * 16) != cudaSuccess || cudaMalloc(&m->d_thi, SIX_MAX_CLASSES * sizeof(uint32_t)) != cudaSuccess ||
        cudaMalloc(&m->d_tlo, SIX_MAX_CLASSES * sizeof(uint32_t)) != cudaSuccess || cudaMalloc(&m->d_thr16, (2 * q6 + 15) / 16 * 16) != cudaSuccess ||
        cudaMalloc(&m->d_acc, (size_t)n_multi * 192 * sizeof(unsigned long long)) != cudaSuccess) {
        snprintf(g_b200mc_err, sizeof(g_b200mc_err), "cudaMalloc failed (%zu bytes per colour)", bytes);
        cudaGetLastError();
        destroy(m); return B200MC_ERR_CUDA;
    }
    cudaMemsetAsync(m->c[0], 0, bytes, m->stream);   // init_sixclock_order: all states 0 (pad bytes too)
    cudaMemsetAsync(m->c[1], 0, bytes, m->stream);
    const size_t want = SIX_MAX_CLASSES * sizeof(uint32_t) + (q6 + 15) / 16 * 16;
    int dev = 0, maxsm = 0, occ = 1;
    m->sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&m->sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    m->cls_in_smem = want <= (size_t)maxsm ? 1 : 0;
    m->smem_bytes = (int)(m->cls_in_smem ? want : SIX_MAX_CLASSES * sizeof(uint32_t));
    if ((m->cls_in_smem ? cudaFuncSetAttribute(sixclock_pass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, m->smem_bytes)
                        : cudaFuncSetAttribute(sixclock_pass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, m->smem_bytes)) != cudaSuccess) {
        snprintf(g_b200mc_err, sizeof(g_b200mc_err), "cudaFuncSetAttribute(smem) failed");
        destroy(m); return B200MC_ERR_CUDA;
    }
    if (m->cls_in_smem) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sixclock_pass_kernel<true>, 256, m->smem_bytes);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sixclock_pass_kernel<false>, 256, m->smem_bytes);
    if (occ < 1) occ = 1;
    const int64_t need = ((int64_t)n_multi * ny * nvr + 255) / 256;
    m->grid = (int)(need < (int64_t)m->sms * occ ? need : (int64_t)m->sms * occ);
    // direct lookup: 2 q^6 bytes of thresholds in shared memory, one block of SIX_DIRECT_THREADS threads per SM
    m->direct = 0; m->threads = 256;
    { const char* t = getenv("B200MC_SIX_PREFETCH"); m->prefetch = (t && atoi(t) == 1) ? 1 : 0; }
    {
        const size_t wantd = (2 * q6 + 15) / 16 * 16;
        const char* t = getenv("B200MC_SIX_DIRECT");
        int occd = 0;
        if (wantd <= (size_t)maxsm && !(t && atoi(t) == 0) &&
            cudaFuncSetAttribute(sixclock_pass_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wantd) == cudaSuccess &&
            cudaFuncSetAttribute(sixclock_pass_kernel<false, true, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wantd) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occd, sixclock_pass_kernel<false, true>, SIX_DIRECT_THREADS, wantd) == cudaSuccess && occd >= 1) {
            m->direct = 1; m->threads = SIX_DIRECT_THREADS; m->smem_bytes = (int)wantd;
            // q = 6: 1024-thread blocks (64 registers) measured 4-6 % faster than 768 x 80 (873 / 951 vs 843 / 893 flips/ns at
            // 16384^2 x 1 / x 4 samples); B200MC_SIX_THREADS=768 selects the other instantiation for A/B runs
            const char* tt = getenv("B200MC_SIX_THREADS");
            if (!(tt && atoi(tt) == 768) && mstate == 6 &&
                cudaFuncSetAttribute(sixclock_pass_kernel<false, true, 6, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wantd) == cudaSuccess) m->threads = 1024;
            const int64_t needd = ((int64_t)n_multi * ny * nvr + m->threads - 1) / m->threads;
            m->grid = (int)(needd < (int64_t)m->sms * occd ? needd : (int64_t)m->sms * occd);
        } else cudaGetLastError();
    }
    int rc = build_tables(m);
    if (rc) { destroy(m); return rc; }
    *out = m;
    return B200MC_OK;
}
int b200mc_sixclock_destroy(void* h) { if (h) destroy(HS(h)); return B200MC_OK; }
int b200mc_sixclock_set_stream(void* h, void* s) { CHECK_S(h); HS(h)->stream = (cudaStream_t)s; return B200MC_OK; }
int b200mc_sixclock_skip_curand_clock(void* h, int64_t n_skip)
{
    CHECK_S(h);
    if (n_skip < 0) ARG_FAIL("n_skip < 0");
    const int64_t per = 2 * HS(h)->nx * HS(h)->ny;  // uniforms per update_metropolis, :95
    HS(h)->draw += (uint64_t)((n_skip + per - 1) / per);
    return B200MC_OK;
}
int b200mc_sixclock_init_sixclock_order(void* h)
{
    CHECK_S(h);
    Six* m = HS(h);
    m->obs_valid = false;
    const size_t bytes = m->rep_bytes * (size_t)m->n_multi;
    CK(cudaMemsetAsync(m->c[0], 0, bytes, m->stream));
    CK(cudaMemsetAsync(m->c[1], 0, bytes, m->stream));
    return B200MC_OK;
}
int b200mc_sixclock_set_kbt(void* h, double kbt) { CHECK_S(h); if (!(kbt > 0.0)) ARG_FAIL("kbt must be > 0"); HS(h)->beta = 1 / kbt; return build_tables(HS(h)); }
int b200mc_sixclock_update_metropolis(void* h) { CHECK_S(h); return sweep(HS(h)); }
int b200mc_sixclock_update_metropolis_n(void* h, int32_t n) { CHECK_S(h); for (int i = 0; i < n; ++i) { int rc = sweep(HS(h)); if (rc) return rc; } return B200MC_OK; }
int b200mc_sixclock_update_with_rnds(void* h, const double* rnds) { CHECK_S(h); return update_with_rnds(HS(h), rnds); }
int b200mc_sixclock_get_histograms(void* h, int64_t* hist, int64_t* bond_right, int64_t* bond_up)
{
    CHECK_S(h);
    Six* m = HS(h);
    int rc = measure(m);
    if (rc) return rc;
    for (int j = 0; j < m->n_multi; ++j)
        for (int c = 0; c < m->q; ++c) {
            if (hist) hist[j * m->q + c] = m->obs[(size_t)j * 192 + c];
            if (bond_right) bond_right[j * m->q + c] = m->obs[(size_t)j * 192 + 64 + c];
            if (bond_up) bond_up[j * m->q + c] = m->obs[(size_t)j * 192 + 128 + c];
        }
    return B200MC_OK;
}
int b200mc_sixclock_calc_energy(void* h, double* res)
{
    CHECK_S(h);
    Six* m = HS(h);
    if (!res) ARG_FAIL("null output");
    int rc = measure(m);
    if (rc) return rc;
    const double nall_inv = 1.0 / (double)(m->nx * m->ny);  // :13
    for (int j = 0; j < m->n_multi; ++j) {
        // sum of state_center_right_up_to_energy(c, r, u) = -cos((r - c) psi) - cos((u - c) psi), :177
        double e = 0.0;
        for (int d = 0; d < m->q; ++d)
            e -= (double)(m->obs[(size_t)j * 192 + 64 + d] + m->obs[(size_t)j * 192 + 128 + d]) * m->magne[d];
        res[j] = e * nall_inv;
    }
    return B200MC_OK;
}
int b200mc_sixclock_calc_magne(void* h, double* res)
{
    CHECK_S(h);
    Six* m = HS(h);
    if (!res) ARG_FAIL("null output");
    int rc = measure(m);
    if (rc) return rc;
    const double nall_inv = 1.0 / (double)(m->nx * m->ny);
    for (int j = 0; j < m->n_multi; ++j) {
        double s = 0.0;
        for (int c = 0; c < m->q; ++c) s += (double)m->obs[(size_t)j * 192 + c] * m->magne[c];
        res[j] = s * nall_inv;
    }
    return B200MC_OK;
}
int b200mc_sixclock_get_sixclock(void* h, int32_t* out)
{
    CHECK_S(h);
    if (!out) ARG_FAIL("null output");
    Six* m = HS(h);
    int rc = stage(m);
    if (rc) return rc;
    const size_t N = (size_t)m->nx * m->ny;
    for (int j = 0; j < m->n_multi; ++j) {
        COUNT_LAUNCH();
        sixclock_export_kernel<<<(unsigned)((N + 255) / 256), 256, 0, m->stream>>>(m->c[0] + j * m->rep_bytes, m->c[1] + j * m->rep_bytes, (int)m->nx, (int)m->ny, m->pitch, m->d_stage);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(out + (size_t)j * N, m->d_stage, N * sizeof(int32_t), cudaMemcpyDeviceToHost, m->stream));
        CK(cudaStreamSynchronize(m->stream));
    }
    return B200MC_OK;
}
int b200mc_sixclock_set_sixclock(void* h, const int32_t* in)
{
    CHECK_S(h);
    if (!in) ARG_FAIL("null input");
    Six* m = HS(h);
    const size_t N = (size_t)m->nx * m->ny;
    for (size_t i = 0; i < N * (size_t)m->n_multi; ++i)
        if (in[i] < 0 || in[i] >= m->q) ARG_FAIL("sixclock: state %d at element %zu outside 0..%d", in[i], i, m->q - 1);
    int rc = stage(m);
    if (rc) return rc;
    m->obs_valid = false;
    for (int j = 0; j < m->n_multi; ++j) {
        CK(cudaMemcpyAsync(m->d_stage, in + (size_t)j * N, N * sizeof(int32_t), cudaMemcpyHostToDevice, m->stream));
        COUNT_LAUNCH();
        sixclock_import_kernel<<<(unsigned)((N + 255) / 256), 256, 0, m->stream>>>(m->c[0] + j * m->rep_bytes, m->c[1] + j * m->rep_bytes, (int)m->nx, (int)m->ny, m->pitch, m->d_stage);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(m->stream));
    }
    return B200MC_OK;
}
int b200mc_sixclock_get_dual(void* h, int32_t* even, int32_t* odd)
{
    CHECK_S(h);
    if (!even || !odd) ARG_FAIL("null output");
    Six* m = HS(h);
    int rc = stage(m);
    if (rc) return rc;
    const size_t Nh = (size_t)m->nxh * m->ny;
    int32_t* dst[2] = {even, odd};
    for (int j = 0; j < m->n_multi; ++j)
        for (int c = 0; c < 2; ++c) {
            COUNT_LAUNCH();
            sixclock_export_half_kernel<<<(unsigned)((Nh + 255) / 256), 256, 0, m->stream>>>(m->c[c] + j * m->rep_bytes, m->nxh, (int)m->ny, m->pitch, m->d_stage);
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(dst[c] + (size_t)j * Nh, m->d_stage, Nh * sizeof(int32_t), cudaMemcpyDeviceToHost, m->stream));
            CK(cudaStreamSynchronize(m->stream));
        }
    return B200MC_OK;
}
int b200mc_sixclock_set_dual(void* h, const int32_t* even, const int32_t* odd)
{
    CHECK_S(h);
    if (!even || !odd) ARG_FAIL("null input");
    Six* m = HS(h);
    const size_t Nh = (size_t)m->nxh * m->ny;
    const int32_t* src[2] = {even, odd};
    for (int c = 0; c < 2; ++c)
        for (size_t i = 0; i < Nh * (size_t)m->n_multi; ++i)
            if (src[c][i] < 0 || src[c][i] >= m->q) ARG_FAIL("sixclock: state %d outside 0..%d", src[c][i], m->q - 1);
    int rc = stage(m);
    if (rc) return rc;
    m->obs_valid = false;
    for (int j = 0; j < m->n_multi; ++j)
        for (int c = 0; c < 2; ++c) {
            CK(cudaMemcpyAsync(m->d_stage, src[c] + (size_t)j * Nh, Nh * sizeof(int32_t), cudaMemcpyHostToDevice, m->stream));
            COUNT_LAUNCH();
            sixclock_import_half_kernel<<<(unsigned)((Nh + 255) / 256), 256, 0, m->stream>>>(m->c[c] + j * m->rep_bytes, m->nxh, (int)m->ny, m->pitch, m->d_stage);
            CK(cudaGetLastError());
            CK(cudaStreamSynchronize(m->stream));
        }
    return B200MC_OK;
}
int b200mc_sixclock_get_states_to_prob(void* h, double* out)
{
    CHECK_S(h);
    for (size_t i = 0; i < HS(h)->prob.size(); ++i) out[i] = HS(h)->prob[i];
    return B200MC_OK;
}
int b200mc_sixclock_get_energy_table(void* h, double* out)
{
    CHECK_S(h);
    for (size_t i = 0; i < HS(h)->e3.size(); ++i) out[i] = HS(h)->e3[i];
    return B200MC_OK;
}
int64_t b200mc_sixclock_nx(void* h) { return h ? HS(h)->nx : -1; }
int64_t b200mc_sixclock_ny(void* h) { return h ? HS(h)->ny : -1; }
int64_t b200mc_sixclock_nall(void* h) { return h ? HS(h)->nx * HS(h)->ny : -1; }
int32_t b200mc_sixclock_mstate(void* h) { return h ? HS(h)->q : -1; }
int b200mc_sixclock_set_sample_offset(void* h, int32_t first_sample)
{
    CHECK_S(h);
    if (first_sample < 0) ARG_FAIL("first_sample must be >= 0");
    HS(h)->sample0 = first_sample;
    return B200MC_OK;
}
int32_t b200mc_sixclock_n_multi(void* h) { return h ? HS(h)->n_multi : -1; }
double b200mc_sixclock_kbt(void* h) { return h ? 1 / HS(h)->beta : 0.0; }
double b200mc_sixclock_beta(void* h) { return h ? HS(h)->beta : 0.0; }
int b200mc_sixclock_set_timing(void* h, int32_t on) { CHECK_S(h); HS(h)->timing = on != 0; HS(h)->ev_used = 0; return B200MC_OK; }
int b200mc_sixclock_get_timing(void* h, int64_t* launches, double* total_ms)
{
    CHECK_S(h);
    Six* m = HS(h);
    CK(cudaStreamSynchronize(m->stream));
    double tot = 0.0;
    for (size_t i = 0; i + 1 < m->ev_used; i += 2) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, m->evs[i], m->evs[i + 1]));
        tot += ms;
    }
    if (launches) *launches = (int64_t)(m->ev_used / 2);
    if (total_ms) *total_ms = tot;
    return B200MC_OK;
}
int b200mc_sixclock_sync(void* h) { CHECK_S(h); CK(cudaStreamSynchronize(HS(h)->stream)); return B200MC_OK; }

}  // extern "C"
