// Ising 2D / 3D with TRUE PERIODIC boundaries (torus): host-side handle, kernels and C ABI (b200mc_ising_torus_*).
//
// Not a reference module: the reference's Ising types are helical and only valid for odd nx
// (src/ising3d_gpu_m.f90:60-62,196), so "L = 1024^3" (BASELINE.json north_star, BASELINE.md C2 "1024^3
// periodic", C1 "1024^2 periodic", C5 "65536^2 periodic") is not a shape its kernels can run.  This module
// is that input: the reference's update rule, tables, value conventions and observables
// (update_sub src/ising3d_gpu_m.f90:189-206, src/ising2d_gpu_m.f90:148-162; calc_*_sum :239-276 / :198-228) on the
// torus, colour = (x + y + z) & 1, colour 0 first.  CPU restatement: oracle/oracle.c orc_isingp_*.
//
// Layout.  One byte per site, colours split, each colour row-major [z][y][xi] with xi = x >> 1 fastest: a 128-bit
// vector = 16 consecutive xi of one row, R = nx / 32 vectors per row, no halo (the wrap is index arithmetic).  The other
// colour's neighbours of vector (z, y, j) are the vectors at the SAME j in rows y +- 1 and planes z +- 1, the vector at
// the same place (x - 1 or x + 1, depending on the row's parity) and that vector moved by ONE BYTE (the other x
// neighbour): four funnel shifts plus the edge byte of the adjacent vector, fetched with one SHFL from the neighbouring
// lane.  Every load is a 128-byte-aligned LDG.128.
//
// Strip kernel (rows of at least 32 vectors, nx % 1024 == 0).  A warp owns 32 adjacent vectors and walks ROWS
// consecutive rows of one plane; the other colour's row y + 1 it loads for row y is the centre row of y + 1 and the
// y - 1 row of y + 2, so it stays in registers (rolling three-row window).  Per vector: own + ONE row of the plane + two
// z rows = 4 loads and a store, against 7 + 1 in the helical fold, whose x / y / z offsets are whole-vector offsets of
// a folded ring and cannot be shared between steps.  The helical pass is bound by L1TEX wavefronts (35 per warp-vector,
// profiles/r01_ising3d_ncu_full.md: 73 %); here it is 21 + one SHFL.  The arithmetic after the loads (Philox block,
// byte-parallel sums, PRMT threshold lookup, deferred ties, fused E / M) is ising_core / ising_drain of
// ising_kernels.cuh, unchanged: RNG contract = the ring models' with (position, lane) = (vector index, byte).
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <new>
#include <vector>
#include "../../include/b200mc.h"
#include "ising_kernels.cuh"
#include "ising_tables.cuh"
#include "ring.cuh"

namespace {

struct TorusArgs {
    RingPassArgs a;      // own / oth (no halo: H = 0), draw, colour, seed, acc, ticket + the self-cleaning pointers
    int R, ny, nz;       // vectors per row, rows per plane, planes (1 in 2D)
    int strips, yblocks; // strip kernel: R / 32, ny / ROWS
    int ntickets;        // strips * yblocks * nz
    int64_t nvec;        // R * ny * nz
    int64_t nx;          // sites per row (both colours)
    // slab mode (planes split over the ranks): zwrap = 0 -> the planes below / above the owned ones are ghost planes in the same
    // array (no wrap); z0 = global index of local plane 0 (row parity; the Philox counter takes a.p0 = z0 * R * ny).
    // A launch covers the planes zi0, zi0 + zstride, ...: the boundary launch of a slab pass is planes {0, nz - 1}.
    int zwrap, z0, zi0, zstride;
};

// The x neighbour that is not at the same xi: row parity par = (y + z + colour) & 1 of the row being updated.
//   par = 1: xi + 1 -> byte i <- byte i + 1, last byte <- byte 0 of the next vector      (edge = next.x)
//   par = 0: xi - 1 -> byte i <- byte i - 1, first byte <- byte 15 of the previous vector (edge = prev.w)
__device__ __forceinline__ uint4 torus_shift(uint4 B, int par, uint32_t edge)
{
    const uint32_t w0 = par ? B.x : edge, w1 = par ? B.y : B.x, w2 = par ? B.z : B.y, w3 = par ? B.w : B.z, w4 = par ? edge : B.w;
    const uint32_t sh = par ? 8u : 24u;
    return make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
}

#ifndef TORUS_ZM_EVICT
#define TORUS_ZM_EVICT 0
#endif
// the z - 1 row is the last use of that row in the pass: optionally marked evict-first in L2
__device__ __forceinline__ uint4 ld_other_last(const uint4* p, uint64_t pol)
{
#if TORUS_ZM_EVICT
    uint4 r;
    asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
    return r;
#else
    return ld_other(p);
#endif
}

__device__ __forceinline__ uint32_t ld_word_nc(const uint32_t* p)
{
    uint32_t r;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

// generic gather of the NNB neighbour vectors of vector v (any R): nb = {same xi, shifted, y-, y+ (, z+, z-)}
template <int NNB>
__device__ __forceinline__ void torus_gather(const uint4* oth, int R, int ny, int nz, int64_t v, uint32_t colour, uint4 (&nb)[NNB],
                                             int zwrap = 1, int z0 = 0)
{
    const int64_t row = v / R;
    const int col = (int)(v - row * R);
    const int z = (int)(row / ny), y = (int)(row - (int64_t)z * ny);
    const int par = (y + z + z0 + (int)colour) & 1;
    const uint4 B = ld_other(oth + v);
    const int ncol = par ? (col + 1 == R ? 0 : col + 1) : (col == 0 ? R - 1 : col - 1);
    const uint32_t edge = ld_word_nc(reinterpret_cast<const uint32_t*>(oth + row * R + ncol) + (par ? 0 : 3));
    nb[0] = B;
    nb[1] = torus_shift(B, par, edge);
    const int ym = y == 0 ? ny - 1 : y - 1, yp = y + 1 == ny ? 0 : y + 1;
    const int64_t zb = (int64_t)z * ny;
    nb[2] = ld_other(oth + (zb + ym) * R + col);
    nb[3] = ld_other(oth + (zb + yp) * R + col);
    if (NNB == 6) {
        const int zp = (zwrap && z + 1 == nz) ? 0 : z + 1, zm = (zwrap && z == 0) ? nz - 1 : z - 1;   // (slab mode: ghost planes -1 and nz)
        nb[4] = ld_other(oth + ((int64_t)zp * ny + y) * R + col);
        nb[5] = ld_other(oth + ((int64_t)zm * ny + y) * R + col);
    }
}

// end of a measuring pass: the lane's sums -> warp -> acc; the last warp of the launch stores the totals in pinned host memory
template <int NNB>
__device__ __forceinline__ void torus_finish_sums(const RingPassArgs& a, uint32_t accX, uint32_t accM, int corrX, int corrM)
{
    const int lane = threadIdx.x & 31;
    long long x = (long long)NNB * accM - 2ll * accX + corrX, mm = (long long)accM + corrM;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        x += __shfl_down_sync(0xffffffffu, x, o);
        mm += __shfl_down_sync(0xffffffffu, mm, o);
    }
    if (lane == 0) {
        if (x) atomicAdd(a.acc, (unsigned long long)x);
        if (mm) atomicAdd(a.acc + 1, (unsigned long long)mm);
        if (a.host_out) {
            __threadfence();
            const unsigned int total = gridDim.x * (blockDim.x >> 5);
            if (atomicAdd(a.done_warps, 1u) == total - 1u) {
                __threadfence();
                const unsigned long long t0 = atomicAdd(a.acc, 0ull), t1 = atomicAdd(a.acc + 1, 0ull);
                *a.done_warps = 0u;
                a.host_out[0] = t0;
                a.host_out[1] = t1;
                __threadfence_system();
            }
        }
    }
}

#ifndef TORUS_MINB
#define TORUS_MINB(NNB) ((NNB) == 6 ? 3 : 4)   // resident blocks per SM the registers are allocated for (80 / 64)
#endif
#ifndef TORUS_NB3
#define TORUS_NB3 2         // rows per batch in 3D (4 loads each)
#endif
#ifndef TORUS_NB2
#define TORUS_NB2 2         // rows per batch in 2D (2 loads each); measured: 2 rows x 4 blocks 1621, 2 x 3 1564, 4 x 3 1480 flips/ns at 65536^2
#endif
#ifndef TORUS_PF
#define TORUS_PF 3          // L2 prefetch of the next ticket: bit 0 own rows, bit 1 rows of the leading z plane
#endif
#ifndef TORUS_PFD
#define TORUS_PFD 1024      // which ticket is prefetched: D > 0 = the ticket D after the current one; 0 = this warp's next ticket, which lies
                            // about one ticket per resident warp (3552) ahead of the frontier: 29 MB of prefetched lines waiting in L2 next to
                            // the 15 MB window of the other colour and the dirty lines of the own one, and 14 % more DRAM reads than the
                            // algorithmic figure.  Measured at 1024^3: 0 -> 1705, 256 -> 1740, 512 -> 1748, 768..1536 -> 1757 flips/ns
#endif
#ifndef TORUS_RC2D
#define TORUS_RC2D 1        // 2D: kernels with the row pitch of 16384^2 / 65536^2 compiled in
#endif
#ifndef TORUS_SPLIT_DEFAULT
#define TORUS_SPLIT_DEFAULT 0
#endif
#ifndef TORUS_ROWS
#define TORUS_ROWS 8        // rows per ticket when ny allows
#endif

// ---- strip kernel: R % 32 == 0, ny % ROWS == 0, ROWS even ---------------------------------------------------------
// The rows of one ticket.  PAR0 = parity of the first row (constant per plane and colour: y0 is even), RC = R when it is
// a compile-time constant (32: nx = 1024, one strip per row -- every row offset is an immediate and no strip has an
// edge inside a row), else 0.  The first version computed every address from (z, y, col) per row, selected the shift
// direction at run time and ran 233 instructions per warp-vector against the helical pass's 180: 373 us per pass with
// L1TEX at 43 % (profiles/r02z_torus_ncu.md) -- the pass time follows the instruction count.
// Rows are taken in batches of NB = 2: ALL loads of a batch are issued before the first SHFL --
// the shuffle needs its row's data, and with one row's loads behind it each row paid a full memory round trip.  The
// batch loop is not unrolled (two parity variants of a 2-row body = 700 instructions; fully unrolled over 8 rows the
// kernel was 2900 instructions, 46 KB, and 20 % slower).
template <int NNB, int METHOD, bool MEASURE, int ROWS, int PAR0, int RC>
__device__ __forceinline__ void torus_rows(const TorusArgs& t, const IsingTab& tab, uint4* pown, const uint4* pz, const uint4* pzp,
                                           const uint4* pzm, const uint4* prow_m, const uint4* prow_p, int vbase, int col, int lane,
                                           uint32_t cz, uint32_t cw, uint64_t pol, uint32_t qaddr, uint32_t cntaddr, uint32_t& bX,
                                           uint32_t& bM, uint32_t& accX, uint32_t& accM, int& corrX, int& corrM)
{
    const RingPassArgs& a = t.a;
    const int R = RC ? RC : t.R;
    constexpr int DN = MEASURE ? NNB : 0;
    constexpr int NB = ROWS < 4 ? 2 : (NNB == 6 ? TORUS_NB3 : TORUS_NB2);
    static_assert(ROWS % NB == 0 && NB % 2 == 0, "whole batches of an even number of rows");
    uint4 A = ld_other(prow_m);   // row y0 - 1 (wrapped)
    uint4 B = ld_other(pz);       // row y0
    int v = vbase;
#pragma unroll 1
    for (int b = 0; b < ROWS / NB; ++b) {
        uint4 C[NB], Zp[NB], Zm[NB], O[NB];
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            // row y + 1 of row i of the batch: the next row's centre (the row after the ticket's last one may wrap)
            C[i] = ld_other((i == NB - 1 && b == ROWS / NB - 1) ? prow_p : pz + (i + 1) * R);
            if (NNB == 6) { Zp[i] = ld_other(pzp + i * R); Zm[i] = ld_other_last(pzm + i * R, pol); }
            O[i] = ld_own(pown + i * R, pol);
        }
        // strips narrower than the row: the lane at the end of the strip takes the edge byte of the adjacent vector from memory
        // (requested here, with the batch's loads: behind the SHFL it was a dependent load in the middle of every row's arithmetic)
        uint32_t E[NB];
        if (RC != 32 && t.strips > 1) {
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                const int par = PAR0 ^ (i & 1);
                E[i] = 0u;
                if (lane == (par ? 31 : 0)) {
                    const int ncol = par ? (col + 1 == R ? 0 : col + 1) : (col == 0 ? R - 1 : col - 1);
                    E[i] = ld_word_nc(reinterpret_cast<const uint32_t*>(pz + i * R + (ncol - col)) + (par ? 0 : 3));
                }
            }
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int par = PAR0 ^ (i & 1);
            const uint4 ctr = i == 0 ? B : C[i - 1];
            // edge byte of the adjacent vector: from the neighbouring lane, or the word loaded above
            uint32_t edge = __shfl_sync(0xffffffffu, par ? ctr.x : ctr.w, (lane + (par ? 1 : 31)) & 31);
            if (RC != 32 && t.strips > 1 && lane == (par ? 31 : 0)) edge = E[i];
            uint4 nb[NNB];
            nb[0] = ctr;
            nb[1] = par ? make_uint4(__funnelshift_r(ctr.x, ctr.y, 8), __funnelshift_r(ctr.y, ctr.z, 8), __funnelshift_r(ctr.z, ctr.w, 8), __funnelshift_r(ctr.w, edge, 8))
                        : make_uint4(__funnelshift_l(edge, ctr.x, 8), __funnelshift_l(ctr.x, ctr.y, 8), __funnelshift_l(ctr.y, ctr.z, 8), __funnelshift_l(ctr.z, ctr.w, 8));
            nb[2] = i == 0 ? A : (i == 1 ? B : C[i - 2]);
            nb[3] = C[i];
            if (NNB == 6) { nb[4] = Zp[i]; nb[5] = Zm[i]; }
            ising_core<NNB, METHOD, false, MEASURE>(v + i * R, pown + i * R, O[i], nb, (uint32_t)(v + i * R) + (uint32_t)a.p0, cz, cw, a, tab, pol, qaddr, cntaddr, false, bX, bM, 0u);
            if ((i & 1) && i != NB - 1) {
                __syncwarp();
                if (lds32(cntaddr) > TQ_CAP - 64) {
                    const int2 d = ising_drain<METHOD, false, DN>(qaddr, cntaddr, a.own, a, tab, 0u);
                    corrX += d.x; corrM += d.y;
                }
            }
        }
        A = C[NB - 2];
        B = C[NB - 1];
        pz += NB * R; pzp += NB * R; pzm += NB * R; pown += NB * R; v += NB * R;
        __syncwarp();
        if (lds32(cntaddr) > TQ_CAP - 64) {
            const int2 d = ising_drain<METHOD, false, DN>(qaddr, cntaddr, a.own, a, tab, 0u);
            corrX += d.x; corrM += d.y;
        }
        if (MEASURE) ising_fold_sums(bX, bM, accX, accM);   // at most 4 vectors per fold
    }
}

template <int NNB, int METHOD, bool MEASURE, int ROWS, int RC>
__global__ void __launch_bounds__(256, TORUS_MINB(NNB))
torus_strip_kernel(const __grid_constant__ TorusArgs t, const __grid_constant__ IsingTab tab)
{
    static_assert(ROWS % 2 == 0, "the row parity alternates: two rows per queue look");
    __shared__ uint4 tq[8][TQ_CAP][2];
    __shared__ uint32_t tq_cnt[8];
    const RingPassArgs& a = t.a;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t qaddr = (uint32_t)__cvta_generic_to_shared(&tq[warp][0][0]);
    uint32_t cntaddr = (uint32_t)__cvta_generic_to_shared(&tq_cnt[warp]);
    pin32(qaddr);
    pin32(cntaddr);
    if (lane == 0) tq_cnt[warp] = 0;
    __syncwarp();
    if (blockIdx.x == 0) {   // self-cleaning launches (RingPassArgs)
        if (a.ticket_reset && threadIdx.x < TK_NCNT) a.ticket_reset[threadIdx.x * 64] = 0u;
        if (!MEASURE && a.acc_reset && threadIdx.x < 2) a.acc_reset[threadIdx.x] = 0ull;
    }
    const uint64_t pol = l2_policy_evict_first();
    uint4* own = a.own;
    const uint4* oth = a.oth;
    const int R = RC ? RC : t.R, ny = t.ny, nz = t.nz;
    const int plane = R * ny;            // vectors per z plane (nvec < 2^31)
    const uint32_t cz = (uint32_t)a.draw;
    const uint32_t cw = (uint32_t)((a.draw >> 32) & 0xFFFFu) | (a.colour << 16);
    const int gwarp = blockIdx.x * (blockDim.x >> 5) + warp;
    unsigned int* tk = a.ticket + (gwarp % TK_NCNT) * 64;
    const int tk_base = gwarp % TK_NCNT;
    uint32_t accX = 0, accM = 0, bX = 0, bM = 0;
    int corrX = 0, corrM = 0;
    constexpr int DN = MEASURE ? NNB : 0;
    int cur, nxt = 0, nx2 = 0;
    if (lane == 0) {
        nxt = (int)atomicAdd(tk, 1u) * TK_NCNT + tk_base;
        nx2 = (int)atomicAdd(tk, 1u) * TK_NCNT + tk_base;
    }
    cur = __shfl_sync(0xffffffffu, nxt, 0);
    nxt = __shfl_sync(0xffffffffu, nx2, 0);
    const int strips = RC ? RC / 32 : t.strips;
    // first vector of ticket q: (plane, block of ROWS rows, strip of 32 vectors), strips fastest, planes slowest -- all resident
    // warps work in one narrow window of the lattice, and the z planes of the other colour stay in L2 between their uses
    auto ticket_vec = [&](int q, int& zi, int& y0) -> int {
        const int xs = q % strips, rest = q / strips;
        const int yb = rest % t.yblocks;
        zi = t.zi0 + (rest / t.yblocks) * t.zstride;
        y0 = yb * ROWS;
        return zi * plane + y0 * R + xs * 32;
    };
    while (cur < t.ntickets) {
        if (lane == 0) nx2 = (int)atomicAdd(tk, 1u) * TK_NCNT + tk_base;   // two tickets ahead
        const int pft = TORUS_PFD ? cur + TORUS_PFD : nxt;
        if (TORUS_PF && pft < t.ntickets) {
            // the streams of the next ticket that come from DRAM -- its own rows and the rows of the leading z plane -- are
            // requested into L2 now: ROWS rows x 512 bytes = 4 lines per row, lane -> (row, line)
            int nzi, ny0;
            const int nv0 = ticket_vec(pft, nzi, ny0);
#pragma unroll
            for (int l = lane; l < 4 * ROWS; l += 32) {
                const int nv = nv0 + (l >> 2) * R;
                if (TORUS_PF & 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(own + nv) + (l & 3) * 128));
                if (NNB == 6 && (TORUS_PF & 2)) {
                    const int nvz = nv + ((t.zwrap && nzi + 1 == nz) ? -nzi : 1) * plane;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(oth + nvz) + (l & 3) * 128));
                }
            }
        }
        int zi, y0;
        const int v0 = ticket_vec(cur, zi, y0) + lane;     // this lane's vector in the ticket's first row
        const int col = (RC == 32) ? lane : (v0 - zi * plane - y0 * R);
        const uint4* pz = oth + v0;
        const uint4* pzp = oth + (v0 + ((t.zwrap && zi + 1 == nz) ? -zi : 1) * plane);
        const uint4* pzm = oth + (v0 + ((t.zwrap && zi == 0) ? nz - 1 : -1) * plane);
        const uint4* prow_m = pz + (y0 == 0 ? ny - 1 : -1) * R;
        const uint4* prow_p = pz + (y0 + ROWS == ny ? ROWS - ny : ROWS) * R;
        if ((zi + t.z0 + (int)a.colour) & 1)
            torus_rows<NNB, METHOD, MEASURE, ROWS, 1, RC>(t, tab, own + v0, pz, pzp, pzm, prow_m, prow_p, v0, col, lane, cz, cw, pol, qaddr, cntaddr,
                                                          bX, bM, accX, accM, corrX, corrM);
        else
            torus_rows<NNB, METHOD, MEASURE, ROWS, 0, RC>(t, tab, own + v0, pz, pzp, pzm, prow_m, prow_p, v0, col, lane, cz, cw, pol, qaddr, cntaddr,
                                                          bX, bM, accX, accM, corrX, corrM);
        cur = nxt;
        nxt = __shfl_sync(0xffffffffu, nx2, 0);
    }
    const int2 d = ising_drain<METHOD, false, DN>(qaddr, cntaddr, own, a, tab, 0u);
    if (MEASURE) torus_finish_sums<NNB>(a, accX, accM, corrX + d.x, corrM + d.y);
}

// ---- generic kernel: any nx % 32 == 0; one vector per lane and step, every neighbour loaded ----------------------
template <int NNB, int METHOD, bool MEASURE>
__global__ void __launch_bounds__(256, TORUS_MINB(NNB))
torus_pass_kernel(const __grid_constant__ TorusArgs t, const __grid_constant__ IsingTab tab)
{
    __shared__ uint4 tq[8][TQ_CAP][2];
    __shared__ uint32_t tq_cnt[8];
    const RingPassArgs& a = t.a;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t qaddr = (uint32_t)__cvta_generic_to_shared(&tq[warp][0][0]);
    uint32_t cntaddr = (uint32_t)__cvta_generic_to_shared(&tq_cnt[warp]);
    if (lane == 0) tq_cnt[warp] = 0;
    __syncwarp();
    const uint64_t pol = l2_policy_evict_first();
    const uint32_t cz = (uint32_t)a.draw;
    const uint32_t cw = (uint32_t)((a.draw >> 32) & 0xFFFFu) | (a.colour << 16);
    uint32_t accX = 0, accM = 0, bX = 0, bM = 0;
    int corrX = 0, corrM = 0;
    constexpr int DN = MEASURE ? NNB : 0;
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t base = ((int64_t)blockIdx.x * (blockDim.x >> 5) + warp) * 32; base < t.nvec; base += nwarps * 32) {
        const int64_t v = base + lane;
        if (v < t.nvec) {
            uint4 nb[NNB];
            torus_gather<NNB>(a.oth, t.R, t.ny, t.nz, v, a.colour, nb, t.zwrap, t.z0);
            const uint4 o = ld_own(a.own + v, pol);
            ising_core<NNB, METHOD, false, MEASURE>((int)v, a.own + v, o, nb, (uint32_t)v + (uint32_t)a.p0, cz, cw, a, tab, pol, qaddr, cntaddr, false, bX, bM, 0u);
            if (MEASURE) ising_fold_sums(bX, bM, accX, accM);
        }
        __syncwarp();
        if (lds32(cntaddr) > TQ_CAP - 64) {
            const int2 d = ising_drain<METHOD, false, DN>(qaddr, cntaddr, a.own, a, tab, 0u);
            corrX += d.x; corrM += d.y;
        }
    }
    const int2 d = ising_drain<METHOD, false, DN>(qaddr, cntaddr, a.own, a, tab, 0u);
    if (MEASURE) torus_finish_sums<NNB>(a, accX, accM, corrX + d.x, corrM + d.y);
}

// ---- reference-stream pass: uniforms from a caller array indexed by site, real64 compare (parity tests with arbitrary uniforms) ----
template <int NNB, int METHOD>
__global__ void __launch_bounds__(256)
torus_pass_randoms_kernel(const __grid_constant__ TorusArgs t, const __grid_constant__ IsingTabF64 tab, const double* __restrict__ randoms)
{
    const RingPassArgs& a = t.a;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < t.nvec; v += stride) {
        uint4 nb[NNB];
        torus_gather<NNB>(a.oth, t.R, t.ny, t.nz, v, a.colour, nb);
        uint4 o = a.own[v];
        uint4 S = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int j = 0; j < NNB; ++j) { S.x += nb[j].x; S.y += nb[j].y; S.z += nb[j].z; S.w += nb[j].w; }
        uint32_t ow[4] = {o.x, o.y, o.z, o.w};
        const uint32_t sw[4] = {S.x, S.y, S.z, S.w};
        const int64_t row = v / t.R;
        const int col = (int)(v - row * t.R);
        const int z = (int)(row / t.ny), y = (int)(row - (int64_t)z * t.ny);
        const int par = (y + z + (int)a.colour) & 1;
#pragma unroll
        for (int b = 0; b < 16; ++b) {
            const int64_t x = 2 * ((int64_t)col * 16 + b) + par;
            const int64_t i = x + t.nx * row;
            const uint32_t s = (ow[b >> 2] >> (8 * (b & 3))) & 0xFFu;
            const uint32_t Sb = (sw[b >> 2] >> (8 * (b & 3))) & 0xFFu;
            const double w = tab.w[(METHOD == METHOD_METROPOLIS ? s * 8 : 0) + Sb];
            if (randoms[i] > w) {
                if (METHOD == METHOD_HEATBATH) ow[b >> 2] &= ~(0xFFu << (8 * (b & 3)));
                continue;
            }
            if (METHOD == METHOD_METROPOLIS) ow[b >> 2] ^= 1u << (8 * (b & 3));
            else ow[b >> 2] = (ow[b >> 2] & ~(0xFFu << (8 * (b & 3)))) | (1u << (8 * (b & 3)));
        }
        a.own[v] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
}

// ---- E + M in one pass over the colour-1 vectors (every bond has exactly one colour-1 end): acc[0] += X, acc[1] += sum(s) ----
template <int NNB>
__global__ void __launch_bounds__(256)
torus_measure_kernel(const uint4* __restrict__ c0, const uint4* __restrict__ c1, int R, int ny, int nz, int64_t nvec, unsigned long long* acc,
                     int zwrap, int z0)
{
    long long part[2] = {0, 0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        uint4 nb[NNB];
        torus_gather<NNB>(c0, R, ny, nz, v, 1u, nb, zwrap, z0);
        const uint4 o = c1[v];
        uint4 X = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int j = 0; j < NNB; ++j) { X.x += nb[j].x ^ o.x; X.y += nb[j].y ^ o.y; X.z += nb[j].z ^ o.z; X.w += nb[j].w ^ o.w; }
        part[0] += (((X.x + X.y) + (X.z + X.w)) * 0x01010101u) >> 24;   // bytes <= 24, total <= 96
        part[1] += ((((o.x + nb[0].x) + (o.y + nb[0].y)) + ((o.z + nb[0].z) + (o.w + nb[0].w))) * 0x01010101u) >> 24;
    }
    block_atomic_add<2>(acc, part);
}

// ---- host int32 arrays s[x + nx (y + ny z)] <-> the two colour arrays ----
__global__ void torus_export_kernel(const uint8_t* __restrict__ c0, const uint8_t* __restrict__ c1, int64_t nx, int64_t ny, int64_t n, int32_t* out, int pm1, int64_t z0)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t x = i % nx, row = i / nx, y = row % ny, z = row / ny;
    const uint8_t s = (((x + y + z + z0) & 1) ? c1 : c0)[row * (nx / 2) + (x >> 1)];
    out[i] = pm1 ? (s ? 1 : -1) : (int32_t)s;
}
__global__ void torus_import_kernel(uint8_t* c0, uint8_t* c1, int64_t nx, int64_t ny, int64_t n, const int32_t* __restrict__ in, int pm1, int64_t z0)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t x = i % nx, row = i / nx, y = row % ny, z = row / ny;
    (((x + y + z + z0) & 1) ? c1 : c0)[row * (nx / 2) + (x >> 1)] = pm1 ? (uint8_t)(in[i] > 0) : (uint8_t)in[i];
}

// ---- slab mode, direct transport: ghost planes over peer memory ------------------------------------------------------
// After a colour pass every rank stores its first owned plane of the colour into the ghost plane above rank - 1's slab and
// its last owned plane into the ghost plane below rank + 1's (arrays mapped with CUDA IPC, stores over NVLink), then the
// last block to finish publishes a sequence number in both neighbours' flag words (system-scope release) and waits for
// theirs: when the kernel ends, the ghost planes the next pass reads have landed.  No NCCL kernel (which cannot run beside
// or right behind the pass without a launch of its own and a rendezvous), no host synchronisation.  Only the boundary
// planes of a neighbour's pass read its ghost planes, and that pass finished before this rank could start the pass whose
// results it pushes here (it waited for that neighbour's previous flag): no write-after-read hazard.
__global__ void __launch_bounds__(256)
torus_push_kernel(const uint4* __restrict__ first, const uint4* __restrict__ last, uint4* __restrict__ prev_high, uint4* __restrict__ next_low,
                  int64_t nplane, unsigned int* done, unsigned int* sig_prev, unsigned int* sig_next, const unsigned int* wait_prev,
                  const unsigned int* wait_next, unsigned int seq)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * nplane; i += stride) {
        if (i < nplane) prev_high[i] = first[i];
        else next_low[i - nplane] = last[i - nplane];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(done, 1u) == gridDim.x - 1u) {
        *done = 0u;
        __threadfence_system();
        st_release_sys(sig_prev, seq);
        st_release_sys(sig_next, seq);
        // (bounded: a neighbour that never arrives -- a rank that died -- aborts this kernel after a few seconds instead of hanging the GPU)
        unsigned int spins = 0;
        while ((int)(ld_acquire_sys(wait_prev) - seq) < 0) { __nanosleep(200); if (++spins > (1u << 25)) __trap(); }
        while ((int)(ld_acquire_sys(wait_next) - seq) < 0) { __nanosleep(200); if (++spins > (1u << 25)) __trap(); }
    }
}

#define TORUS_FLAG_WORDS 512   // own flag buffer: word 32 (2 colour + side) = pushes received from prev (side 0) / next (side 1); word 256 = finished blocks
#define TORUS_IPC_BYTES 192

#define TORUS_MAGIC 0x544F5255

struct Torus {
    int magic;
    int ndim;
    int64_t nx, ny, nz;     // nz = 1 in 2D
    int R;
    int64_t nvec, N;
    uint4* vec[2];          // first OWNED plane of each colour
    uint4* alloc[2];        // the allocations (slab mode: one ghost plane in front of and one behind the owned planes)
    // slab mode: planes [z0, z0 + nz) of a lattice of nz_glob planes, one process per GPU; ghost planes through ncclSend/Recv
    int rank, nranks;
    int64_t z0, nz_glob, N_glob;
    void* comm;
    cudaStream_t comm_stream;
    cudaEvent_t ev_boundary, ev_halo;
    // direct transport (CUDA IPC): the neighbours' colour arrays and flag words
    bool p2p;
    unsigned int* flags;
    uint4* peer_alloc[2][2];       // [prev / next][colour]
    unsigned int* peer_flags[2];
    void* peer_maps[6];
    int n_peer_maps;
    unsigned int push_seq[2];
    bool split;             // slab pass = boundary planes, then exchange beside the interior planes (else: one launch, then the exchange)
    cudaStream_t stream;
    double beta;
    uint32_t seed;
    uint64_t draw;
    int method;
    IsingHostTables tabs;
    unsigned long long* d_acc;
    unsigned long long* h_acc;
    unsigned int* d_ticket;
    int32_t* d_io;          // staging for get_spins / set_spins (allocated on first use)
    double* d_randoms;
    int grid, rows;         // resident grid; rows per ticket of the strip kernel (0: generic kernel)
    bool obs_valid;
    int64_t obs_e, obs_m;
    bool want_fused, fused_pending, h_acc_pending;
    bool timing;
    std::vector<cudaEvent_t> evs;
    size_t ev_used;
    bool alive;
};

int torus_destroy(Torus* m)
{
    if (!m) return B200MC_OK;
    cudaFree(m->alloc[0]); cudaFree(m->alloc[1]); cudaFree(m->d_acc);
    for (int i = 0; i < m->n_peer_maps; ++i) cudaIpcCloseMemHandle(m->peer_maps[i]);
    cudaFree(m->flags);
    if (m->comm) dist_comm_destroy(m->comm);
    if (m->comm_stream) cudaStreamDestroy(m->comm_stream);
    if (m->ev_boundary) cudaEventDestroy(m->ev_boundary);
    if (m->ev_halo) cudaEventDestroy(m->ev_halo); cudaFree(m->d_ticket); cudaFree(m->d_io); cudaFree(m->d_randoms);
    cudaFreeHost(m->h_acc);
    for (cudaEvent_t e : m->evs) cudaEventDestroy(e);
    m->alive = false;
    m->magic = 0;
    delete m;
    return B200MC_OK;
}

int torus_tables(Torus* m) { return ising_build_host_tables(m->ndim, m->method, m->beta, m->seed, &m->tabs); }

int torus_halo(struct Torus* m, int colour, cudaStream_t st);

int torus_create(void** out, int ndim, int64_t nx, int64_t ny, int64_t nz, double kbt, int32_t iseed,
                 int rank = 0, int nranks = 1, const char* nccl_id = nullptr)
{
    if (!out) ARG_FAIL("null handle pointer");
    *out = nullptr;
    if (ndim != 2 && ndim != 3) ARG_FAIL("ndim must be 2 or 3");
    if (ndim == 2) nz = 1;
    if (nx < 32 || nx % 32 != 0) ARG_FAIL("periodic Ising: nx must be a positive multiple of 32 (16 sites of a colour per 128-bit vector), got %lld", (long long)nx);
    if (ny < 2 || ny % 2 != 0) ARG_FAIL("periodic Ising: ny must be even and >= 2 (checkerboard colouring on the torus), got %lld", (long long)ny);
    if (ndim == 3 && (nz < 2 || nz % 2 != 0)) ARG_FAIL("periodic Ising: nz must be even and >= 2, got %lld", (long long)nz);
    if (!(kbt > 0.0)) ARG_FAIL("kbt must be > 0");
    if (nx / 32 * ny * nz >= (int64_t)1 << 31) ARG_FAIL("lattice too large: 2^31 vectors per colour");
    if (nranks < 1 || rank < 0 || rank >= nranks) ARG_FAIL("invalid rank %d of %d", rank, nranks);
    const int64_t nz_glob = nz;
    if (nranks > 1) {
        if (ndim != 3) ARG_FAIL("periodic Ising slabs: 3D only (planes along z)");
        if (nx % 1024 != 0) ARG_FAIL("periodic Ising slabs: nx must be a multiple of 1024 (strip kernel)");
        if (nz % nranks != 0 || nz / nranks < 2) ARG_FAIL("periodic Ising slabs: nz = %lld must be a multiple of the %d ranks with at least 2 planes each", (long long)nz, nranks);
        if (!nccl_id) ARG_FAIL("slab mode needs the NCCL unique id of the job (b200mc_dist_unique_id on rank 0, broadcast by the caller)");
        nz = nz / nranks;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        snprintf(g_b200mc_err, sizeof(g_b200mc_err), "no CUDA device: this library has no CPU fallback");
        return B200MC_ERR_CUDA;
    }
    Torus* m = new (std::nothrow) Torus();
    if (!m) ARG_FAIL("out of host memory");
    m->magic = TORUS_MAGIC; m->ndim = ndim; m->nx = nx; m->ny = ny; m->nz = nz;
    m->R = (int)(nx / 32); m->nvec = (int64_t)m->R * ny * nz; m->N = nx * ny * nz;
    m->vec[0] = m->vec[1] = m->alloc[0] = m->alloc[1] = nullptr;
    m->rank = rank; m->nranks = nranks; m->z0 = (int64_t)rank * nz; m->nz_glob = nz_glob; m->N_glob = nx * ny * nz_glob;
    m->comm = nullptr; m->comm_stream = nullptr; m->ev_boundary = m->ev_halo = nullptr;
    m->p2p = false; m->flags = nullptr; m->n_peer_maps = 0; m->push_seq[0] = m->push_seq[1] = 0;
    { const char* ts = getenv("B200MC_TORUS_SPLIT"); m->split = ts ? atoi(ts) != 0 : TORUS_SPLIT_DEFAULT; }
    m->stream = 0; m->beta = 1 / kbt; m->seed = (uint32_t)iseed; m->draw = 0; m->method = METHOD_METROPOLIS;
    m->d_acc = nullptr; m->h_acc = nullptr; m->d_ticket = nullptr; m->d_io = nullptr; m->d_randoms = nullptr;
    m->obs_valid = false; m->want_fused = false; m->fused_pending = false; m->h_acc_pending = false;
    m->timing = false; m->ev_used = 0; m->alive = true;
    const int64_t plane = (int64_t)m->R * ny, ghost = nranks > 1 ? plane : 0;
    if (nranks > 1) {
        const int rc = dist_comm_init(&m->comm, rank, nranks, nccl_id);
        if (rc) { torus_destroy(m); return rc; }
        // the exchange runs beside the interior launch, whose blocks fill every SM for the whole pass: on a stream of the highest
        // priority its (few) blocks are placed first when both become runnable at the end of the boundary launch
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        if (cudaStreamCreateWithPriority(&m->comm_stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
            cudaEventCreateWithFlags(&m->ev_boundary, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&m->ev_halo, cudaEventDisableTiming) != cudaSuccess) {
            snprintf(g_b200mc_err, sizeof(g_b200mc_err), "cannot create the comm stream / events");
            torus_destroy(m);
            return B200MC_ERR_CUDA;
        }
    }
    if (cudaMalloc(&m->alloc[0], (size_t)(m->nvec + 2 * ghost) * 16) != cudaSuccess || cudaMalloc(&m->alloc[1], (size_t)(m->nvec + 2 * ghost) * 16) != cudaSuccess ||
        cudaMalloc(&m->d_acc, 2 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaHostAlloc(&m->h_acc, 2 * sizeof(unsigned long long), cudaHostAllocDefault) != cudaSuccess ||
        cudaMalloc(&m->d_ticket, (2 * TK_NCNT * 64 + 64) * sizeof(unsigned int)) != cudaSuccess ||
        cudaMemset(m->d_ticket, 0, (2 * TK_NCNT * 64 + 64) * sizeof(unsigned int)) != cudaSuccess) {
        snprintf(g_b200mc_err, sizeof(g_b200mc_err), "cudaMalloc failed (%lld bytes per colour)", (long long)m->nvec * 16);
        cudaGetLastError();
        torus_destroy(m);
        return B200MC_ERR_CUDA;
    }
    m->vec[0] = m->alloc[0] + ghost; m->vec[1] = m->alloc[1] + ghost;
    int dev = 0, sms = 148, occ = 3;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (ndim == 3) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, torus_strip_kernel<6, METHOD_METROPOLIS, true, TORUS_ROWS, 32>, 256, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, torus_strip_kernel<4, METHOD_METROPOLIS, true, TORUS_ROWS, 0>, 256, 0);
    if (occ < 1) occ = 1;
    const int64_t need = (m->nvec + 255) / 256;
    m->grid = (int)(need < (int64_t)sms * occ ? need : (int64_t)sms * occ);
    // strip kernel: rows of whole 32-vector strips; 8 rows per ticket when ny allows, else 2.  Small lattices (fewer tickets than
    // a few per resident warp) and B200MC_TORUS_GENERIC=1 use the generic kernel.
    m->rows = 0;
    const char* tg = getenv("B200MC_TORUS_GENERIC");
    if (m->R % 32 == 0 && (nranks > 1 || !(tg && atoi(tg) != 0))) m->rows = ny % TORUS_ROWS == 0 ? TORUS_ROWS : 2;
    int rc = torus_tables(m);
    if (rc) { torus_destroy(m); return rc; }
    if (cudaMemsetAsync(m->alloc[0], 1, (size_t)(m->nvec + 2 * ghost) * 16, m->stream) != cudaSuccess ||
        cudaMemsetAsync(m->alloc[1], 1, (size_t)(m->nvec + 2 * ghost) * 16, m->stream) != cudaSuccess) {   // set_allup_spin (ghost planes too)
        snprintf(g_b200mc_err, sizeof(g_b200mc_err), "cudaMemset failed");
        torus_destroy(m);
        return B200MC_ERR_CUDA;
    }
    *out = m;
    return B200MC_OK;
}

void torus_args(Torus* m, int colour, TorusArgs& t)
{
    t = TorusArgs();
    RingPassArgs& a = t.a;
    a.own = m->vec[colour];
    a.oth = m->vec[colour ^ 1];
    a.nvec = m->nvec;
    a.seed = m->seed;
    a.colour = (uint32_t)colour;
    a.draw = m->draw;
    a.acc = m->d_acc;
    a.mask_from = 0x7FFFFFFF;
    a.Lfold = 0; a.Nc = INT64_MAX;   // (ising_drain: every byte lane of every vector holds a site)
    a.p0 = m->z0 * (int64_t)m->R * m->ny;   // global index of local vector 0: the Philox counter of a site does not depend on the number of ranks
    t.R = m->R; t.ny = (int)m->ny; t.nz = (int)m->nz; t.nvec = m->nvec; t.nx = m->nx;
    t.zwrap = m->nranks == 1; t.z0 = (int)m->z0; t.zi0 = 0; t.zstride = 1;
}

// slab mode: my first owned plane of `colour` becomes the ghost plane above rank - 1's slab, my last owned plane the ghost
// plane below rank + 1's (ring of ranks: the torus closes between rank P - 1 and rank 0)
int torus_push(Torus* m, int colour, cudaStream_t st)
{
    const int64_t plane = (int64_t)m->R * m->ny;
    const uint4* v = m->vec[colour];
    const unsigned int seq = ++m->push_seq[colour];
    COUNT_LAUNCH();
    torus_push_kernel<<<296, 256, 0, st>>>(v, v + (m->nz - 1) * plane,
                                          m->peer_alloc[0][colour] + plane + m->nvec,   // rank - 1: the ghost plane above its owned planes
                                          m->peer_alloc[1][colour],                     // rank + 1: the ghost plane below
                                          plane, m->flags + 256,
                                          m->peer_flags[0] + 32 * (2 * colour + 1),     // I am rank - 1's "next"
                                          m->peer_flags[1] + 32 * (2 * colour + 0),     // and rank + 1's "prev"
                                          m->flags + 32 * (2 * colour + 0), m->flags + 32 * (2 * colour + 1), seq);
    CK(cudaGetLastError());
    return B200MC_OK;
}

int torus_halo(Torus* m, int colour, cudaStream_t st)
{
    const int64_t plane = (int64_t)m->R * m->ny;
    uint4* v = m->vec[colour];
    return dist_exchange_ring(m->comm, m->rank, m->nranks, v, v + (m->nz - 1) * plane, v - plane, v + m->nz * plane, (size_t)plane * 16, st);
}

template <int NNB>
int torus_launch_pass(Torus* m, int colour, bool fuse, bool fuse_next)
{
    TorusArgs t;
    torus_args(m, colour, t);
    m->obs_valid = false; m->fused_pending = false; m->h_acc_pending = false;
    if (m->timing) {
        while (m->evs.size() < m->ev_used + 2) { cudaEvent_t e; CK(cudaEventCreate(&e)); m->evs.push_back(e); }
        CK(cudaEventRecord(m->evs[m->ev_used], m->stream));
    }
    if (m->nranks > 1) {
        // slab pass: the two boundary planes first (one launch: planes {0, nz - 1}), their exchange on the comm stream while the
        // interior planes are updated; the next launch on the compute stream waits for the ghost planes
        t.strips = m->R / 32; t.yblocks = (int)(m->ny / m->rows);
        const int per_plane = t.strips * t.yblocks;
        if (m->split) {
            if (fuse) CK(cudaMemsetAsync(m->d_acc, 0, 2 * sizeof(unsigned long long), m->stream));
            CK(cudaMemsetAsync(m->d_ticket, 0, 2 * TK_NCNT * 64 * sizeof(unsigned int), m->stream));
        }
        for (int part = m->split ? 0 : 1; part < 2; ++part) {
            TorusArgs u = t;
            u.a.ticket = m->d_ticket + part * TK_NCNT * 64;
            if (!m->split) {   // one launch per pass: self-cleaning counters and sums as on a single GPU
                u.a.ticket = m->d_ticket + colour * TK_NCNT * 64;
                u.a.ticket_reset = m->d_ticket + (colour ^ 1) * TK_NCNT * 64;
                if (colour == 0 && fuse_next) u.a.acc_reset = m->d_acc;
            }
            if (part == 0) { u.zi0 = 0; u.zstride = (int)m->nz - 1; u.ntickets = 2 * per_plane; }
            else if (m->split) { u.zi0 = 1; u.zstride = 1; u.ntickets = ((int)m->nz - 2) * per_plane; }
            else { u.zi0 = 0; u.zstride = 1; u.ntickets = (int)m->nz * per_plane; }   // all planes in one launch, the exchange after it
            if (u.ntickets > 0) {
                const int64_t need = ((int64_t)u.ntickets + 7) / 8;
                const int grid = (int)(need < (int64_t)m->grid ? need : (int64_t)m->grid);
                COUNT_LAUNCH();
#define SPASS(METHOD, MEAS, ROWS, RC) torus_strip_kernel<NNB, METHOD, MEAS, ROWS, RC><<<grid, 256, 0, m->stream>>>(u, m->tabs.tab)
#define SPASS2(METHOD, MEAS) do { if (m->rows == TORUS_ROWS) { if (m->R == 32) SPASS(METHOD, MEAS, TORUS_ROWS, 32); else SPASS(METHOD, MEAS, TORUS_ROWS, 0); } else SPASS(METHOD, MEAS, 2, 0); } while (0)
                if (m->method == METHOD_METROPOLIS) { if (fuse) SPASS2(METHOD_METROPOLIS, true); else SPASS2(METHOD_METROPOLIS, false); }
                else { if (fuse) SPASS2(METHOD_HEATBATH, true); else SPASS2(METHOD_HEATBATH, false); }
#undef SPASS2
#undef SPASS
                CK(cudaGetLastError());
            }
            if (part == 0) {
                CK(cudaEventRecord(m->ev_boundary, m->stream));
                CK(cudaStreamWaitEvent(m->comm_stream, m->ev_boundary, 0));
                const int rc = torus_halo(m, colour, m->comm_stream);
                if (rc) return rc;
                CK(cudaEventRecord(m->ev_halo, m->comm_stream));
            }
        }
        if (m->split) CK(cudaStreamWaitEvent(m->stream, m->ev_halo, 0));
        else { const int rc = m->p2p ? torus_push(m, colour, m->stream) : torus_halo(m, colour, m->stream); if (rc) return rc; }
        if (m->timing) { CK(cudaEventRecord(m->evs[m->ev_used + 1], m->stream)); m->ev_used += 2; }
        return B200MC_OK;
    }
    COUNT_LAUNCH();
    if (m->rows) {
        t.strips = m->R / 32; t.yblocks = (int)(m->ny / m->rows); t.ntickets = t.strips * t.yblocks * (int)m->nz;
        t.a.ticket = m->d_ticket + colour * TK_NCNT * 64;
        t.a.ticket_reset = m->d_ticket + (colour ^ 1) * TK_NCNT * 64;
        if (colour == 0 && fuse_next) t.a.acc_reset = m->d_acc;
        if (fuse) { t.a.host_out = m->h_acc; t.a.done_warps = m->d_ticket + 2 * TK_NCNT * 64; m->h_acc_pending = true; }
        const int64_t need = ((int64_t)t.ntickets + 7) / 8;
        const int grid = (int)(need < (int64_t)m->grid ? need : (int64_t)m->grid);
#define SPASS(METHOD, MEAS, ROWS, RC) torus_strip_kernel<NNB, METHOD, MEAS, ROWS, RC><<<grid, 256, 0, m->stream>>>(t, m->tabs.tab)
        // row pitch as a compile-time constant for the sizes the benchmarks use (3D: nx = 1024; 2D: 16384, 65536), else a kernel argument
#define SPASS2(METHOD, MEAS) do { if (m->rows == TORUS_ROWS) { \
            if (m->R == 32) SPASS(METHOD, MEAS, TORUS_ROWS, 32); \
            else if (TORUS_RC2D && NNB == 4 && m->R == 512) SPASS(METHOD, MEAS, TORUS_ROWS, (NNB == 4 && TORUS_RC2D ? 512 : 0)); \
            else if (TORUS_RC2D && NNB == 4 && m->R == 2048) SPASS(METHOD, MEAS, TORUS_ROWS, (NNB == 4 && TORUS_RC2D ? 2048 : 0)); \
            else SPASS(METHOD, MEAS, TORUS_ROWS, 0); } else SPASS(METHOD, MEAS, 2, 0); } while (0)
        if (m->method == METHOD_METROPOLIS) { if (fuse) SPASS2(METHOD_METROPOLIS, true); else SPASS2(METHOD_METROPOLIS, false); }
        else { if (fuse) SPASS2(METHOD_HEATBATH, true); else SPASS2(METHOD_HEATBATH, false); }
#undef SPASS2
#undef SPASS
    } else {
        if (fuse) CK(cudaMemsetAsync(m->d_acc, 0, 2 * sizeof(unsigned long long), m->stream));
#define GPASS(METHOD, MEAS) torus_pass_kernel<NNB, METHOD, MEAS><<<m->grid, 256, 0, m->stream>>>(t, m->tabs.tab)
        if (m->method == METHOD_METROPOLIS) { if (fuse) GPASS(METHOD_METROPOLIS, true); else GPASS(METHOD_METROPOLIS, false); }
        else { if (fuse) GPASS(METHOD_HEATBATH, true); else GPASS(METHOD_HEATBATH, false); }
#undef GPASS
    }
    CK(cudaGetLastError());
    if (m->timing) { CK(cudaEventRecord(m->evs[m->ev_used + 1], m->stream)); m->ev_used += 2; }
    return B200MC_OK;
}

// One MCS: colour 0, colour 1; no halo refresh (the wrap is index arithmetic).  When the caller measured after the
// previous update (the drivers' loop), the second pass accumulates E / M itself.
int torus_sweep(Torus* m, bool allow_fuse = true)
{
    if (m->fused_pending) m->want_fused = false;   // the sums of the previous sweep were never asked for
    const bool fuse = allow_fuse && m->want_fused;
    for (int colour = 0; colour < 2; ++colour) {
        const int rc = m->ndim == 3 ? torus_launch_pass<6>(m, colour, fuse && colour == 1, fuse) : torus_launch_pass<4>(m, colour, fuse && colour == 1, fuse);
        if (rc) return rc;
    }
    m->fused_pending = fuse;
    m->draw += 1;
    return B200MC_OK;
}

int torus_measure(Torus* m, int64_t* e, int64_t* mag)
{
    if (!m->obs_valid) {
        const bool direct = m->fused_pending && m->h_acc_pending;   // already stored in h_acc by the kernel
        if (!m->fused_pending) {
            CK(cudaMemsetAsync(m->d_acc, 0, 2 * sizeof(unsigned long long), m->stream));
            COUNT_LAUNCH();
            const int zwrap = m->nranks == 1;
            if (m->ndim == 3) torus_measure_kernel<6><<<m->grid, 256, 0, m->stream>>>(m->vec[0], m->vec[1], m->R, (int)m->ny, (int)m->nz, m->nvec, m->d_acc, zwrap, (int)m->z0);
            else torus_measure_kernel<4><<<m->grid, 256, 0, m->stream>>>(m->vec[0], m->vec[1], m->R, (int)m->ny, (int)m->nz, m->nvec, m->d_acc, zwrap, (int)m->z0);
            CK(cudaGetLastError());
        }
        m->fused_pending = false;
        m->h_acc_pending = false;
        m->want_fused = true;
        if (m->nranks > 1) {   // every rank gets the sums of the whole lattice
            const int rc = dist_allreduce_u64(m->comm, m->d_acc, 2, m->stream);
            if (rc) return rc;
        }
        if (!direct) CK(cudaMemcpyAsync(m->h_acc, m->d_acc, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, m->stream));
        CK(cudaStreamSynchronize(m->stream));
        const int64_t X = (int64_t)m->h_acc[0], sum = (int64_t)m->h_acc[1];
        const int nnb = m->ndim == 3 ? 6 : 4;
        m->obs_e = -(int64_t)(nnb / 2) * m->N_glob + 2 * X;   // E = -(bonds) + 2 X, bonds = (nnb / 2) N
        m->obs_m = 2 * sum - m->N_glob;
        m->obs_valid = true;
    }
    if (e) *e = m->obs_e;
    if (mag) *mag = m->obs_m;
    return B200MC_OK;
}

int torus_set_random(Torus* m)
{
    m->obs_valid = false; m->fused_pending = false;
    for (int c = 0; c < 2; ++c) {
        COUNT_LAUNCH();
        ring_random_bits_kernel<<<(unsigned)((m->nvec + 255) / 256), 256, 0, m->stream>>>(m->vec[c], m->nvec, 0, m->z0 * (int64_t)m->R * m->ny, m->seed, m->draw, (uint32_t)c, 0);
    }
    CK(cudaGetLastError());
    if (m->nranks > 1)
        for (int c = 0; c < 2; ++c) { const int rc = torus_halo(m, c, m->stream); if (rc) return rc; }
    m->draw += 1;
    return B200MC_OK;
}

int torus_update_with_randoms(Torus* m, const double* randoms)
{
    if (!randoms) ARG_FAIL("null randoms");
    if (m->nranks > 1) { snprintf(g_b200mc_err, sizeof(g_b200mc_err), "update_with_randoms: single-GPU handles only"); return B200MC_ERR_UNSUPPORTED; }
    if (!m->d_randoms) CK(cudaMalloc(&m->d_randoms, (size_t)m->N * sizeof(double)));
    CK(cudaMemcpyAsync(m->d_randoms, randoms, (size_t)m->N * sizeof(double), cudaMemcpyHostToDevice, m->stream));
    m->obs_valid = false; m->fused_pending = false;
    const unsigned grid = (unsigned)((m->nvec + 255) / 256);
    for (int colour = 0; colour < 2; ++colour) {
        TorusArgs t;
        torus_args(m, colour, t);
        COUNT_LAUNCH();
        if (m->ndim == 3) {
            if (m->method == METHOD_METROPOLIS) torus_pass_randoms_kernel<6, METHOD_METROPOLIS><<<grid, 256, 0, m->stream>>>(t, m->tabs.tabf, m->d_randoms);
            else torus_pass_randoms_kernel<6, METHOD_HEATBATH><<<grid, 256, 0, m->stream>>>(t, m->tabs.tabf, m->d_randoms);
        } else {
            if (m->method == METHOD_METROPOLIS) torus_pass_randoms_kernel<4, METHOD_METROPOLIS><<<grid, 256, 0, m->stream>>>(t, m->tabs.tabf, m->d_randoms);
            else torus_pass_randoms_kernel<4, METHOD_HEATBATH><<<grid, 256, 0, m->stream>>>(t, m->tabs.tabf, m->d_randoms);
        }
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(m->stream));   // the host array may be reused by the caller
    return B200MC_OK;
}

int torus_io(Torus* m, int32_t* out, const int32_t* in)
{
    if (!m->d_io) CK(cudaMalloc(&m->d_io, (size_t)m->N * sizeof(int32_t)));
    const int pm1 = m->ndim == 2;   // value conventions of the reference types: 2D -1 / +1, 3D 0 / 1
    const unsigned grid = (unsigned)((m->N + 255) / 256);
    if (in) {
        for (int64_t i = 0; i < m->N; ++i) {
            const int32_t s = in[i];
            if (pm1 ? (s != 1 && s != -1) : (s != 0 && s != 1)) ARG_FAIL("set_spins: value %d at element %lld is not a spin of this model", s, (long long)i);
        }
        CK(cudaMemcpyAsync(m->d_io, in, (size_t)m->N * sizeof(int32_t), cudaMemcpyHostToDevice, m->stream));
        COUNT_LAUNCH();
        torus_import_kernel<<<grid, 256, 0, m->stream>>>(reinterpret_cast<uint8_t*>(m->vec[0]), reinterpret_cast<uint8_t*>(m->vec[1]), m->nx, m->ny, m->N, m->d_io, pm1, m->z0);
        CK(cudaGetLastError());
        m->obs_valid = false; m->fused_pending = false;
        if (m->nranks > 1)
            for (int c = 0; c < 2; ++c) { const int rc = torus_halo(m, c, m->stream); if (rc) return rc; }
    } else {
        COUNT_LAUNCH();
        torus_export_kernel<<<grid, 256, 0, m->stream>>>(reinterpret_cast<const uint8_t*>(m->vec[0]), reinterpret_cast<const uint8_t*>(m->vec[1]), m->nx, m->ny, m->N, m->d_io, pm1, m->z0);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(out, m->d_io, (size_t)m->N * sizeof(int32_t), cudaMemcpyDeviceToHost, m->stream));
    }
    CK(cudaStreamSynchronize(m->stream));
    return B200MC_OK;
}

}  // namespace

#define T(h) (reinterpret_cast<Torus*>(h))
#define CHECK_T(h)                                                                            \
    do {                                                                                      \
        if (!(h) || T(h)->magic != TORUS_MAGIC || !T(h)->alive) ARG_FAIL("invalid handle");   \
    } while (0)

extern "C" {

int b200mc_ising_torus_create(void** h, int32_t ndim, int64_t nx, int64_t ny, int64_t nz, double kbt, int32_t iseed)
{
    return torus_create(h, ndim, nx, ny, nz, kbt, iseed);
}
int b200mc_ising_torus_create_slab(void** h, int32_t ndim, int64_t nx, int64_t ny, int64_t nz, double kbt, int32_t iseed, int32_t rank,
                                   int32_t nranks, const char nccl_id[128])
{
    return torus_create(h, ndim, nx, ny, nz, kbt, iseed, rank, nranks, nccl_id);
}
int b200mc_ising_torus_p2p_handles(void* h, char out[192])
{
    CHECK_T(h);
    Torus* m = T(h);
    if (m->nranks < 2) ARG_FAIL("p2p_handles: not a slab handle");
    if (!m->flags) {
        CK(cudaMalloc(&m->flags, TORUS_FLAG_WORDS * sizeof(unsigned int)));
        CK(cudaMemset(m->flags, 0, TORUS_FLAG_WORDS * sizeof(unsigned int)));
    }
    static_assert(sizeof(cudaIpcMemHandle_t) * 3 == TORUS_IPC_BYTES, "cudaIpcMemHandle_t size");
    cudaIpcMemHandle_t hd[3];
    CK(cudaIpcGetMemHandle(&hd[0], m->alloc[0]));
    CK(cudaIpcGetMemHandle(&hd[1], m->alloc[1]));
    CK(cudaIpcGetMemHandle(&hd[2], m->flags));
    memcpy(out, hd, TORUS_IPC_BYTES);
    return B200MC_OK;
}
int b200mc_ising_torus_p2p_connect(void* h, const char prev[192], const char next[192])
{
    CHECK_T(h);
    Torus* m = T(h);
    if (m->nranks < 2) ARG_FAIL("p2p_connect: not a slab handle");
    if (!m->flags) ARG_FAIL("p2p_connect: call p2p_handles first");
    const char* src[2] = {prev, next};
    const int nopen = m->nranks == 2 ? 1 : 2;   // with two ranks both neighbours are the same process
    for (int side = 0; side < nopen; ++side) {
        cudaIpcMemHandle_t hd[3];
        memcpy(hd, src[side], TORUS_IPC_BYTES);
        void* ptr[3] = {nullptr, nullptr, nullptr};
        for (int j = 0; j < 3; ++j) {
            const cudaError_t e = cudaIpcOpenMemHandle(&ptr[j], hd[j], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                snprintf(g_b200mc_err, sizeof(g_b200mc_err), "cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
                cudaGetLastError();
                for (int i = 0; i < m->n_peer_maps; ++i) cudaIpcCloseMemHandle(m->peer_maps[i]);
                m->n_peer_maps = 0;
                return B200MC_ERR_UNSUPPORTED;
            }
            m->peer_maps[m->n_peer_maps++] = ptr[j];
        }
        m->peer_alloc[side][0] = (uint4*)ptr[0];
        m->peer_alloc[side][1] = (uint4*)ptr[1];
        m->peer_flags[side] = (unsigned int*)ptr[2];
    }
    if (nopen == 1) {
        m->peer_alloc[1][0] = m->peer_alloc[0][0];
        m->peer_alloc[1][1] = m->peer_alloc[0][1];
        m->peer_flags[1] = m->peer_flags[0];
    }
    m->p2p = true;
    return B200MC_OK;
}
int b200mc_ising_torus_rank_info(void* h, int32_t* rank, int32_t* nranks, int64_t* z0, int64_t* nz_local)
{
    CHECK_T(h);
    if (rank) *rank = T(h)->rank;
    if (nranks) *nranks = T(h)->nranks;
    if (z0) *z0 = T(h)->z0;
    if (nz_local) *nz_local = T(h)->ndim == 3 ? T(h)->nz : 0;
    return B200MC_OK;
}
int b200mc_ising_torus_destroy(void* h)
{
    if (!h) return B200MC_OK;
    CHECK_T(h);
    return torus_destroy(T(h));
}
int b200mc_ising_torus_set_stream(void* h, void* s) { CHECK_T(h); T(h)->stream = (cudaStream_t)s; return B200MC_OK; }
int b200mc_ising_torus_skip_curand(void* h, int64_t n_skip)
{
    CHECK_T(h);
    if (n_skip < 0) ARG_FAIL("n_skip < 0");
    T(h)->draw += (uint64_t)((n_skip + T(h)->N_glob - 1) / T(h)->N_glob);   // as the helical modules: ceil(n_skip / nall) draws
    return B200MC_OK;
}
int b200mc_ising_torus_set_allup_spin(void* h)
{
    CHECK_T(h);
    T(h)->obs_valid = false; T(h)->fused_pending = false;
    // slab mode: the owned planes only; the ghost planes are refreshed through the two-sided exchange, so that no neighbour can
    // push into a ghost plane this rank is about to overwrite (direct transport)
    CK(cudaMemsetAsync(T(h)->vec[0], 1, (size_t)T(h)->nvec * 16, T(h)->stream));
    CK(cudaMemsetAsync(T(h)->vec[1], 1, (size_t)T(h)->nvec * 16, T(h)->stream));
    if (T(h)->nranks > 1)
        for (int c = 0; c < 2; ++c) { const int rc = torus_halo(T(h), c, T(h)->stream); if (rc) return rc; }
    return B200MC_OK;
}
int b200mc_ising_torus_set_random_spin(void* h) { CHECK_T(h); return torus_set_random(T(h)); }
int b200mc_ising_torus_set_kbt(void* h, double kbt) { CHECK_T(h); if (!(kbt > 0.0)) ARG_FAIL("kbt must be > 0"); T(h)->beta = 1 / kbt; return torus_tables(T(h)); }
int b200mc_ising_torus_set_beta(void* h, double beta) { CHECK_T(h); if (!(beta >= 0.0)) ARG_FAIL("beta must be >= 0"); T(h)->beta = beta; return torus_tables(T(h)); }
int b200mc_ising_torus_set_method(void* h, int32_t method)
{
    CHECK_T(h);
    if (method != METHOD_METROPOLIS && method != METHOD_HEATBATH) ARG_FAIL("unknown method %d", method);
    T(h)->method = method;
    return torus_tables(T(h));
}
int b200mc_ising_torus_update(void* h) { CHECK_T(h); return torus_sweep(T(h)); }
int b200mc_ising_torus_update_n(void* h, int32_t n)
{
    CHECK_T(h);
    for (int i = 0; i < n; ++i) { const int rc = torus_sweep(T(h), i == n - 1); if (rc) return rc; }
    return B200MC_OK;
}
int b200mc_ising_torus_update_with_randoms(void* h, const double* randoms) { CHECK_T(h); return torus_update_with_randoms(T(h), randoms); }
int b200mc_ising_torus_calc_energy_sum(void* h, int64_t* e) { CHECK_T(h); return torus_measure(T(h), e, nullptr); }
int b200mc_ising_torus_calc_magne_sum(void* h, int64_t* m) { CHECK_T(h); return torus_measure(T(h), nullptr, m); }
int b200mc_ising_torus_measure(void* h, int64_t* e, int64_t* m) { CHECK_T(h); return torus_measure(T(h), e, m); }
int b200mc_ising_torus_get_spins(void* h, int32_t* out) { CHECK_T(h); if (!out) ARG_FAIL("null output"); return torus_io(T(h), out, nullptr); }
int b200mc_ising_torus_set_spins(void* h, const int32_t* in) { CHECK_T(h); if (!in) ARG_FAIL("null input"); return torus_io(T(h), nullptr, in); }
int64_t b200mc_ising_torus_nall(void* h) { return (h && T(h)->magic == TORUS_MAGIC) ? T(h)->N_glob : -1; }
int64_t b200mc_ising_torus_nx(void* h) { return (h && T(h)->magic == TORUS_MAGIC) ? T(h)->nx : -1; }
int64_t b200mc_ising_torus_ny(void* h) { return (h && T(h)->magic == TORUS_MAGIC) ? T(h)->ny : -1; }
int64_t b200mc_ising_torus_nz(void* h) { return (h && T(h)->magic == TORUS_MAGIC) ? (T(h)->ndim == 3 ? T(h)->nz_glob : 0) : -1; }
double b200mc_ising_torus_beta(void* h) { return (h && T(h)->magic == TORUS_MAGIC) ? T(h)->beta : 0.0; }
double b200mc_ising_torus_kbt(void* h) { return (h && T(h)->magic == TORUS_MAGIC) ? 1 / T(h)->beta : 0.0; }
int b200mc_ising_torus_get_table(void* h, double out[16]) { CHECK_T(h); for (int i = 0; i < 16; ++i) out[i] = T(h)->tabs.w[i]; return B200MC_OK; }
int b200mc_ising_torus_set_timing(void* h, int32_t on) { CHECK_T(h); T(h)->timing = on != 0; T(h)->ev_used = 0; return B200MC_OK; }
int b200mc_ising_torus_get_timing(void* h, int64_t* launches, double* total_ms)
{
    CHECK_T(h);
    CK(cudaStreamSynchronize(T(h)->stream));
    double tot = 0.0;
    for (size_t i = 0; i + 1 < T(h)->ev_used; i += 2) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, T(h)->evs[i], T(h)->evs[i + 1]));
        tot += ms;
    }
    if (launches) *launches = (int64_t)(T(h)->ev_used / 2);
    if (total_ms) *total_ms = tot;
    return B200MC_OK;
}
int b200mc_ising_torus_sync(void* h) { CHECK_T(h); CK(cudaStreamSynchronize(T(h)->stream)); return B200MC_OK; }

}  // extern "C"
