// Device kernels for the helical Ising models on the folded-ring layout.
//
// One colour pass (reference: update_sub, src/ising3d_gpu_m.f90:189-206 and
// src/ising2d_gpu_m.f90:148-162) processes, per thread and iteration, ONE
// 128-bit vector = 16 sites of the colour being updated:
//   * 1 + nnb aligned 128-bit loads (own vector, nnb neighbour vectors of the
//     other colour), 1 aligned 128-bit store;
//   * neighbour sums, table lookup and accept test are byte-parallel inside
//     32-bit registers (4 sites per instruction); the per-site acceptance
//     threshold is fetched with PRMT used as an 8-entry byte lookup table;
//   * one Philox4x32-10 call yields the 16 random bytes of the vector.
//
// Accept test.  The reference accepts iff u <= w(S, s) with u a real64 uniform
// and w a real64 table.  Here u has 32-bit resolution, u = (U+1) 2^-32, and
// u <= w  <=>  U < thr,  thr = floor(w 2^32)  (0 <= thr <= 2^32).  U is built
// lazily: its top 7 bits b7 come from the vector's Philox block; only when
// b7 == thr >> 25 (probability 2^-7 per undecided site) are the low 25 bits
// drawn from a second Philox block.  Stored per (S, s): T' = 128 - (thr >> 25)
// (a byte) and thr & (2^25 - 1).
#pragma once
#include "common.cuh"
#include "ring.cuh"
#include <cooperative_groups.h>

struct IsingTab {
    uint32_t tlo, thi;     // bytes T'(idx) for idx = 0..3 / 4..7
    uint32_t low25[8];     // thr & (2^25-1) per idx
    uint32_t rk0[10];      // Philox round keys seed + r*W0 (uniform: precomputed on the host)
};
// Table index: METROPOLIS idx = number of neighbours ALIGNED with the site's own spin
// (k' = s ? S : nnb - S; the zero-field acceptance ws(S, s) only depends on it),
// HEATBATH idx = S (number of up neighbours).

// Philox4x32-10 with the key schedule supplied as constants (rk0 from the host,
// k1 = tag + r*W1 folded at compile time): 20 IMAD.WIDE + 20 LOP3 per block.
template <uint32_t TAG>
__device__ __forceinline__ uint4 philox_rk(uint4 c, const uint32_t (&rk0)[10])
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

struct RingPassArgs {
    uint4* own;         // colour being updated, vector index 0 = position -H
    const uint4* oth;   // the other colour
    int64_t nvec;       // owned positions (L); L + 2H < 2^31 is checked at create
    int64_t H;
    int64_t p0;         // global position of the first owned vector
    int64_t off[6];     // neighbour vector offsets for this colour
    uint32_t seed;
    uint32_t colour;
    uint64_t draw;
    unsigned int* ticket;  // work counter for ordered scheduling (nullptr: static round-robin)
    // Self-cleaning launches (single-GPU ticket passes): the pass of colour c counts on counter set c and block 0 clears
    // set c ^ 1 (idle: the pass that used it has finished, the one that will has not started) -- no memset between the
    // passes.  acc_reset: the first pass of a sweep clears the sums the fused second pass will add to.  host_out: the
    // last warp of the fused pass to finish stores the two sums straight into pinned host memory (done_warps counts).
    unsigned int* ticket_reset;
    unsigned long long* acc_reset;
    unsigned long long* host_out;
    unsigned int* done_warps;
    int chunk;             // vectors per ticket (multiple of 32)
    // PUSH variant (slab mode: colour pass fused with the halo exchange).  The first nb owned vectors
    // are the HIGH halo of rank-1, the vectors from hi_start on the LOW halo of rank+1: the kernel
    // stores those results a second time straight into the neighbour's halo over NVLink (peer pointers
    // from cudaIpcOpenMemHandle), rotated by one byte-lane where the ring closes (rank 0 <-> rank P-1).
    // Chunks [0, blo) and [jhi, nchunks) hold them ("boundary chunks", nbchunks in all); the other
    // chunks are interior.  q_total = number of tickets (virtual chunks) of the launch.
    uint4* peer_lo;        // rank-1's high halo of this colour
    uint4* peer_hi;        // rank+1's low halo of this colour
    int rot_lo, rot_hi;    // lane rotation of the pushed copy: +1 lane b <- b-1, -1 lane b <- b+1, 0 none
    int nb;                // H
    int hi_start;          // Lloc - H
    int nbchunks, q_total, nopush;
    int hi_tickets, hi_first, lo_end;  // in vectors: 128 * (chunks from jhi on), 128 * jhi, 128 * blo
    unsigned long long* dbg_wait;  // debug: ns block 0 spent waiting for the neighbours' flags (summed over launches)
    unsigned int* done;    // completed boundary chunks of this launch (local)
    unsigned int* sig_prev;       // rank-1's "from next" flag, rank+1's "from prev" flag (peer memory)
    unsigned int* sig_next;
    const unsigned int* wait_prev;  // my flags: pushes received so far from rank-1 / rank+1
    const unsigned int* wait_next;
    unsigned int wait_seq, sig_seq;
    // MEASURE variant (second colour pass of a sweep): acc[0] += X, acc[1] += sum(s) as ising_measure_kernel
    unsigned long long* acc;
    // batch of independent samples (blockIdx.y): sample j lives rstride vectors further in both colour arrays,
    // draws the counters (position, j, draw, ...) and adds its sums to acc[2 j], acc[2 j + 1]
    int64_t rstride;
    // MEASURE on folds whose last positions hold no site in the high lanes (16 does not divide Nc): vectors from
    // mask_from on (local index) keep only the lanes b with b * Lfold + position < Nc in the sums
    int mask_from;
    int64_t Lfold, Nc;
    int wrap_mode;   // cooperative sweep kernel: 0 = the halo is refreshed after every pass, read it; 1 / 2 = no refresh between the passes,
                     // wrapped positions are rebuilt from the owned sites (1: H <= L and ptail >= H, vector-granular; 2: byte-granular)
};

__device__ __forceinline__ uint4 rot_lanes(uint4 s, int dir)
{
    uint4 o = s;
    if (dir > 0) {
        o.x = __funnelshift_l(s.w, s.x, 8); o.y = __funnelshift_l(s.x, s.y, 8);
        o.z = __funnelshift_l(s.y, s.z, 8); o.w = __funnelshift_l(s.z, s.w, 8);
    } else if (dir < 0) {
        o.x = __funnelshift_r(s.x, s.y, 8); o.y = __funnelshift_r(s.y, s.z, 8);
        o.z = __funnelshift_r(s.z, s.w, 8); o.w = __funnelshift_r(s.w, s.x, 8);
    }
    return o;
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p)
{
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

enum { METHOD_METROPOLIS = 0, METHOD_HEATBATH = 1 };

// Deferred tie resolution.  A tie (b7 == thr >> 25) happens for 1 site in 128,
// i.e. in almost every warp-iteration, so resolving it inline would make every
// warp pay for a second Philox block.  Instead the (rare) lane that sees a tie
// treats it as "reject", pushes a 32-byte record into a per-warp shared-memory
// queue and goes on; when the queue is half full (and at kernel end) the warp
// drains it with all 32 lanes busy and patches the accepted bytes in global memory.
#define TQ_CAP 96   // records per warp: drained when more than TQ_CAP - 64 are parked, checked every second vector
#define TK_NCNT 8

// keep a value in a register: stops the compiler from rematerialising address
// arithmetic from kernel parameters inside the hot loop
template <typename T>
__device__ __forceinline__ void pin64(T*& p) { asm volatile("" : "+l"(p)); }
__device__ __forceinline__ void pin32(uint32_t& x) { asm volatile("" : "+r"(x)); }

__device__ __forceinline__ uint32_t lds32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// stage 1 for 8 sites: own words (w0 = lanes 0-3, w1 = lanes 4-7 of the group),
// neighbour sums (s0, s1), random words (ra -> lanes (0,4,1,5), rb -> lanes (2,6,3,7)).
// Returns the compare words zA, zB (bit 7 of a byte clear <=> accept), the permuted own
// words oA, oB and ix = nibble-packed table indices | own spin << 3 (selector order).
template <int NNB, int METHOD>
__device__ __forceinline__ void ising_stage1(uint32_t w0, uint32_t w1, uint32_t s0, uint32_t s1,
                                             uint32_t ra, uint32_t rb, const IsingTab& tab,
                                             uint32_t& ix, uint32_t& oA, uint32_t& oB,
                                             uint32_t& zA, uint32_t& zB)
{
    const uint32_t sp = s0 + (s1 << 4);  // nibble-packed sums: byte j = S(lane j) | S(lane 4+j) << 4
    const uint32_t op = w0 + (w1 << 4);  // nibble-packed own spins, same order
    uint32_t idx;
    if (METHOD == METHOD_METROPOLIS) {
        // k' = s ? S : nnb - S, per nibble:  nnb - S = (S ^ 7) - (7 - nnb)   (no borrows: S <= nnb)
        const uint32_t t = op ^ 0x11111111u;        // 1 where the spin is down
        idx = (sp ^ (t * 7u)) - t * (uint32_t)(7 - NNB);
    } else {
        idx = sp;
    }
    ix = idx + op * 8u;
    oA = prmt(w0, w1, 0x5140u);  // lanes (0,4,1,5)
    oB = prmt(w0, w1, 0x7362u);  // lanes (2,6,3,7)
    const uint32_t tA = prmt(tab.tlo, tab.thi, idx);
    const uint32_t tB = prmt(tab.tlo, tab.thi, idx >> 16);
    // z = b7 + 128 - T7 per byte (no carries: <= 255).  bit 7 clear <=> b7 < T7 <=> accept.
    zA = (ra & 0x7F7F7F7Fu) + tA;
    zB = (rb & 0x7F7F7F7Fu) + tB;
}

template <int METHOD>
__device__ __forceinline__ void ising_finish(uint32_t& w0, uint32_t& w1, uint32_t oA, uint32_t oB,
                                             uint32_t zA, uint32_t zB)
{
    const uint32_t sA = zA >> 7, sB = zB >> 7;
    uint32_t nA, nB;
    if (METHOD == METHOD_METROPOLIS) { nA = oA ^ (~sA & 0x01010101u); nB = oB ^ (~sB & 0x01010101u); }
    else { nA = ~sA & 0x01010101u; nB = ~sB & 0x01010101u; }
    w0 = prmt(nA, nB, 0x6420u);
    w1 = prmt(nA, nB, 0x7531u);
}

// conservative tie flags: bit 7 set in every byte equal to 0x80 (may also flag the byte
// above a 0x00 byte; the drain re-tests exactly)
__device__ __forceinline__ uint32_t tie_flags(uint32_t z) { return z & ~(z - 0x01010101u); }

// exact per-byte "== 0x80" test -> 4-bit mask
__device__ __forceinline__ uint32_t tie_bits(uint32_t z)
{
    const uint32_t t = z ^ 0x80808080u;
    const uint32_t e = ~(((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t) & 0x80808080u;  // 0x80 where byte == 0
    return ((e >> 7) * 0x00204081u) >> 21 & 0xFu;  // gather bits 0,8,16,24 -> 0..3
}

// Returns (dX, dM): what the ties accepted here add to this lane's running sums of the MEASURE
// variant (the main loop counted them as rejected).  NNB = 0: not measuring.
template <int METHOD, bool PUSH, int NNB>
__device__ __noinline__ int2 ising_drain(uint32_t qaddr, uint32_t cntaddr, uint4* own, const RingPassArgs& a,
                                         const IsingTab& tab, uint32_t rep = 0)
{
    int2 delta = make_int2(0, 0);
    __syncwarp();
    const uint32_t n = lds32(cntaddr);
    const int lane = threadIdx.x & 31;
    for (uint32_t r = lane; r < n; r += 32) {
        uint4 r0, r1;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r0.x), "=r"(r0.y), "=r"(r0.z), "=r"(r0.w) : "r"(qaddr + r * 32));
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r1.x), "=r"(r1.y), "=r"(r1.z), "=r"(r1.w) : "r"(qaddr + r * 32 + 16));
        const uint32_t v = r0.x;
        uint8_t* rbytes = nullptr;  // PUSH: the pushed copy of this vector in the neighbour's halo
        int rrot = 0;
        if (PUSH) {
            if ((int)v < a.nb) { rbytes = reinterpret_cast<uint8_t*>(a.peer_lo + v); rrot = a.rot_lo; }
            else if ((int)v >= a.hi_start) { rbytes = reinterpret_cast<uint8_t*>(a.peer_hi + ((int)v - a.hi_start)); rrot = a.rot_hi; }
        }
        uint8_t* bytes = reinterpret_cast<uint8_t*>(own + v);
        // 16-bit tie mask, bit m = byte position m of the stage-1 Philox block
        uint32_t mask = tie_bits(r0.y) | (tie_bits(r0.z) << 4) | (tie_bits(r0.w) << 8) | (tie_bits(r1.x) << 12);
        const uint64_t pglob = (uint64_t)(a.p0 + v) | ((uint64_t)rep << 32);
        while (mask) {
            const int m = __ffs(mask) - 1;
            mask &= mask - 1;
            const int w = m >> 2, j = m & 3;
            const uint4 R = philox_rk<TAG_ISING>(mk_ctr(pglob, a.draw, a.colour, 1u + w), tab.rk0);
            const uint32_t rj = j == 0 ? R.x : j == 1 ? R.y : j == 2 ? R.z : R.w;
            const uint32_t ixw = (w >> 1) ? r1.z : r1.y;
            const uint32_t nib = (ixw >> (16 * (w & 1) + 4 * j)) & 0xFu;
            if ((rj & 0x1FFFFFFu) < tab.low25[nib & 7u]) {
                const int lb = (m & 8) | ((m & 7) >> 1) | ((m & 1) << 2);  // byte position -> lane
                const uint8_t nv = (METHOD == METHOD_METROPOLIS) ? (uint8_t)((nib >> 3) ^ 1u) : (uint8_t)1;
                bytes[lb] = nv;
                if (PUSH && rbytes) rbytes[(lb + rrot) & 15] = nv;
                if (NNB && (int64_t)lb * a.Lfold + (a.p0 + v) < a.Nc) {   // (a site-less tail lane is not part of the sums)
                    // Metropolis: nib & 7 = k' aligned neighbours of the old spin; counted NNB - k' unequal, now k'.
                    // Heat-bath: nib & 7 = S up neighbours; counted as down (S unequal), now up (NNB - S).
                    const int t = (int)(nib & 7u);
                    delta.x += (METHOD == METHOD_METROPOLIS) ? 2 * t - NNB : NNB - 2 * t;
                    delta.y += (METHOD == METHOD_METROPOLIS) ? (nv ? 1 : -1) : 1;
                }
            }
        }
    }
    __syncwarp();
    if (lane == 0) asm volatile("st.shared.u32 [%0], %1;" ::"r"(cntaddr), "r"(0u) : "memory");
    __syncwarp();
    return delta;
}

// One vector (16 sites of the colour being updated) of one lane: loads, Philox block, byte-parallel
// accept test, store; ties parked in the warp's queue; optional halo push and fused E/M sums.
// MEASURE on a fold with site-less tail positions: byte mask of the lanes of vector v that hold a site
static __device__ __noinline__ uint4 ising_tail_keep(int64_t pg, int64_t Lfold, int64_t Nc)
{
    uint32_t keep[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        keep[w] = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if ((int64_t)(4 * w + j) * Lfold + pg < Nc) keep[w] |= 0xFFu << (8 * j);
    }
    return make_uint4(keep[0], keep[1], keep[2], keep[3]);
}

// byte-wise partial sums of ising_core -> the lane's running totals
__device__ __forceinline__ void ising_fold_sums(uint32_t& bX, uint32_t& bM, uint32_t& accX, uint32_t& accM)
{
    // sum of the four byte lanes.  Each lane is < 256 (the callers fold after at most 4 vectors: <= 96 / <= 32 per lane), but the
    // lanes of bX together can reach 384, so the one-multiply byte sum (total < 256) is only good for bM; bX goes through
    // 16-bit fields first
    accX += (((bX & 0x00FF00FFu) + ((bX >> 8) & 0x00FF00FFu)) * 0x00010001u) >> 16;
    accM += (bM * 0x01010101u) >> 24;
    bX = 0u; bM = 0u;
}

// everything after the loads: o = own vector, nb = the NNB neighbour vectors
template <int NNB, int METHOD, bool PUSH, bool MEASURE>
__device__ __forceinline__ void ising_core(int v, uint4* po, uint4 o, const uint4 (&nb)[NNB], uint32_t cx, uint32_t cz, uint32_t cw,
                                           const RingPassArgs& a, const IsingTab& tab, uint64_t pol, uint32_t qaddr,
                                           uint32_t cntaddr, bool is_b, uint32_t& accX, uint32_t& accM, uint32_t cy = 0u)
{
    // counter (p0 + v, 0, draw_lo, draw_hi | colour << 16 | sub << 24): positions are < 2^31
    const uint4 r = philox_rk<TAG_ISING>(make_uint4(cx, cy, cz, cw), tab.rk0);
    uint4 S = make_uint4(nb[0].x + nb[1].x, nb[0].y + nb[1].y, nb[0].z + nb[1].z, nb[0].w + nb[1].w);
#pragma unroll
    for (int j = 2; j < NNB; ++j) { S.x += nb[j].x; S.y += nb[j].y; S.z += nb[j].z; S.w += nb[j].w; }
    uint32_t ix0, ix1, oA0, oB0, oA1, oB1, zA0, zB0, zA1, zB1;
    ising_stage1<NNB, METHOD>(o.x, o.y, S.x, S.y, r.x, r.y, tab, ix0, oA0, oB0, zA0, zB0);
    ising_stage1<NNB, METHOD>(o.z, o.w, S.z, S.w, r.z, r.w, tab, ix1, oA1, oB1, zA1, zB1);
    const uint32_t tie = (tie_flags(zA0) | tie_flags(zB0) | tie_flags(zA1) | tie_flags(zB1)) & 0x80808080u;
    if (tie) {  // rare per lane: park the record, resolve later (ties count as reject below)
        uint32_t slot;
        asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(slot) : "r"(cntaddr) : "memory");
        const uint32_t ra = qaddr + slot * 32;
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(ra), "r"((uint32_t)v), "r"(zA0), "r"(zB0), "r"(zA1) : "memory");
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(ra + 16), "r"(zB1), "r"(ix0), "r"(ix1), "r"(0u) : "memory");
    }
    ising_finish<METHOD>(o.x, o.y, oA0, oB0, zA0, zB0);
    ising_finish<METHOD>(o.z, o.w, oA1, oB1, zA1, zB1);
    st_own(po, o, pol);
    if (MEASURE) {
        // fused E + M on the values this pass leaves behind.  With s the new spins of this colour and S the number
        // of up neighbours, the unequal-neighbour count is  X = sum_{s=0} S + sum_{s=1} (NNB - S)
        //   = sum S - 2 sum S s + NNB sum s,  and  sum S (over all sites of this colour) = NNB sum(s of the other colour),
        // so the launch only accumulates  accX = sum S s  and  accM = sum s (both colours):  X = NNB accM - 2 accX.
        // (IDP.4A is a slow pipe on sm_100: the byte sums use full-rate IMAD / LOP3 / IADD3 instead.)
        uint4 om = o, n0 = nb[0];
        if (v >= a.mask_from) {   // rare: the few tail positions of the fold
            const uint4 keep = ising_tail_keep(a.p0 + v, a.Lfold, a.Nc);
            om.x &= keep.x; om.y &= keep.y; om.z &= keep.z; om.w &= keep.w;
            n0.x &= keep.x; n0.y &= keep.y; n0.z &= keep.z; n0.w &= keep.w;
        }
        // accX / accM are BYTE-WISE partial sums here (four byte lanes each): a vector adds at most 24 / 8 per byte, so the caller
        // folds them (ising_fold_sums) at least every 10 vectors -- once per ticket in the unrolled loop instead of once per vector
        accX += ((S.x & (om.x * 255u)) + (S.y & (om.y * 255u))) + ((S.z & (om.z * 255u)) + (S.w & (om.w * 255u)));
        accM += ((om.x + om.y) + (om.z + om.w)) + ((n0.x + n0.y) + (n0.z + n0.w));
    }
    if (PUSH && is_b && !(a.nopush & 1)) {  // second copy straight into the neighbour's halo (NVLink store)
        if (v < a.nb) a.peer_lo[v] = rot_lanes(o, a.rot_lo);
        else if (v >= a.hi_start) a.peer_hi[v - a.hi_start] = rot_lanes(o, a.rot_hi);
    }
}

// COHERENT: 0 = other colour through the read-only (nc) path; 1 = cooperative sweep kernel (L2 loads, wrapped positions
// rebuilt from the owned sites); 2 = plain L2 loads -- the boundary tickets of a slab pass, whose halo vectors the
// neighbouring GPUs store over NVLink while earlier blocks of this launch may already have pulled the same 128-byte
// line through L1 (an acquire on the flags does not invalidate the nc path)
template <int NNB, int METHOD, bool PUSH, bool MEASURE, bool FULLWARP, int COHERENT = 0>
__device__ __forceinline__ void ising_vec(int v, uint4* po, const uint4* const (&q)[NNB], uint32_t cx, uint32_t cz, uint32_t cw,
                                          const RingPassArgs& a, const IsingTab& tab, uint64_t pol, uint32_t qaddr,
                                          uint32_t cntaddr, bool is_b, uint32_t& accX, uint32_t& accM, uint32_t cy = 0u)
{
    const uint4 o = ld_own(po, pol);
    uint4 nb[NNB];
    // (loading the x+ vector as a shuffle of the neighbouring lane's x- vector instead of a second, overlapping
    // 512-byte load was tried: 4 SHFL + a one-lane load made the pass 20 % slower -- the MIO queue is the busiest unit)
#pragma unroll
    for (int j = 0; j < NNB; ++j) {
        if (COHERENT == 2) nb[j] = ld_other_coherent(q[j]);
        else if (COHERENT == 1) {
            // cooperative sweep kernel (single GPU, p0 = 0), wrap_mode != 0: positions outside [0, min(ptail, L)) are rebuilt
            // from the owned sites instead of read from the halo, which is then only refreshed at the end of the launch
            const int pj = v + (int)a.off[j];
            const int safe = a.mask_from < (int)a.Lfold ? a.mask_from : (int)a.Lfold;
            nb[j] = (a.wrap_mode == 0 || (unsigned)pj < (unsigned)safe)
                        ? ld_other_coherent(q[j])
                        : ring_load_wrapped(a.oth, a.Lfold, a.H, a.Nc, (int64_t)safe, a.wrap_mode == 1, (int64_t)pj);
        } else nb[j] = ld_other(q[j]);
    }
    ising_core<NNB, METHOD, PUSH, MEASURE>(v, po, o, nb, cx, cz, cw, a, tab, pol, qaddr, cntaddr, is_b, accX, accM, cy);
}

#ifndef TK_CHUNK
#define TK_CHUNK 128
#endif

// vectors per ticket: 4 warp-iterations, unrolled so that the 8 stream addresses are
                      // computed once per ticket and reached with immediate offsets (+512 B per iteration)

// resident blocks per SM the register allocation is tuned for (measured, profiles/r01_history.md): the 3D
// kernels are fastest with ~80 registers and 3 blocks (1646 vs 1505 flips/ns with 64 registers and 4 blocks),
// the 2D kernels with 64 registers and 4 blocks
#ifndef PASS_MINB
#define PASS_MINB(NNB, PUSH) (((NNB) == 6) ? 3 : 4)
#endif
// The body of a colour pass: everything one block does for the vectors (tickets) of `a`.  tq / tq_cnt: the block's
// tie queues (8 warps x TQ_CAP records) and their fill counts in shared memory.
template <int NNB, int METHOD, bool ORDERED, bool PUSH, bool MEASURE, bool BATCH, int CH = TK_CHUNK>
__device__ __forceinline__ void ising_pass_body(const RingPassArgs& a, const IsingTab& tab, uint4 (*tq)[TQ_CAP][2], uint32_t* tq_cnt)
{
    static_assert(!PUSH || ORDERED, "the fused update + halo push kernel uses ticket scheduling");
    static_assert(CH % 32 == 0 && (CH == TK_CHUNK || !ORDERED), "tickets are TK_CHUNK vectors");
    constexpr int COH = CH != TK_CHUNK ? 1 : 0;   // 32-vector chunks = the cooperative multi-pass kernel: the other colour was written in this launch
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t qaddr = (uint32_t)__cvta_generic_to_shared(&tq[warp][0][0]);
    uint32_t cntaddr = (uint32_t)__cvta_generic_to_shared(&tq_cnt[warp]);
    pin32(qaddr);
    pin32(cntaddr);
    if (lane == 0) tq_cnt[warp] = 0;
    __syncwarp();
    if (ORDERED && !PUSH && blockIdx.x == 0) {   // self-cleaning launches (see RingPassArgs)
        if (a.ticket_reset && threadIdx.x < TK_NCNT) a.ticket_reset[threadIdx.x * 64] = 0u;
        if (!MEASURE && a.acc_reset && threadIdx.x < 2) a.acc_reset[threadIdx.x] = 0ull;
    }
    if (PUSH) {
        // the halo cells this pass reads were pushed by the neighbours during their previous pass
        if (threadIdx.x == 0) {
            unsigned long long t0 = 0, t1 = 0;
            if (blockIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            while ((int)(ld_acquire_sys(a.wait_prev) - a.wait_seq) < 0) __nanosleep(100);
            while ((int)(ld_acquire_sys(a.wait_next) - a.wait_seq) < 0) __nanosleep(100);
            if (blockIdx.x == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); atomicAdd(a.dbg_wait, t1 - t0); }
        }
        __syncthreads();
    }
    const uint64_t pol = l2_policy_evict_first();
    static_assert(!BATCH || (!ORDERED && !PUSH), "batches of samples use the static round-robin launch");
    const uint32_t rep = BATCH ? blockIdx.y : 0u;          // sample of the batch (a constant 0 keeps the single-sample code unchanged)
    uint4* own = a.own + (a.H + (int64_t)rep * a.rstride);
    const int nvec = (int)a.nvec;
    const uint4* pn[NNB];
#pragma unroll
    for (int j = 0; j < NNB; ++j) pn[j] = a.oth + (a.H + a.off[j] + (int64_t)rep * a.rstride);
    const uint32_t cx0 = (uint32_t)a.p0, cz = (uint32_t)a.draw;
    const uint32_t cw = (uint32_t)((a.draw >> 32) & 0xFFFFu) | (a.colour << 16);
    // ordered mode: every warp takes TK_CHUNK-vector chunks from a global counter, so that all
    // resident warps work inside one narrow, advancing window of the lattice (the z-neighbour
    // planes then stay in L2 between their three uses).  The next ticket is fetched one chunk
    // ahead; no block-level synchronisation anywhere in the loop.
    const int nwarps_grid = gridDim.x * (blockDim.x >> 5);
    // TK_NCNT interleaved counters (256 B apart -> different L2 slices): counter c hands out the
    // chunks c, c + NCNT, c + 2 NCNT, ...; one same-address atomic stream would cap the ticket rate
    const int gwarp = blockIdx.x * (blockDim.x >> 5) + warp;
    unsigned int* tk = a.ticket + (gwarp % TK_NCNT) * 64;
    const int tk_base = (gwarp % TK_NCNT) * CH, tk_scale = TK_NCNT;
    const int vlimit = PUSH ? a.q_total * CH : nvec;  // PUSH: tickets count VIRTUAL chunks
    uint32_t accX = 0, accM = 0;  // MEASURE: this lane's sums (< 2^31: at most ~10^5 sites per lane and launch)
    uint32_t bX = 0, bM = 0;      // byte-wise partial sums of the current ticket (ising_core), folded into accX / accM below
    int corrX = 0, corrM = 0;
    constexpr int DN = MEASURE ? NNB : 0;
    // ticket -> first vector of the chunk (-1: a virtual chunk of the PUSH schedule that maps to nothing)
    auto chunk_base = [&](int t, bool& boundary) -> int {
        boundary = false;
        if (!PUSH) return t;
        // Slab mode.  The chunks holding the LAST H owned vectors (the low halo of rank+1) are handed
        // out first; all other chunks follow in natural order, which
        // starts with the first H owned vectors (the high halo of rank-1).  Both halo blocks are thus
        // on their way over NVLink early in the pass and land while the interior is being updated.
        const int vb = t >= a.hi_tickets ? t - a.hi_tickets : t + a.hi_first;   // hi_tickets = 128 nbhi, hi_first = 128 jhi
        boundary = vb < a.lo_end || vb >= a.hi_first;
        return vb;
    };
    // The two streams of a chunk that come from DRAM (the own colour, read once per pass, and the leading
    // z / y neighbour plane, touched for the first time) are prefetched into L2 one ticket ahead: lanes
    // 0-15 fetch the 16 lines of the own chunk, lanes 16-31 those of the leading neighbour chunk.
    const char* pf_base = (lane < 16) ? reinterpret_cast<const char*>(own) : reinterpret_cast<const char*>(pn[NNB - 2]);
    pf_base += (lane & 15) * 128;
    int cur, nxt = 0, nx2 = 0;
    if (ORDERED) {
        if (lane == 0) {
            nxt = (int)atomicAdd(tk, (unsigned)CH) * tk_scale + tk_base;
            nx2 = (int)atomicAdd(tk, (unsigned)CH) * tk_scale + tk_base;
        }
        cur = __shfl_sync(0xffffffffu, nxt, 0);
        nxt = __shfl_sync(0xffffffffu, nx2, 0);
    } else {
        cur = gwarp * CH;
        nxt = cur + nwarps_grid * CH;
    }
    while (cur < vlimit) {
        if (ORDERED && lane == 0) nx2 = (int)atomicAdd(tk, (unsigned)CH) * tk_scale + tk_base;  // two tickets ahead
        // (prefetching a fixed distance ahead of the current chunk instead -- what the torus kernel does -- was measured here too:
        //  no effect, 16 K / 64 K / 128 K vectors ahead; profiles/r02ah_prefetch_distance.log)
        if (nxt < vlimit) {
            bool bn;
            const int bnext = chunk_base(nxt, bn);
            if (CH == TK_CHUNK && bnext >= 0 && bnext + CH <= nvec && !(a.nopush & 2))
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pf_base + (size_t)bnext * 16));
        }
        bool is_b;
        const int base = chunk_base(cur, is_b);
        if (PUSH && base < 0) { cur = nxt; nxt = __shfl_sync(0xffffffffu, nx2, 0); continue; }
        const int v0 = base + lane;
        uint4* po = own + v0;
        const uint4* q[NNB];
#pragma unroll
        for (int j = 0; j < NNB; ++j) q[j] = pn[j] + v0;
        const uint32_t cx = cx0 + (uint32_t)v0;
        if (base + CH <= nvec && !(PUSH && is_b)) {
            // full (interior) chunk: no bounds checks, no halo push, the queue is looked at every second vector
#pragma unroll
            for (int u = 0; u < CH / 32; ++u) {
                const uint4* qu[NNB];
#pragma unroll
                for (int j = 0; j < NNB; ++j) qu[j] = q[j] + 32 * u;
                ising_vec<NNB, METHOD, false, MEASURE, true, COH>(v0 + 32 * u, po + 32 * u, qu, cx + 32u * u, cz, cw, a, tab, pol, qaddr,
                                                             cntaddr, false, bX, bM, rep);
                if ((u & 1) || CH < 64) {   // (a 32-vector chunk is one vector per lane: look at the queue after it)
                    __syncwarp();
                    if (lds32(cntaddr) > TQ_CAP - 64) {
                        const int2 d = ising_drain<METHOD, PUSH, DN>(qaddr, cntaddr, own, a, tab, rep);
                        corrX += d.x; corrM += d.y;
                    }
                }
            }
            if (MEASURE) ising_fold_sums(bX, bM, accX, accM);   // (at most 4 vectors: bytes <= 96 / 32)
        } else {
#pragma unroll 1
            for (int u = 0; u < CH / 32; ++u) {
                if (v0 + 32 * u < nvec) {
                    const uint4* qu[NNB];
#pragma unroll
                    for (int j = 0; j < NNB; ++j) qu[j] = q[j] + 32 * u;
                    ising_vec<NNB, METHOD, PUSH, MEASURE, false, (PUSH ? 2 : COH)>(v0 + 32 * u, po + 32 * u, qu, cx + 32u * u, cz, cw, a, tab, pol, qaddr,
                                                                 cntaddr, is_b, bX, bM, rep);
                    if (MEASURE) ising_fold_sums(bX, bM, accX, accM);
                }
                __syncwarp();
                if (lds32(cntaddr) > TQ_CAP - 64) {
                    const int2 d = ising_drain<METHOD, PUSH, DN>(qaddr, cntaddr, own, a, tab, rep);
                    corrX += d.x; corrM += d.y;
                }
            }
        }
        if (PUSH && is_b) {
            // this chunk's ties are resolved (remote copies patched too), its stores are performed
            // system-wide, and the warp that completes the LAST boundary chunk of the launch tells both
            // neighbours that their halo of this colour is complete
            const int2 d = ising_drain<METHOD, PUSH, DN>(qaddr, cntaddr, own, a, tab, rep);
            corrX += d.x; corrM += d.y;
            __threadfence_system();
            __syncwarp();
            if (lane == 0 && atomicAdd(a.done, 1u) == (unsigned)a.nbchunks - 1u) {
                *a.done = 0;
                __threadfence_system();
                st_release_sys(a.sig_prev, a.sig_seq);
                st_release_sys(a.sig_next, a.sig_seq);
            }
        }
        cur = nxt;
        if (ORDERED) nxt = __shfl_sync(0xffffffffu, nx2, 0);
        else nxt += nwarps_grid * CH;
    }
    const int2 d = ising_drain<METHOD, PUSH, DN>(qaddr, cntaddr, own, a, tab, rep);
    if (MEASURE) {
        long long x = (long long)NNB * accM - 2ll * accX + corrX + d.x, mm = (long long)accM + corrM + d.y;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            x += __shfl_down_sync(0xffffffffu, x, o);
            mm += __shfl_down_sync(0xffffffffu, mm, o);
        }
        if (lane == 0) {
            if (x) atomicAdd(a.acc + 2 * rep, (unsigned long long)x);
            if (mm) atomicAdd(a.acc + 2 * rep + 1, (unsigned long long)mm);
            if (ORDERED && !PUSH && !BATCH && a.host_out) {
                // the last warp of the launch to get here reads the totals back (atomics at L2, ordered by the fences)
                // and stores them into pinned host memory: update -> calc_magne_sum -> calc_energy_sum needs no copy
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
}

template <int NNB, int METHOD, bool ORDERED, bool PUSH = false, bool MEASURE = false, bool BATCH = false>
__global__ void __launch_bounds__(256, PASS_MINB(NNB, PUSH))
ising_pass_kernel(const __grid_constant__ RingPassArgs a, const __grid_constant__ IsingTab tab)
{
    __shared__ uint4 tq[8][TQ_CAP][2];
    __shared__ uint32_t tq_cnt[8];
    ising_pass_body<NNB, METHOD, ORDERED, PUSH, MEASURE, BATCH>(a, tab, tq, tq_cnt);
}

// Slab mode, one launch per colour pass.  The first nb_blocks blocks of the grid (the ones the hardware
// starts first) take the boundary tickets of `ab` -- update + store into the neighbours' halos over NVLink,
// flag handshake -- and then join the other blocks on the interior tickets of `ai`, which is the plain body:
// the interior never waits for a neighbour, and the latency of the boundary work (system-scope fences, peer
// stores, flag waits) is hidden behind the interior blocks resident on the same SMs.
template <int NNB, int METHOD, bool MEASURE>
__device__ __forceinline__ void ising_slab_boundary(const RingPassArgs& ab, const IsingTab& tab, uint4 (*tq)[TQ_CAP][2], uint32_t* tq_cnt)
{
    ising_pass_body<NNB, METHOD, true, true, MEASURE, false>(ab, tab, tq, tq_cnt);
}

template <int NNB, int METHOD, bool MEASURE>
__global__ void __launch_bounds__(256, PASS_MINB(NNB, true))
ising_slab_kernel(const __grid_constant__ RingPassArgs ab, const __grid_constant__ RingPassArgs ai,
                  const __grid_constant__ IsingTab tab, const int nb_blocks)
{
    __shared__ uint4 tq[8][TQ_CAP][2];
    __shared__ uint32_t tq_cnt[8];
    if ((int)blockIdx.x < nb_blocks) ising_slab_boundary<NNB, METHOD, MEASURE>(ab, tab, tq, tq_cnt);
    ising_pass_body<NNB, METHOD, true, false, MEASURE, false>(ai, tab, tq, tq_cnt);
}

// Small lattices (at most a few vectors per resident thread): n_sweeps whole sweeps in ONE cooperative launch.
// At the reference's default sizes (app/ising2d_gpu_relaxation.f90:6-7: 1001 x 1000) a colour pass is ~1 us of
// work, and the per-sweep sequence of the plain path -- two pass launches, two to four halo launches, memsets -- is
// bound by launch latency (27 us per MCS).  Here every block runs colour pass -> grid barrier (-> halo refresh ->
// grid barrier when the halo is wide), twice per sweep, for all sweeps; static 32-vector chunks (one vector per
// lane) spread the lattice over the whole grid.  Sums: `series` != nullptr -> sweep i adds {X, sum s} to series[2 i], series[2 i + 1]
// (run_relaxation); else fuse_last -> the last sweep adds them to a[1].acc (update + calc_*_sum).
struct IsingCoopArgs {
    RingPassArgs a[2];          // per colour; draw = the first sweep's
    int64_t L, H, Nc, ptail;    // halo refresh (single GPU: the whole fold)
    int halo_fast;              // H <= L && ptail >= H: vector-granular halo + byte-granular tail, else all byte-granular
    int n_sweeps, fuse_last;
    unsigned long long* series;
    unsigned long long* host_out;   // fuse_last: the two sums are also stored here (pinned host memory) by the kernel itself
};

template <int NNB, int METHOD>
__global__ void __launch_bounds__(256, 2)
ising_coop_kernel(const __grid_constant__ IsingCoopArgs s, const __grid_constant__ IsingTab tab)
{
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ uint4 tq[8][TQ_CAP][2];
    __shared__ uint32_t tq_cnt[8];
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gthreads = (int64_t)gridDim.x * blockDim.x;
    // update_norishiro (src/ising3d_gpu_m.f90:102-122) for one colour array, by the whole grid
    auto halo = [&](uint4* vec) {
        if (s.halo_fast) {
            for (int64_t v = gtid; v < 2 * s.H; v += gthreads) ring_halo_fast_item(vec, s.L, s.H, s.Nc, v);
            for (int64_t t = gtid; t < (s.L - s.ptail) * 16; t += gthreads)
                ring_halo_generic_item(reinterpret_cast<uint8_t*>(vec), s.L, s.H, s.Nc, s.ptail, 2 * s.H, t);
        } else {
            for (int64_t t = gtid; t < (2 * s.H + s.L - s.ptail) * 16; t += gthreads)
                ring_halo_generic_item(reinterpret_cast<uint8_t*>(vec), s.L, s.H, s.Nc, s.ptail, 0, t);
        }
    };
    // fuse_last: the accumulators are cleared here (the pass that adds to them comes after at least one grid barrier)
    if (s.fuse_last && !s.series && gtid == 0) { s.a[1].acc[0] = 0ull; s.a[1].acc[1] = 0ull; }
    for (int i = 0; i < s.n_sweeps; ++i) {
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
            RingPassArgs b = s.a[c];
            b.draw += (uint64_t)i;
            const bool meas = c == 1 && (s.series != nullptr || (s.fuse_last && i == s.n_sweeps - 1));
            if (meas) {
                if (s.series) b.acc = s.series + 2 * (size_t)i;
                ising_pass_body<NNB, METHOD, false, false, true, false, 32>(b, tab, tq, tq_cnt);
            } else {
                ising_pass_body<NNB, METHOD, false, false, false, false, 32>(b, tab, tq, tq_cnt);
            }
            grid.sync();
            if (s.a[0].wrap_mode == 0) {   // wide halos (3D): refresh after every pass, one more barrier
                halo(b.own);
                grid.sync();
            }
        }
    }
    // narrow halos (2D): the passes rebuilt wrapped neighbour positions on the fly (ring_load_wrapped); the halo vectors
    // and tail lanes every other kernel reads are refreshed once, here
    if (s.a[0].wrap_mode != 0) { halo(s.a[0].own); halo(s.a[1].own); }
    // (after the last barrier every block's atomics have landed)
    if (s.fuse_last && !s.series && s.host_out && gtid == 0) {
        s.host_out[0] = __ldcg(s.a[1].acc);
        s.host_out[1] = __ldcg(s.a[1].acc + 1);
    }
}

// ---------------------------------------------------------------------------
// TMA-staged colour pass (single-GPU launches; experiment, B200MC_TUNE bit 7).
// ising_pass_kernel is bound by the L1TEX request path: eight 512-byte requests per warp-step (a coalesced
// LDG.128 is 4 wavefronts).  Here a ninth warp of the block feeds the copy engine: per tile of 256 vectors
// one `cp.async.bulk` per input stream (1-D, no tensor map: every stream is a contiguous run of vectors in
// the folded layout; the x- / x+ windows overlap in all but one vector and are fetched once, 257 vectors) into
// a ring of TMA_STAGES shared-memory stages, completion on a "full" mbarrier per stage; the eight consumer
// warps read their 32 vectors back with LDS.128, update, store with STG and release the stage on an "empty"
// mbarrier.  Tiles are drawn from the same interleaved ticket counters as the plain kernel.
// ---------------------------------------------------------------------------
#define TMA_STAGES 2
#define TMA_TILE 256                 // vectors per tile (8 consumer warps x 32)
#define TMA_SLOT (16 * (TMA_TILE + 2))   // bytes per stream slot (257 vectors used by the x window)

__device__ __forceinline__ void mbar_init(uint32_t mb, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mb, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t mb)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mb) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t mb, uint32_t phase)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(mb), "r"(phase)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t mb, uint32_t phase)
{
    int spins = 0;
    while (!mbar_try_wait(mb, phase)) { if (++spins > (1 << 22)) __trap(); }   // a lost arrival aborts instead of hanging the GPU
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mb)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(mb)
                 : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}

template <int NNB, int METHOD, bool MEASURE>
__global__ void __launch_bounds__(288)
ising_pass_tma_kernel(const __grid_constant__ RingPassArgs a, const __grid_constant__ IsingTab tab)
{
    constexpr int NSTREAM = NNB;                   // own, x window, y+, y- (, z+, z-)
    constexpr uint32_t STAGE_BYTES = NSTREAM * TMA_SLOT;
    constexpr uint32_t TX_BYTES = (NSTREAM - 1) * (16u * TMA_TILE) + 16u * (TMA_TILE + 1);
    extern __shared__ __align__(128) uint8_t tma_smem[];
    __shared__ uint4 tq[8][64][2];   // drained when more than 32 records are parked, checked every step
    __shared__ uint32_t tq_cnt[8];
    __shared__ __align__(8) unsigned long long mb_full[TMA_STAGES], mb_empty[TMA_STAGES];
    __shared__ int tile_base[TMA_STAGES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t buf0 = (uint32_t)__cvta_generic_to_shared(tma_smem);
    const uint32_t full0 = (uint32_t)__cvta_generic_to_shared(&mb_full[0]);
    const uint32_t empty0 = (uint32_t)__cvta_generic_to_shared(&mb_empty[0]);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < TMA_STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp < 8 && lane == 0) tq_cnt[warp] = 0;
    __syncthreads();
    uint4* own = a.own + a.H;
    const int nvec = (int)a.nvec;
    if (warp == 8) {
        // ---- producer: one lane feeds the copy engine ----
        if (lane == 0) {
            const uint4* src[NSTREAM];
            src[0] = own;
            src[1] = a.oth + (a.H + a.off[0]);     // 257-vector x window: x- = [v], x+ = [v + 1]
#pragma unroll
            for (int j = 2; j < NSTREAM; ++j) src[j] = a.oth + (a.H + a.off[j]);
            unsigned int* tk = a.ticket + (blockIdx.x % TK_NCNT) * 64;
            const int tk_base = (blockIdx.x % TK_NCNT) * TMA_TILE;
            int base_next = (int)atomicAdd(tk, (unsigned)TMA_TILE) * TK_NCNT + tk_base;
            for (uint32_t it = 0;; ++it) {
                const uint32_t st = it % TMA_STAGES, round = it / TMA_STAGES;
                const int base = base_next;
                if (base < nvec) base_next = (int)atomicAdd(tk, (unsigned)TMA_TILE) * TK_NCNT + tk_base;   // one tile ahead: its latency hides behind this tile
                if (round > 0) mbar_wait(empty0 + 8 * st, (round - 1) & 1u);   // the consumers have released this stage
                tile_base[st] = base < nvec ? base : -1;
                const uint32_t mb = full0 + 8 * st, dst = buf0 + st * STAGE_BYTES;
                if (base + TMA_TILE <= nvec) {
                    mbar_expect_tx(mb, TX_BYTES);
                    bulk_g2s(dst, src[0] + base, 16u * TMA_TILE, mb);
                    bulk_g2s(dst + TMA_SLOT, src[1] + base, 16u * (TMA_TILE + 1), mb);
#pragma unroll
                    for (int j = 2; j < NSTREAM; ++j) bulk_g2s(dst + j * TMA_SLOT, src[j] + base, 16u * TMA_TILE, mb);
                } else {
                    mbar_arrive(mb);   // partial last tile (plain loads) or the end marker
                }
                if (base >= nvec) break;
            }
        }
        return;
    }
    // ---- consumers ----
    uint32_t qaddr = (uint32_t)__cvta_generic_to_shared(&tq[warp][0][0]);
    uint32_t cntaddr = (uint32_t)__cvta_generic_to_shared(&tq_cnt[warp]);
    const uint64_t pol = l2_policy_evict_first();
    const uint32_t cx0 = (uint32_t)a.p0, cz = (uint32_t)a.draw;
    const uint32_t cw = (uint32_t)((a.draw >> 32) & 0xFFFFu) | (a.colour << 16);
    uint32_t accX = 0, accM = 0, bX = 0, bM = 0;
    int corrX = 0, corrM = 0;
    constexpr int DN = MEASURE ? NNB : 0;
    for (uint32_t it = 0;; ++it) {
        const uint32_t st = it % TMA_STAGES, round = it / TMA_STAGES;
        mbar_wait(full0 + 8 * st, round & 1u);
        const int base = tile_base[st];
        if (base < 0) break;
        const int v = base + 32 * warp + lane;
        if (base + TMA_TILE <= nvec) {
            const uint32_t bufs = buf0 + st * STAGE_BYTES + 16u * (32 * warp + lane);
            const uint4 o = lds128(bufs);
            uint4 nb[NNB];
            nb[0] = lds128(bufs + TMA_SLOT);
            nb[1] = lds128(bufs + TMA_SLOT + 16u);
#pragma unroll
            for (int j = 2; j < NNB; ++j) nb[j] = lds128(bufs + j * TMA_SLOT);
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * st);   // this warp's 32 vectors are in registers
            ising_core<NNB, METHOD, false, MEASURE>(v, own + v, o, nb, cx0 + (uint32_t)v, cz, cw, a, tab, pol, qaddr, cntaddr, false,
                                                    bX, bM);
            if (MEASURE) ising_fold_sums(bX, bM, accX, accM);
        } else {
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * st);
            if (v < nvec) {
                const uint4* qu[NNB];
#pragma unroll
                for (int j = 0; j < NNB; ++j) qu[j] = a.oth + (a.H + a.off[j]) + v;
                ising_vec<NNB, METHOD, false, MEASURE, false>(v, own + v, qu, cx0 + (uint32_t)v, cz, cw, a, tab, pol, qaddr, cntaddr,
                                                              false, bX, bM);
                if (MEASURE) ising_fold_sums(bX, bM, accX, accM);
            }
        }
        __syncwarp();
        if (lds32(cntaddr) > 32) {
            const int2 d = ising_drain<METHOD, false, DN>(qaddr, cntaddr, own, a, tab);
            corrX += d.x; corrM += d.y;
        }
    }
    const int2 d = ising_drain<METHOD, false, DN>(qaddr, cntaddr, own, a, tab);
    if (MEASURE) {
        long long x = (long long)NNB * accM - 2ll * accX + corrX + d.x, mm = (long long)accM + corrM + d.y;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            x += __shfl_down_sync(0xffffffffu, x, o);
            mm += __shfl_down_sync(0xffffffffu, mm, o);
        }
        if (lane == 0) {
            if (x) atomicAdd(a.acc, (unsigned long long)x);
            if (mm) atomicAdd(a.acc + 1, (unsigned long long)mm);
        }
    }
}

// ---------------------------------------------------------------------------
// Reference-stream pass: uniforms come from a device array in the reference's
// own index order (randoms(idx), idx = i+1) and are compared as real64 against
// the real64 table, exactly like update_sub.  Used for parity tests with
// arbitrary uniforms and for the cuRAND-stream mode.  Not the fast path.
// wtab[s*8 + S].
// ---------------------------------------------------------------------------
struct IsingTabF64 { double w[16]; };

template <int NNB, int METHOD>
__global__ void __launch_bounds__(256)
ising_pass_randoms_kernel(const __grid_constant__ RingPassArgs a, const __grid_constant__ IsingTabF64 tab,
                          const double* __restrict__ randoms, int64_t L, int64_t Nc)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < a.nvec; v += stride) {
        const int64_t q = v + a.H;
        uint4 o = a.own[q];
        uint4 S = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int j = 0; j < NNB; ++j) {
            const uint4 n = a.oth[q + a.off[j]];
            S.x += n.x; S.y += n.y; S.z += n.z; S.w += n.w;
        }
        uint32_t ow[4] = {o.x, o.y, o.z, o.w};
        const uint32_t sw[4] = {S.x, S.y, S.z, S.w};
        const int64_t p = a.p0 + v;
#pragma unroll
        for (int b = 0; b < 16; ++b) {
            const int64_t k = (int64_t)b * L + p;
            if (k >= Nc) continue;
            const int64_t i = 2 * k + a.colour;
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
        a.own[q] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
}

// ---------------------------------------------------------------------------
// Fused energy + magnetisation (reference: two OpenACC reductions that re-read
// the lattice, src/ising3d_gpu_m.f90:239-276, src/ising2d_gpu_m.f90:198-228).
// One pass over the colour-1 vectors: every bond has exactly one colour-1 end,
// so  X = sum over colour-1 sites of #(neighbours with a different spin)
// counts every anti-aligned bond once and  E = -(nnb/2) N + 2 X  exactly;
// sum(s) is accumulated for both colours.  acc[0] += X, acc[1] += sum(s).
// ---------------------------------------------------------------------------
template <int NNB>
__global__ void __launch_bounds__(256)
ising_measure_kernel(const uint4* __restrict__ c0, const uint4* __restrict__ c1, int64_t nvec,
                     int64_t H, int64_t p0, const int64_t* offs /* colour-1 offsets */, int64_t L,
                     int64_t Nc, int64_t ptail, unsigned long long* acc, int64_t rstride = 0)
{
    c0 += (size_t)blockIdx.y * (size_t)rstride;   // sample of the batch
    c1 += (size_t)blockIdx.y * (size_t)rstride;
    acc += 2 * blockIdx.y;
    __shared__ int64_t off[6];
    if (threadIdx.x < 6) off[threadIdx.x] = offs[threadIdx.x];
    __syncthreads();
    long long part[2] = {0, 0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        const int64_t q = v + H;
        const int64_t p = p0 + v;
        uint4 o = c1[q];
        uint4 a0 = c0[q];  // colour-0 vector at the same position: owned sites for sum(s)
        uint4 X = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int j = 0; j < NNB; ++j) {
            const uint4 n = c0[q + off[j]];
            X.x += n.x ^ o.x; X.y += n.y ^ o.y; X.z += n.z ^ o.z; X.w += n.w ^ o.w;
        }
        if (p >= ptail) {
            // mask lanes that hold no site (k = b L + p >= Nc)
            uint32_t keep[4];
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                keep[w] = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if ((int64_t)(4 * w + j) * L + p < Nc) keep[w] |= 0xFFu << (8 * j);
            }
            X.x &= keep[0]; X.y &= keep[1]; X.z &= keep[2]; X.w &= keep[3];
            o.x &= keep[0]; o.y &= keep[1]; o.z &= keep[2]; o.w &= keep[3];
            a0.x &= keep[0]; a0.y &= keep[1]; a0.z &= keep[2]; a0.w &= keep[3];
        }
        // byte sums with a full-rate multiply (IDP.4A is a slow pipe on sm_100): bytes of the word sums are <= 24 / <= 8
        const uint32_t x = (((X.x + X.y) + (X.z + X.w)) * 0x01010101u) >> 24;
        const uint32_t m = ((((o.x + a0.x) + (o.y + a0.y)) + ((o.z + a0.z) + (o.w + a0.w))) * 0x01010101u) >> 24;
        part[0] += x;
        part[1] += m;
    }
    block_atomic_add<2>(acc, part);
}

// set_random_spin (reference: set_random_spin_sub, src/ising3d_gpu_m.f90:91-100):
// s = (u < 0.5) with u = (U+1) 2^-32, U = R[lane & 3] of
// philox(ctr(p, draw, colour, lane >> 2), (seed, TAG_INIT)).
static __global__ void __launch_bounds__(256)
ring_random_bits_kernel(uint4* own, int64_t nvec, int64_t H, int64_t p0, uint32_t seed,
                        uint64_t draw, uint32_t colour, int64_t rstride = 0)
{
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvec) return;
    own += (size_t)blockIdx.y * (size_t)rstride;   // sample of the batch: counters (position | sample << 32, ...)
    uint32_t w[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const uint4 r = philox4x32_10(mk_ctr((uint64_t)(p0 + v) | ((uint64_t)blockIdx.y << 32), draw, colour, g), make_uint2(seed, TAG_INIT));
        // (U+1) 2^-32 < 0.5  <=>  U < 2^31 - 1
        w[g] = (r.x < 0x7FFFFFFFu ? 1u : 0u) | (r.y < 0x7FFFFFFFu ? 0x100u : 0u) |
               (r.z < 0x7FFFFFFFu ? 0x10000u : 0u) | (r.w < 0x7FFFFFFFu ? 0x1000000u : 0u);
    }
    own[v + H] = make_uint4(w[0], w[1], w[2], w[3]);
}
