// Bit-packed (multi-spin coded) helical Ising 2D / 3D: one BIT per site, Metropolis.  The second storage format the north
// star names ("int8 or bit-packed multi-spin-coded Ising lattices"); same reference path as ising.cu:
// update_sub + curandGenerate (src/ising3d_gpu_m.f90:174-206, src/ising2d_gpu_m.f90:138-162), update_norishiro_sub
// (:111-122), the two OpenACC reductions (:239-276), set_random_spin (:84-100), spins() (:232-236).
//
// Layout.  The helical lattice is a ring of N sites, colour = parity of the 0-based linear index i, colour-site index
// k = i >> 1 (ring.cuh).  Each colour ring of Nc = N / 2 sites is folded into 128 bit-lanes of L = Nc / 128 positions
// (Nc must be a multiple of 128): site k -> lane k / L, position k % L, stored as bit (lane & 31) of word (lane >> 5) of the
// 128-bit vector `position`.  Every neighbour offset of the ring is a whole-vector offset, so a thread that owns one
// vector (128 sites) reads 1 + nnb aligned 128-bit vectors and all arithmetic is bit-parallel, 32 sites per instruction.
// A position beyond [0, L) belongs to the next / previous lane: the H = max|offset| halo vectors on each side are the
// vectors of the other end of the fold rotated by one bit (bits_halo_kernel) -- 3/8 byte of HBM traffic per attempted
// flip against 3 bytes for the int8 layout.
//
// Update.  Bit-sliced count of the aligned neighbours k' (two full adders + a 2-bit add: 10 LOP3 per 32 sites in 3D);
// sites with k' <= nnb/2 flip (dE <= 0), the others -- classes k' = 4, 5, 6 (3D) / 3, 4 (2D) -- flip iff U < thr[k'],
// thr = floor(w 2^32), exactly the reference's `randoms(idx) > ws(...)` test on u = (U + 1) 2^-32.  The comparison is
// bit-serial on the 8 leading bits: plane j of a vector's uniforms is one Philox block (4 words = bit j of the 128
// uniforms), compared with bit j of the site's threshold; a site is decided at the first plane where the two differ.  The
// 2^-8 of the non-trivial sites that are still undecided after 8 planes take the remaining 24 bits of their uniform from a
// per-site Philox word (a scalar tail, about 0.35 sites per vector near T_c; with planes alone a warp needed about 13 of
// them before its 4096 sites were all decided).  Exact at the full 32 bits (U == thr: not accepted).
//
// RNG contract (CPU restatement: oracle/rng_contract.c, orc_isingbits_uniforms): vector p, colour c, sweep `draw`:
//   R_j = philox(ctr(p, draw, c, sub = j), (seed, TAG_ISNB)), j = 0..7;  S = philox(ctr(p, draw, c, sub = 32 + (lane >> 2)), same key)
//   site lane = 32 w + b:  U = sum_j ((R_j[w] >> b) & 1) << (31 - j)  |  S[lane & 3] & 0xFFFFFF
// set_random_spin: R = philox(ctr(p, draw, c, 0), (seed, TAG_INIB)), U = ((R[w] >> b) & 1) << 31 (the reference only tests
// u < 0.5: spin = 1 - bit).
#include <math.h>
#include <stdlib.h>
#include <new>
#include <vector>
#include "../../include/b200mc.h"
#include "common.cuh"
#include "ising_kernels.cuh"   // philox_rk
#include "ring.cuh"            // dist_* (NCCL resolved at run time)

namespace {

#define TAG_ISNB 0x49534E42u /* "ISNB": accept uniforms of the bit-packed Ising models */
#define BITS_PLANES 8                                   /* leading bits of a uniform that come from bit planes */
#define BITS_LOW_MASK ((1u << (32 - BITS_PLANES)) - 1u)
#define TAG_INIB 0x494E4942u /* "INIB": their set_random_spin */

struct BitsPassArgs {
    uint4* own;          // colour being updated, index 0 = position 0
    const uint4* oth;
    int64_t L;           // vectors this launch covers (the rank's share of every lane in slab mode, or a part of it)
    int64_t pbeg;        // first local vector of this launch (slab mode: boundary / interior launches)
    int64_t p0;          // global position of local vector 0 (slab mode; 0 otherwise): the RNG counters use global positions
    int64_t off[6];      // neighbour vector offsets: x-, x+, y+, y-, z+, z-
    uint32_t thr[3];     // thresholds of the non-trivial classes: 3D k' = 4, 5, 6; 2D k' = 3, 4
    uint32_t always[3];  // 0 / ~0: the class accepts every proposal (thr = 2^32: beta = 0)
    uint32_t colour;
    uint64_t draw;
    uint32_t rk0[10];
    unsigned long long* acc;   // measure kernel: acc[0] += sum k' over the colour-1 sites, acc[1] += sum s over both colours
};

__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | (a & c) | (b & c); }

// bit-sliced number of aligned neighbours k' = c0 + 2 c1 + 4 c2 of the 32 sites of a word
template <int NNB>
__device__ __forceinline__ void bits_count(uint32_t s, const uint32_t (&n)[6], uint32_t& c0, uint32_t& c1, uint32_t& c2)
{
    const uint32_t t = ~s;   // neighbour j is aligned iff n_j ^ t
    if (NNB == 6) {
        // (n0, n1, n2) and (n3, n4, n5) through one full adder each; the aligned versions of a group's sum / carry are
        // x ^ t and maj ^ t (three inputs: parity flips, majority is self-dual)
        const uint32_t x = n[0] ^ n[1] ^ n[2], cx = maj3(n[0], n[1], n[2]) ^ t;
        const uint32_t y = n[3] ^ n[4] ^ n[5], cy = maj3(n[3], n[4], n[5]) ^ t;
        c0 = x ^ y;
        const uint32_t carry = (x ^ t) & (y ^ t);
        c1 = cx ^ cy ^ carry;
        c2 = maj3(cx, cy, carry);
    } else {
        const uint32_t x = n[0] ^ n[1] ^ n[2], cx = maj3(n[0], n[1], n[2]) ^ t;
        c0 = x ^ n[3];
        const uint32_t carry = (x ^ t) & (n[3] ^ t);
        c1 = cx ^ carry;
        c2 = cx & carry;
    }
}

// MEASURE (second colour pass of a sweep when the caller measures every MCS): the pass also accumulates acc[0] += sum of the
// aligned-neighbour counts of the NEW colour-1 spins (k' of a flipped site becomes nnb - k') and acc[1] += sum s over both
// colours (the colour-0 vector at the same position is the x- neighbour vector), like bits_measure_kernel.
template <int NNB, int MINB, bool MEASURE>
__global__ void __launch_bounds__(256, MINB)
bits_pass_kernel(const __grid_constant__ BitsPassArgs a)
{
    long long part[2] = {0, 0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = a.pbeg + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < a.pbeg + a.L; p += stride) {
        const uint4 o = a.own[p];
        uint4 nb[NNB];
#pragma unroll
        for (int j = 0; j < NNB; ++j) nb[j] = ld_other(a.oth + p + a.off[j]);
        const uint32_t s[4] = {o.x, o.y, o.z, o.w};
        uint32_t sel0[4], sel1[4], und[4], lt[4], cc2[4];
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            uint32_t n[6] = {0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
            for (int j = 0; j < NNB; ++j) n[j] = w == 0 ? nb[j].x : w == 1 ? nb[j].y : w == 2 ? nb[j].z : nb[j].w;
            uint32_t c0, c1, c2;
            bits_count<NNB>(s[w], n, c0, c1, c2);
            // class selectors: threshold = sel1 ? thr[2] : (sel0 ? thr[1] : thr[0]) and the non-trivial sites
            uint32_t nt;
            if (NNB == 6) { nt = c2; sel0[w] = c0; sel1[w] = c1; cc2[w] = c2; }  // k' = 4: (c1, c0) = (0, 0); 5: (0, 1); 6: (1, 0)
            else { nt = c2 | (c1 & c0); sel0[w] = c2; sel1[w] = c1; cc2[w] = c0; }   // k' = 3: thr[0]; 4: thr[1]  (sel1 / cc2 only carry c1 / c0 for MEASURE)
            const uint32_t f = (sel0[w] & a.always[1]) | (~sel0[w] & a.always[0]);
            const uint32_t A = NNB == 6 ? ((sel1[w] & a.always[2]) | (~sel1[w] & f)) : f;
            und[w] = nt & ~A;
            lt[w] = ~und[w];      // trivial classes (dE <= 0) and always-accept classes flip
        }
        // bit-serial U < thr on the BITS_PLANES leading planes, most significant first
#pragma unroll
        for (int j = 0; j < BITS_PLANES; ++j) {
            if (j >= 3 && !(und[0] | und[1] | und[2] | und[3])) break;
            const uint4 R = philox_rk<TAG_ISNB>(mk_ctr((uint64_t)(a.p0 + p), a.draw, a.colour, (uint32_t)j), a.rk0);
            const uint32_t r[4] = {R.x, R.y, R.z, R.w};
            const uint32_t P0 = 0u - ((a.thr[0] >> (31 - j)) & 1u), P1 = 0u - ((a.thr[1] >> (31 - j)) & 1u),
                           P2 = 0u - ((a.thr[2] >> (31 - j)) & 1u);
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const uint32_t f = (sel0[w] & P1) | (~sel0[w] & P0);
                const uint32_t T = NNB == 6 ? ((sel1[w] & P2) | (~sel1[w] & f)) : f;
                lt[w] |= und[w] & ~r[w] & T;       // uniform bit 0, threshold bit 1: U < thr
                und[w] &= ~(r[w] ^ T);             // still equal: undecided
            }
        }
        // the sites whose leading planes equal their threshold's (2^-BITS_PLANES of the non-trivial ones): the low
        // 32 - BITS_PLANES bits of the uniform come from a per-site Philox word
        // (ONE loop over the four words: a warp runs max-over-lanes iterations, about 2 near T_c; a loop per word ran about
        // one iteration for each of the four words on top of every plane it replaced)
        while (und[0] | und[1] | und[2] | und[3]) {
            const int w = und[0] ? 0 : und[1] ? 1 : und[2] ? 2 : 3;
            const uint32_t m = w == 0 ? und[0] : w == 1 ? und[1] : w == 2 ? und[2] : und[3];
            const uint32_t s0 = w == 0 ? sel0[0] : w == 1 ? sel0[1] : w == 2 ? sel0[2] : sel0[3];
            const uint32_t s1 = w == 0 ? sel1[0] : w == 1 ? sel1[1] : w == 2 ? sel1[2] : sel1[3];
            const int b = __ffs(m) - 1;
            const uint32_t bit = 1u << b;
            const int lane = 32 * w + b;
            const uint4 R = philox_rk<TAG_ISNB>(mk_ctr((uint64_t)(a.p0 + p), a.draw, a.colour, (uint32_t)(32 + (lane >> 2))), a.rk0);
            const uint32_t S = ((lane & 3) == 0 ? R.x : (lane & 3) == 1 ? R.y : (lane & 3) == 2 ? R.z : R.w) & BITS_LOW_MASK;
            const uint32_t cls = NNB == 6 ? ((s1 & bit) ? 2u : ((s0 & bit) ? 1u : 0u)) : ((s0 & bit) ? 1u : 0u);
            const uint32_t t = (cls == 0 ? a.thr[0] : cls == 1 ? a.thr[1] : a.thr[2]) & BITS_LOW_MASK;
            const uint32_t acc = S < t ? bit : 0u;
#pragma unroll
            for (int ww = 0; ww < 4; ++ww)
                if (ww == w) { und[ww] &= ~bit; lt[ww] |= acc; }
        }
        a.own[p] = make_uint4(s[0] ^ lt[0], s[1] ^ lt[1], s[2] ^ lt[2], s[3] ^ lt[3]);
        if (MEASURE) {
            const uint32_t z[4] = {nb[0].x, nb[0].y, nb[0].z, nb[0].w};   // colour 0 at the same position (offset 0 for colour 1)
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                // count bits (b0, b1, b2) of k': 3D (sel0, sel1, cc2); 2D (cc2, sel1, sel0)
                const uint32_t b0 = NNB == 6 ? sel0[w] : cc2[w], b1 = sel1[w], b2 = NNB == 6 ? cc2[w] : sel0[w], f = lt[w];
                const int keep = __popc(b0 & ~f) + 2 * __popc(b1 & ~f) + 4 * __popc(b2 & ~f);
                const int flip = __popc(b0 & f) + 2 * __popc(b1 & f) + 4 * __popc(b2 & f);
                part[0] += keep + NNB * __popc(f) - flip;
                part[1] += __popc(s[w] ^ f) + __popc(z[w]);
            }
        }
    }
    if (MEASURE) block_atomic_add<2>(a.acc, part);
}

// halo refresh of one colour array (index 0 = position -H): position -H + v <- position L - H + v rotated up by one lane,
// position L + v <- position v rotated down by one lane (v = 0 .. H-1); the counterpart of update_norishiro_sub
__global__ void __launch_bounds__(256)
bits_halo_kernel(uint4* vec, int64_t L, int64_t H)
{
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= 2 * H) return;
    if (v < H) {
        const uint4 s = vec[L + v];                 // position L - H + v
        uint4 o;                                     // lane j <- lane j - 1, lane 0 <- lane 127
        o.x = __funnelshift_l(s.w, s.x, 1);
        o.y = __funnelshift_l(s.x, s.y, 1);
        o.z = __funnelshift_l(s.y, s.z, 1);
        o.w = __funnelshift_l(s.z, s.w, 1);
        vec[v] = o;
    } else {
        const int64_t t = v - H;
        const uint4 s = vec[H + t];                 // position t
        uint4 o;                                     // lane j <- lane j + 1, lane 127 <- lane 0
        o.x = __funnelshift_r(s.x, s.y, 1);
        o.y = __funnelshift_r(s.y, s.z, 1);
        o.z = __funnelshift_r(s.z, s.w, 1);
        o.w = __funnelshift_r(s.w, s.x, 1);
        vec[H + L + t] = o;
    }
}

// slab mode: the halo blocks arrive from the neighbour ranks unrotated; the two ranks at the ends of the fold rotate what
// they received by one bit-lane (dir = +1: lane j <- lane j-1, lane 0 <- lane 127; dir = -1: the other way)
__global__ void __launch_bounds__(256)
bits_rotate_kernel(uint4* v, int64_t n, int dir)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 s = v[i];
    uint4 o;
    if (dir > 0) {
        o.x = __funnelshift_l(s.w, s.x, 1); o.y = __funnelshift_l(s.x, s.y, 1);
        o.z = __funnelshift_l(s.y, s.z, 1); o.w = __funnelshift_l(s.z, s.w, 1);
    } else {
        o.x = __funnelshift_r(s.x, s.y, 1); o.y = __funnelshift_r(s.y, s.z, 1);
        o.z = __funnelshift_r(s.z, s.w, 1); o.w = __funnelshift_r(s.w, s.x, 1);
    }
    v[i] = o;
}

// E and M: one pass over the colour-1 vectors.  Every bond has exactly one colour-1 end, so the number of unequal bonds is
// X = nnb Nc - sum over the colour-1 sites of k'; sum s = popcounts of both colours.
template <int NNB>
__global__ void __launch_bounds__(256)
bits_measure_kernel(const __grid_constant__ BitsPassArgs a, const uint4* __restrict__ c0vec)
{
    long long part[2] = {0, 0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < a.L; p += stride) {
        const uint4 o = a.own[p];
        uint4 nb[NNB];
#pragma unroll
        for (int j = 0; j < NNB; ++j) nb[j] = a.oth[p + a.off[j]];
        const uint4 z = c0vec[p];
        const uint32_t s[4] = {o.x, o.y, o.z, o.w}, zw[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            uint32_t n[6] = {0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
            for (int j = 0; j < NNB; ++j) n[j] = w == 0 ? nb[j].x : w == 1 ? nb[j].y : w == 2 ? nb[j].z : nb[j].w;
            uint32_t c0, c1, c2;
            bits_count<NNB>(s[w], n, c0, c1, c2);
            part[0] += __popc(c0) + 2 * __popc(c1) + 4 * __popc(c2);
            part[1] += __popc(s[w]) + __popc(zw[w]);
        }
    }
    block_atomic_add<2>(a.acc, part);
}

__global__ void __launch_bounds__(256)
bits_random_kernel(uint4* own, int64_t L, int64_t p0, uint32_t seed, uint64_t draw, uint32_t colour)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= L) return;
    const uint4 R = philox4x32_10(mk_ctr((uint64_t)(p0 + p), draw, colour, 0u), make_uint2(seed, TAG_INIB));
    own[p] = make_uint4(~R.x, ~R.y, ~R.z, ~R.w);   // spin = 1 iff u < 1/2 iff the bit is 0
}

// spins() in the reference layout spins(1-P : N+P), halo cells included: int32 0/1 (3D) or -1/+1 (2D, pm1)
__global__ void __launch_bounds__(256)
bits_export_kernel(const uint4* c0, const uint4* c1, int64_t N, int64_t L, int64_t P, int pm1, int64_t p0, int64_t Lloc, int32_t* out)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N + 2 * P) return;
    int64_t i = idx - P;
    if (i < 0) i += N;
    if (i >= N) i -= N;
    const int64_t k = i >> 1, lane = k / L, pos = k - lane * L;
    if (pos < p0 || pos >= p0 + Lloc) { out[idx] = INT32_MIN; return; }   // owned by another rank (the host takes the maximum over the ranks)
    const uint32_t* w = reinterpret_cast<const uint32_t*>(((i & 1) ? c1 : c0) + (pos - p0));
    const int32_t b = (int32_t)((w[lane >> 5] >> (lane & 31)) & 1u);
    out[idx] = pm1 ? 2 * b - 1 : b;
}
// inverse: one thread builds one 32-bit word (32 sites, L apart in k)
__global__ void __launch_bounds__(256)
bits_import_kernel(uint4* c0, uint4* c1, int64_t L, int64_t P, int pm1, int64_t p0, int64_t Lloc, const int32_t* in)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 8 * Lloc) return;
    const int colour = (int)(t / (4 * Lloc));
    const int64_t r = t - (int64_t)colour * 4 * Lloc, lpos = r >> 2, pos = p0 + lpos;
    const int w = (int)(r & 3);
    uint32_t word = 0;
    for (int b = 0; b < 32; ++b) {
        const int64_t k = (int64_t)(32 * w + b) * L + pos;
        const int32_t v = in[2 * k + colour + P];
        const int32_t bit = pm1 ? (v + 1) >> 1 : v;
        word |= (uint32_t)(bit & 1) << b;
    }
    reinterpret_cast<uint32_t*>((colour ? c1 : c0) + lpos)[w] = word;
}

struct Bits {
    int ndim, nnb;
    int64_t nx, ny, nz, N, Nc, L, H, P;
    int64_t p0, Lloc;    // slab mode: this rank owns positions [p0, p0 + Lloc) of every bit-lane (single GPU: 0, L)
    int rank, nranks;
    void* comm;          // ncclComm_t when nranks > 1
    cudaStream_t aux, hi; // slab mode: the interior launch of a colour pass (low priority) beside the boundary launches + exchange (high priority)
    cudaEvent_t ev_main, ev_aux, ev_hi;
    int64_t off[2][6];
    uint4* vec[2];       // [colour]: L + 2H vectors, position p at index p + H
    int32_t* stage;
    unsigned long long* d_acc;
    cudaStream_t stream;
    double beta;
    double w[16];        // w[s * 8 + S]: the reference's table
    uint64_t thr[8];     // floor(w 2^32) per k'
    uint32_t seed;
    uint64_t draw;
    int short_blocks;    // one vector per thread (grid = all blocks) instead of a resident grid-stride grid (B200MC_BITS_SHORT, A/B)
    int grid, minb;      // minb: resident blocks per SM the 3D pass is compiled for (3: 85 registers, 4: 64 with a few spills)
    bool obs_valid;
    bool want_fused, fused_pending, swept_since_measure;   // fused measurement, as for the int8 handles (ising.cu)
    int64_t obs_e, obs_m;
    bool timing;
    std::vector<cudaEvent_t> evs;
    size_t ev_used;
};

int build_tables(Bits* m)
{
    const double beta = m->beta;
    for (int i = 0; i < 16; ++i) m->w[i] = 0.0;
    if (m->ndim == 3) {
        // update_ws_ising3d_gpu, src/ising3d_gpu_m.f90:138-172 (same loop nest and expression order)
        static const int32_t spin_map[2] = {-1, 1};
        int64_t et[4][2];
        for (int i1 = 0; i1 <= 1; ++i1)
            for (int i2 = 0; i2 <= 1; ++i2)
                for (int i3 = 0; i3 <= 1; ++i3) {
                    const int s1 = i1 + i2 + i3;
                    const int32_t sum = spin_map[i1] + spin_map[i2] + spin_map[i3];
                    et[s1][0] = -spin_map[0] * sum;
                    et[s1][1] = -spin_map[1] * sum;
                }
        for (int s1 = 0; s1 <= 3; ++s1)
            for (int s2 = 0; s2 <= 3; ++s2) {
                const int64_t e1 = et[s1][0] + et[s2][0];
                const int64_t e2 = et[s1][1] + et[s2][1];
                m->w[0 * 8 + s1 + s2] = fmin(1.0, exp(-beta * (double)(e2 - e1)));
                m->w[1 * 8 + s1 + s2] = fmin(1.0, exp(-beta * (double)(e1 - e2)));
            }
    } else {
        // update_exparr_ising2d_gpu, src/ising2d_gpu_m.f90:122-131; calc_delta_energy :195 with sigma = 2s-1, sum(sigma_nb) = 2S-4
        double exparr[17];
        for (int i = 0; i < 17; ++i) exparr[i] = 1.0;
        for (int diff = 1; diff <= 8; ++diff) exparr[diff + 8] = exp(-beta * diff);
        for (int s = 0; s < 2; ++s)
            for (int S = 0; S <= 4; ++S) m->w[s * 8 + S] = exparr[2 * (2 * s - 1) * (2 * S - 4) + 8];
    }
    const int z = m->nnb;
    for (int k = 0; k <= z; ++k) {
        const double w = m->w[1 * 8 + k];                      // s = 1: k' = S
        if (m->w[0 * 8 + (z - k)] != w) ARG_FAIL("internal: acceptance table is not symmetric");
        uint64_t t = (uint64_t)floor(w * 4294967296.0);        // u <= w  <=>  U < floor(w 2^32)
        if (t > 4294967296ull) t = 4294967296ull;
        m->thr[k] = t;
        if (k <= z / 2 && t != 4294967296ull) ARG_FAIL("internal: dE <= 0 must always be accepted");
    }
    return B200MC_OK;
}

void fill_args(Bits* m, int colour, BitsPassArgs* a)
{
    a->own = m->vec[colour] + m->H; a->oth = m->vec[colour ^ 1] + m->H;
    a->L = m->Lloc; a->p0 = m->p0; a->pbeg = 0;
    for (int t = 0; t < 6; ++t) a->off[t] = m->off[colour][t];
    const int first = m->nnb / 2 + 1;                          // first non-trivial class: k' = 4 (3D) / 3 (2D)
    for (int c = 0; c < 3; ++c) {
        const int k = first + c;
        const uint64_t t = k <= m->nnb ? m->thr[k] : 0;
        a->always[c] = t >= 4294967296ull ? 0xFFFFFFFFu : 0u;
        a->thr[c] = t >= 4294967296ull ? 0xFFFFFFFFu : (uint32_t)t;
    }
    a->colour = (uint32_t)colour; a->draw = m->draw;
    for (int r = 0; r < 10; ++r) a->rk0[r] = m->seed + (uint32_t)r * PHILOX_W0;
    a->acc = m->d_acc;
}

int halo(Bits* m, int colour, cudaStream_t st = nullptr)
{
    if (!st) st = m->stream;
    if (m->nranks == 1) {
        COUNT_LAUNCH();
        bits_halo_kernel<<<(unsigned)((2 * m->H + 255) / 256), 256, 0, st>>>(m->vec[colour], m->L, m->H);
        CK(cudaGetLastError());
        return B200MC_OK;
    }
    // slab mode: my first H owned vectors become the HIGH halo of rank-1, my last H owned vectors the LOW halo of rank+1
    // (ncclSend/Recv over NVLink, one group per colour pass); crossing the end of the fold moves a site to the next bit-lane,
    // so rank 0 rotates its low halo and rank P-1 its high halo after the exchange
    uint4* v = m->vec[colour];
    const size_t bytes = (size_t)m->H * sizeof(uint4);
    int rc = dist_exchange_ring(m->comm, m->rank, m->nranks, v + m->H, v + m->Lloc, v, v + m->H + m->Lloc, bytes, st);
    if (rc) return rc;
    const unsigned nb = (unsigned)((m->H + 255) / 256);
    if (m->rank == 0) { COUNT_LAUNCH(); bits_rotate_kernel<<<nb, 256, 0, st>>>(v, m->H, +1); }
    if (m->rank == m->nranks - 1) { COUNT_LAUNCH(); bits_rotate_kernel<<<nb, 256, 0, st>>>(v + m->H + m->Lloc, m->H, -1); }
    CK(cudaGetLastError());
    return B200MC_OK;
}

// launch one colour pass over local vectors [pbeg, pbeg + n) on stream st
int launch_pass(Bits* m, BitsPassArgs a, int64_t pbeg, int64_t n, bool fuse, cudaStream_t st, bool short_blocks = false)
{
    a.pbeg = pbeg; a.L = n;
    const int64_t need = (n + 255) / 256;
    // short_blocks (the interior launch of the slab pass): one vector per thread, so that blocks retire all the time and the
    // exchange kernels of the other stream find room on the SMs (a resident grid-stride grid would hold them until it ends)
    const int grid = short_blocks ? (int)need : (int)(need < (int64_t)m->grid ? need : (int64_t)m->grid);
    COUNT_LAUNCH();
    if (fuse) {
        if (m->nnb == 6) bits_pass_kernel<6, 3, true><<<grid, 256, 0, st>>>(a);
        else bits_pass_kernel<4, 3, true><<<grid, 256, 0, st>>>(a);
    } else if (m->nnb == 6 && m->minb == 4) bits_pass_kernel<6, 4, false><<<grid, 256, 0, st>>>(a);
    else if (m->nnb == 6) bits_pass_kernel<6, 3, false><<<grid, 256, 0, st>>>(a);
    else bits_pass_kernel<4, 4, false><<<grid, 256, 0, st>>>(a);
    CK(cudaGetLastError());
    return B200MC_OK;
}

int sweep(Bits* m)
{
    if (m->fused_pending) m->want_fused = false;   // the sums of the previous sweep were never asked for
    m->obs_valid = false; m->fused_pending = false;
    // slab mode with room for an interior: the first / last H owned vectors (what the neighbours need) are updated first
    // and exchanged on the handle's stream while the interior runs on a second stream
    const bool split = m->nranks > 1 && m->Lloc >= 3 * m->H && m->aux;
    for (int colour = 0; colour < 2; ++colour) {
        BitsPassArgs a;
        fill_args(m, colour, &a);
        if (m->timing) {
            while (m->evs.size() < m->ev_used + 2) { cudaEvent_t e; CK(cudaEventCreate(&e)); m->evs.push_back(e); }
            CK(cudaEventRecord(m->evs[m->ev_used], m->stream));
        }
        const bool fuse = colour == 1 && m->want_fused;
        if (fuse) CK(cudaMemsetAsync(m->d_acc, 0, 2 * sizeof(unsigned long long), m->stream));
        int rc;
        if (split) {
            // boundary launches + exchange on a HIGH-priority stream, the interior on a low-priority one: the block scheduler
            // serves a later kernel only when the earlier one has no blocks left to dispatch unless the later one has
            // priority, so without it the exchange kernels sat behind the interior's 12 k blocks (measured: no overlap)
            CK(cudaEventRecord(m->ev_main, m->stream));
            CK(cudaStreamWaitEvent(m->hi, m->ev_main, 0));
            CK(cudaStreamWaitEvent(m->aux, m->ev_main, 0));
            if ((rc = launch_pass(m, a, 0, m->H, fuse, m->hi))) return rc;
            if ((rc = launch_pass(m, a, m->Lloc - m->H, m->H, fuse, m->hi))) return rc;
            if ((rc = launch_pass(m, a, m->H, m->Lloc - 2 * m->H, fuse, m->aux, true))) return rc;
            CK(cudaEventRecord(m->ev_aux, m->aux));
            if ((rc = halo(m, colour, m->hi))) return rc;
            CK(cudaEventRecord(m->ev_hi, m->hi));
            CK(cudaStreamWaitEvent(m->stream, m->ev_hi, 0));       // join: the next pass (and anything else) sees the whole colour
            CK(cudaStreamWaitEvent(m->stream, m->ev_aux, 0));
        } else {
            if ((rc = launch_pass(m, a, 0, m->Lloc, fuse, m->stream, m->short_blocks != 0))) return rc;
            if ((rc = halo(m, colour))) return rc;
        }
        if (m->timing) { CK(cudaEventRecord(m->evs[m->ev_used + 1], m->stream)); m->ev_used += 2; }
        if (fuse) m->fused_pending = true;
    }
    m->draw += 1;
    m->swept_since_measure = true;
    return B200MC_OK;
}

int measure(Bits* m)
{
    if (m->obs_valid) return B200MC_OK;
    if (!m->fused_pending) {
        CK(cudaMemsetAsync(m->d_acc, 0, 2 * sizeof(unsigned long long), m->stream));
        BitsPassArgs a;
        fill_args(m, 1, &a);
        COUNT_LAUNCH();
        if (m->nnb == 6) bits_measure_kernel<6><<<m->grid, 256, 0, m->stream>>>(a, m->vec[0] + m->H);
        else bits_measure_kernel<4><<<m->grid, 256, 0, m->stream>>>(a, m->vec[0] + m->H);
        CK(cudaGetLastError());
    }
    m->want_fused = m->swept_since_measure;   // measured after an update: the next sweeps accumulate the sums themselves
    m->swept_since_measure = false;
    m->fused_pending = false;
    if (m->nranks > 1) { int rc = dist_allreduce_u64(m->comm, m->d_acc, 2, m->stream); if (rc) return rc; }
    unsigned long long host[2];
    CK(cudaMemcpyAsync(host, m->d_acc, sizeof(host), cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    const int64_t X = (int64_t)m->nnb * m->Nc - (int64_t)host[0];      // unequal bonds
    m->obs_e = -(int64_t)(m->nnb / 2) * m->N + 2 * X;
    m->obs_m = 2 * (int64_t)host[1] - m->N;
    m->obs_valid = true;
    return B200MC_OK;
}

void destroy(Bits* m)
{
    cudaStreamSynchronize(m->stream);
    if (m->comm) dist_comm_destroy(m->comm);
    if (m->aux) { cudaStreamSynchronize(m->aux); cudaStreamSynchronize(m->hi); cudaStreamDestroy(m->aux); cudaStreamDestroy(m->hi); cudaEventDestroy(m->ev_main); cudaEventDestroy(m->ev_aux); cudaEventDestroy(m->ev_hi); }
    cudaFree(m->vec[0]); cudaFree(m->vec[1]); cudaFree(m->stage); cudaFree(m->d_acc);
    for (cudaEvent_t e : m->evs) cudaEventDestroy(e);
    delete m;
}

int fill(Bits* m, int value)
{
    m->obs_valid = false; m->fused_pending = false;
    const size_t bytes = (size_t)(m->Lloc + 2 * m->H) * sizeof(uint4);
    CK(cudaMemsetAsync(m->vec[0], value ? 0xFF : 0x00, bytes, m->stream));
    CK(cudaMemsetAsync(m->vec[1], value ? 0xFF : 0x00, bytes, m->stream));
    return B200MC_OK;
}

int create(void** out, int ndim, int64_t nx, int64_t ny, int64_t nz, double kbt, int32_t iseed, int rank = 0, int nranks = 1, const char* nccl_id = nullptr)
{
    if (nranks < 1 || rank < 0 || rank >= nranks) ARG_FAIL("bad rank %d / %d", rank, nranks);
    if (nranks > 1 && !nccl_id) ARG_FAIL("slab mode needs the NCCL unique id of the job (b200mc_dist_unique_id on rank 0, broadcast by the caller)");
    if (!out) ARG_FAIL("null handle pointer");
    *out = nullptr;
    if (!(kbt > 0.0)) ARG_FAIL("kbt must be > 0");
    if (ndim == 2) {
        if (nx < 3 || ny < 2 || !(nx & 1) || (ny & 1)) ARG_FAIL("ising2d (bit-packed): nx must be odd and ny even (helical colouring), got %lld x %lld", (long long)nx, (long long)ny);
        nz = 1;
    } else {
        if (nx < 3 || ny < 3 || nz < 2 || !(nx & 1) || !(ny & 1) || (nz & 1))
            ARG_FAIL("ising3d (bit-packed): nx and ny must be odd and nz even (helical colouring), got %lld x %lld x %lld", (long long)nx, (long long)ny, (long long)nz);
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        snprintf(g_b200mc_err, sizeof(g_b200mc_err), "no CUDA device: this library has no CPU fallback");
        return B200MC_ERR_CUDA;
    }
    const int64_t N = nx * ny * nz, Nc = N / 2;
    const int64_t h = (nx - 1) / 2, g = (nx * ny - 1) / 2;
    const int64_t H = (ndim == 3 ? g : h) + 1;
    if (Nc % 128) {
        snprintf(g_b200mc_err, sizeof(g_b200mc_err), "bit-packed Ising: nx ny nz / 2 = %lld must be a multiple of 128 (e.g. the last extent a multiple of 256)", (long long)Nc);
        return B200MC_ERR_UNSUPPORTED;
    }
    const int64_t L = Nc / 128;
    if (L < H) {
        snprintf(g_b200mc_err, sizeof(g_b200mc_err), "bit-packed Ising: fold length %lld shorter than the halo %lld (lattice too thin along the last axis)", (long long)L, (long long)H);
        return B200MC_ERR_UNSUPPORTED;
    }
    if (L % nranks || L / nranks < H) {
        snprintf(g_b200mc_err, sizeof(g_b200mc_err), "bit-packed Ising slabs: the fold (%lld vectors) must split evenly over %d ranks into shares not shorter than the halo (%lld)", (long long)L, nranks, (long long)H);
        return B200MC_ERR_UNSUPPORTED;
    }
    Bits* m = new (std::nothrow) Bits();
    if (!m) ARG_FAIL("out of host memory");
    m->rank = rank; m->nranks = nranks; m->comm = nullptr; m->aux = m->hi = nullptr; m->ev_main = m->ev_aux = m->ev_hi = nullptr; m->Lloc = L / nranks; m->p0 = (int64_t)rank * (L / nranks);
    m->ndim = ndim; m->nnb = ndim == 3 ? 6 : 4; m->nx = nx; m->ny = ny; m->nz = ndim == 3 ? nz : 0;
    m->N = N; m->Nc = Nc; m->L = L; m->H = H; m->P = ndim == 3 ? nx * ny : nx;
    for (int c = 0; c < 2; ++c) {
        const int64_t o[6] = {-1 + c, c, h + c, -h - 1 + c, g + c, -g - 1 + c};
        for (int t = 0; t < 6; ++t) m->off[c][t] = o[t];
    }
    m->vec[0] = m->vec[1] = nullptr; m->stage = nullptr; m->d_acc = nullptr;
    m->stream = 0; m->beta = 1 / kbt; m->seed = (uint32_t)iseed; m->draw = 0; m->obs_valid = false; m->timing = false; m->ev_used = 0;
    m->want_fused = false; m->fused_pending = false; m->swept_since_measure = false;
    const size_t bytes = (size_t)(m->Lloc + 2 * H) * sizeof(uint4);
    if (cudaMalloc(&m->vec[0], bytes) != cudaSuccess || cudaMalloc(&m->vec[1], bytes) != cudaSuccess ||
        cudaMalloc(&m->d_acc, 2 * sizeof(unsigned long long)) != cudaSuccess) {
        snprintf(g_b200mc_err, sizeof(g_b200mc_err), "cudaMalloc failed (%zu bytes per colour)", bytes);
        cudaGetLastError();
        destroy(m); return B200MC_ERR_CUDA;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    { const char* t = getenv("B200MC_BITS_MINB"); m->minb = (t && atoi(t) == 4) ? 4 : 3; }
    // (measured 5-6 % faster than the resident grid: warps whose sites are decided early retire and the next block backfills)
    { const char* t = getenv("B200MC_BITS_SHORT"); m->short_blocks = (t && atoi(t) == 0) ? 0 : 1; }
    const int64_t need = (m->Lloc + 255) / 256;
    const int per_sm = ndim == 3 ? m->minb : 4;
    m->grid = (int)(need < (int64_t)sms * per_sm ? need : (int64_t)sms * per_sm);
    if (nranks > 1 && !(getenv("B200MC_BITS_NOSPLIT") && atoi(getenv("B200MC_BITS_NOSPLIT")))) {
        int plo = 0, phi = 0;
        cudaDeviceGetStreamPriorityRange(&plo, &phi);     // (numerically lower = higher priority)
        if (cudaStreamCreateWithPriority(&m->aux, cudaStreamNonBlocking, plo) != cudaSuccess || cudaStreamCreateWithPriority(&m->hi, cudaStreamNonBlocking, phi) != cudaSuccess ||
            cudaEventCreateWithFlags(&m->ev_main, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&m->ev_aux, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&m->ev_hi, cudaEventDisableTiming) != cudaSuccess) {
            snprintf(g_b200mc_err, sizeof(g_b200mc_err), "cudaStreamCreate / cudaEventCreate failed");
            destroy(m); return B200MC_ERR_CUDA;
        }
    }
    int rc = nranks > 1 ? dist_comm_init(&m->comm, rank, nranks, nccl_id) : B200MC_OK;
    if (!rc) rc = build_tables(m);
    if (!rc) rc = fill(m, 1);      // like the reference's init: all up
    if (rc) { destroy(m); return rc; }
    *out = m;
    return B200MC_OK;
}

int set_random(Bits* m)
{
    m->obs_valid = false; m->fused_pending = false;
    for (int c = 0; c < 2; ++c) {
        COUNT_LAUNCH();
        bits_random_kernel<<<(unsigned)((m->Lloc + 255) / 256), 256, 0, m->stream>>>(m->vec[c] + m->H, m->Lloc, m->p0, m->seed, m->draw, (uint32_t)c);
        CK(cudaGetLastError());
        int rc = halo(m, c);
        if (rc) return rc;
    }
    m->draw += 1;
    return B200MC_OK;
}

int get_spins(Bits* m, int32_t* out)
{
    if (!out) ARG_FAIL("null output");
    const int64_t n = m->N + 2 * m->P;
    if (!m->stage) CK(cudaMalloc(&m->stage, (size_t)n * sizeof(int32_t)));
    COUNT_LAUNCH();
    bits_export_kernel<<<(unsigned)((n + 255) / 256), 256, 0, m->stream>>>(m->vec[0] + m->H, m->vec[1] + m->H, m->N, m->L, m->P, m->ndim == 2, m->p0, m->Lloc, m->stage);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, m->stage, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    return B200MC_OK;
}

int set_spins(Bits* m, const int32_t* in)
{
    if (!in) ARG_FAIL("null input");
    const int64_t n = m->N + 2 * m->P;
    // validate before touching the lattice (the halo cells of `in` are ignored, like the int8 handles)
    const int32_t lo = m->ndim == 2 ? -1 : 0;
    for (int64_t i = 0; i < m->N; ++i) {
        const int32_t v = in[i + m->P];
        if (v != 1 && v != lo) ARG_FAIL("set_spins: value %d at site %lld (must be %s)", v, (long long)(i + 1), m->ndim == 2 ? "-1 / +1" : "0 / 1");
    }
    if (!m->stage) CK(cudaMalloc(&m->stage, (size_t)n * sizeof(int32_t)));
    m->obs_valid = false; m->fused_pending = false;
    CK(cudaMemcpyAsync(m->stage, in, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, m->stream));
    COUNT_LAUNCH();
    bits_import_kernel<<<(unsigned)((8 * m->Lloc + 255) / 256), 256, 0, m->stream>>>(m->vec[0] + m->H, m->vec[1] + m->H, m->L, m->P, m->ndim == 2, m->p0, m->Lloc, m->stage);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(m->stream));
    int rc = halo(m, 0);
    if (!rc) rc = halo(m, 1);
    return rc;
}

}  // namespace

#define HB(h) (reinterpret_cast<Bits*>(h))
#define CHECK_B(h, nd) do { if (!(h) || HB(h)->ndim != (nd)) ARG_FAIL("invalid handle"); } while (0)

extern "C" {

#define BITS_ABI(PFX, ND)                                                                                                         \
    int PFX##_destroy(void* h) { if (h) destroy(HB(h)); return B200MC_OK; }                                                       \
    int PFX##_set_stream(void* h, void* s) { CHECK_B(h, ND); HB(h)->stream = (cudaStream_t)s; return B200MC_OK; }                 \
    int PFX##_skip_curand(void* h, int64_t n) { CHECK_B(h, ND); if (n < 0) ARG_FAIL("n_skip < 0");                                \
        HB(h)->draw += (uint64_t)((n + HB(h)->N - 1) / HB(h)->N); return B200MC_OK; }                                             \
    int PFX##_set_allup_spin(void* h) { CHECK_B(h, ND); return fill(HB(h), 1); }                                                  \
    int PFX##_set_random_spin(void* h) { CHECK_B(h, ND); return set_random(HB(h)); }                                              \
    int PFX##_set_kbt(void* h, double kbt) { CHECK_B(h, ND); if (!(kbt > 0.0)) ARG_FAIL("kbt must be > 0"); HB(h)->beta = 1 / kbt; return build_tables(HB(h)); } \
    int PFX##_set_beta(void* h, double beta) { CHECK_B(h, ND); if (!(beta >= 0.0)) ARG_FAIL("beta must be >= 0"); HB(h)->beta = beta; return build_tables(HB(h)); } \
    int PFX##_update(void* h) { CHECK_B(h, ND); return sweep(HB(h)); }                                                            \
    int PFX##_update_n(void* h, int32_t n) { CHECK_B(h, ND); for (int i = 0; i < n; ++i) { int rc = sweep(HB(h)); if (rc) return rc; } return B200MC_OK; } \
    int PFX##_calc_energy_sum(void* h, int64_t* e) { CHECK_B(h, ND); int rc = measure(HB(h)); if (!rc && e) *e = HB(h)->obs_e; return rc; } \
    int PFX##_calc_magne_sum(void* h, int64_t* mg) { CHECK_B(h, ND); int rc = measure(HB(h)); if (!rc && mg) *mg = HB(h)->obs_m; return rc; } \
    int PFX##_measure(void* h, int64_t* e, int64_t* mg) { CHECK_B(h, ND); int rc = measure(HB(h)); if (!rc) { if (e) *e = HB(h)->obs_e; if (mg) *mg = HB(h)->obs_m; } return rc; } \
    int PFX##_get_spins(void* h, int32_t* out) { CHECK_B(h, ND); return get_spins(HB(h), out); }                                  \
    int PFX##_set_spins(void* h, const int32_t* in) { CHECK_B(h, ND); return set_spins(HB(h), in); }                              \
    int64_t PFX##_nx(void* h) { return h ? HB(h)->nx : -1; }                                                                      \
    int64_t PFX##_ny(void* h) { return h ? HB(h)->ny : -1; }                                                                      \
    int64_t PFX##_nall(void* h) { return h ? HB(h)->N : -1; }                                                                     \
    double PFX##_kbt(void* h) { return h ? 1 / HB(h)->beta : 0.0; }                                                               \
    double PFX##_beta(void* h) { return h ? HB(h)->beta : 0.0; }                                                                  \
    int PFX##_rank_info(void* h, int32_t* rank, int32_t* nranks) { CHECK_B(h, ND); if (rank) *rank = HB(h)->rank; if (nranks) *nranks = HB(h)->nranks; return B200MC_OK; } \
    int PFX##_sync(void* h) { CHECK_B(h, ND); CK(cudaStreamSynchronize(HB(h)->stream)); return B200MC_OK; }                       \
    int PFX##_set_timing(void* h, int32_t on) { CHECK_B(h, ND); HB(h)->timing = on != 0; HB(h)->ev_used = 0; return B200MC_OK; }  \
    int PFX##_get_timing(void* h, int64_t* launches, double* total_ms) { CHECK_B(h, ND); Bits* m = HB(h);                         \
        CK(cudaStreamSynchronize(m->stream)); double tot = 0; for (size_t i = 0; i + 1 < m->ev_used; i += 2) { float ms = 0;      \
            CK(cudaEventElapsedTime(&ms, m->evs[i], m->evs[i + 1])); tot += ms; }                                                 \
        if (launches) *launches = (int64_t)(m->ev_used / 2); if (total_ms) *total_ms = tot; m->ev_used = 0; return B200MC_OK; }

BITS_ABI(b200mc_ising3dp, 3)
BITS_ABI(b200mc_ising2dp, 2)

int b200mc_ising3dp_create(void** h, int64_t nx, int64_t ny, int64_t nz, double kbt, int32_t iseed) { return create(h, 3, nx, ny, nz, kbt, iseed); }
int b200mc_ising2dp_create(void** h, int64_t nx, int64_t ny, double kbt, int32_t iseed) { return create(h, 2, nx, ny, 0, kbt, iseed); }
int b200mc_ising3dp_create_slab(void** h, int64_t nx, int64_t ny, int64_t nz, double kbt, int32_t iseed, int32_t rank, int32_t nranks, const char* nccl_id)
{
    return create(h, 3, nx, ny, nz, kbt, iseed, rank, nranks, nccl_id);
}
int b200mc_ising2dp_create_slab(void** h, int64_t nx, int64_t ny, double kbt, int32_t iseed, int32_t rank, int32_t nranks, const char* nccl_id)
{
    return create(h, 2, nx, ny, 0, kbt, iseed, rank, nranks, nccl_id);
}
int64_t b200mc_ising3dp_nz(void* h) { return h ? HB(h)->nz : -1; }
int b200mc_ising3dp_get_ws(void* h, double out[14])
{
    CHECK_B(h, 3);
    for (int s = 0; s < 2; ++s) for (int S = 0; S <= 6; ++S) out[7 * s + S] = HB(h)->w[s * 8 + S];
    return B200MC_OK;
}

}  // extern "C"
