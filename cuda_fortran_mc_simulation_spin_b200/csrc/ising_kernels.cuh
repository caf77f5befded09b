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

struct IsingTab {
    uint32_t tlo[2], thi[2];  // [s]: bytes T'(S) for S = 0..3 / S = 4..7
    uint32_t low25[2][8];     // [s][S]
};

struct RingPassArgs {
    uint4* own;         // colour being updated, vector index 0 = position -H
    const uint4* oth;   // the other colour
    int64_t nvec;       // owned positions (L)
    int64_t H;
    int64_t p0;         // global position of the first owned vector
    int64_t off[6];     // neighbour vector offsets for this colour
    uint32_t seed;
    uint32_t colour;
    uint64_t draw;
};

enum { METHOD_METROPOLIS = 0, METHOD_HEATBATH = 1 };

// 8 sites: own words (w0 = lanes 0-3, w1 = lanes 4-7 of the group), neighbour
// sums (s0, s1), random words (ra -> lanes (0,4,1,5), rb -> lanes (2,6,3,7)).
template <int METHOD>
__device__ __forceinline__ void ising_group(uint32_t& w0, uint32_t& w1, uint32_t s0, uint32_t s1,
                                            uint32_t ra, uint32_t rb, const IsingTab& tab,
                                            const RingPassArgs& a, uint64_t pglob, int group)
{
    const uint32_t sp = s0 + (s1 << 4);  // nibble-packed sums: byte j = S(lane j) | S(lane 4+j) << 4
    const uint32_t selA = sp, selB = sp >> 16;
    const uint32_t oA = prmt(w0, w1, 0x5140u);  // lanes (0,4,1,5)
    const uint32_t oB = prmt(w0, w1, 0x7362u);  // lanes (2,6,3,7)
    uint32_t tA, tB;
    if (METHOD == METHOD_METROPOLIS) {
        const uint32_t mA = oA * 0xFFu, mB = oB * 0xFFu;  // 0xFF where the spin is up
        const uint32_t uA = prmt(tab.tlo[1], tab.thi[1], selA), dA = prmt(tab.tlo[0], tab.thi[0], selA);
        const uint32_t uB = prmt(tab.tlo[1], tab.thi[1], selB), dB = prmt(tab.tlo[0], tab.thi[0], selB);
        tA = (uA & mA) | (dA & ~mA);
        tB = (uB & mB) | (dB & ~mB);
    } else {
        tA = prmt(tab.tlo[0], tab.thi[0], selA);
        tB = prmt(tab.tlo[0], tab.thi[0], selB);
    }
    // z = b7 + 128 - T7 per byte (no carries: <= 255).  bit 7 clear <=> b7 < T7 <=> accept.
    uint32_t zA = (ra & 0x7F7F7F7Fu) + tA;
    uint32_t zB = (rb & 0x7F7F7F7Fu) + tB;
    // ties (byte == 0x80 <=> b7 == T7): decide with 25 more bits.  Rare.
    const uint32_t tieA = zero_byte_mask(zA ^ 0x80808080u);
    const uint32_t tieB = zero_byte_mask(zB ^ 0x80808080u);
    if (tieA | tieB) {
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            uint32_t tie = half ? tieB : tieA;
            if (!tie) continue;
            const uint32_t o = half ? oB : oA, sel = half ? selB : selA;
            const uint4 r2 = philox4x32_10(mk_ctr(pglob, a.draw, a.colour, 1u + 2u * group + half),
                                           make_uint2(a.seed, TAG_ISING));
            const uint32_t rr[4] = {r2.x, r2.y, r2.z, r2.w};
            uint32_t clr = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // zero_byte_mask can flag a byte above a true zero byte; re-test exactly
                const uint32_t zb = ((half ? zB : zA) >> (8 * j)) & 0xFFu;
                if (zb != 0x80u) continue;
                const uint32_t S = (sel >> (4 * j)) & 0xFu;
                const uint32_t s = (METHOD == METHOD_METROPOLIS) ? ((o >> (8 * j)) & 1u) : 0u;
                if ((rr[j] & 0x1FFFFFFu) < tab.low25[s][S & 7u]) clr |= 0x80u << (8 * j);
            }
            if (half) zB &= ~clr; else zA &= ~clr;
        }
    }
    const uint32_t fA = (~zA >> 7) & 0x01010101u;
    const uint32_t fB = (~zB >> 7) & 0x01010101u;
    uint32_t nA, nB;
    if (METHOD == METHOD_METROPOLIS) { nA = oA ^ fA; nB = oB ^ fB; }
    else { nA = fA; nB = fB; }
    w0 = prmt(nA, nB, 0x6420u);
    w1 = prmt(nA, nB, 0x7531u);
}

template <int NNB, int METHOD>
__global__ void __launch_bounds__(256)
ising_pass_kernel(const __grid_constant__ RingPassArgs a, const __grid_constant__ IsingTab tab)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < a.nvec; v += stride) {
        const int64_t q = v + a.H;
        uint4 o = ld_own(a.own + q);
        uint4 n = ld_other(a.oth + q + a.off[0]);
        uint4 m = ld_other(a.oth + q + a.off[1]);
        uint4 S = make_uint4(n.x + m.x, n.y + m.y, n.z + m.z, n.w + m.w);
#pragma unroll
        for (int j = 2; j < NNB; j += 2) {
            n = ld_other(a.oth + q + a.off[j]);
            m = ld_other(a.oth + q + a.off[j + 1]);
            S.x += n.x + m.x; S.y += n.y + m.y; S.z += n.z + m.z; S.w += n.w + m.w;
        }
        const uint64_t pglob = (uint64_t)(a.p0 + v);
        const uint4 r = philox4x32_10(mk_ctr(pglob, a.draw, a.colour, 0u), make_uint2(a.seed, TAG_ISING));
        ising_group<METHOD>(o.x, o.y, S.x, S.y, r.x, r.y, tab, a, pglob, 0);
        ising_group<METHOD>(o.z, o.w, S.z, S.w, r.z, r.w, tab, a, pglob, 1);
        st_own(a.own + q, o);
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
                     int64_t Nc, int64_t ptail, unsigned long long* acc)
{
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
        uint32_t x = __dp4a(X.x, 0x01010101u, 0u);
        x = __dp4a(X.y, 0x01010101u, x);
        x = __dp4a(X.z, 0x01010101u, x);
        x = __dp4a(X.w, 0x01010101u, x);
        uint32_t m = __dp4a(o.x + a0.x, 0x01010101u, 0u);
        m = __dp4a(o.y + a0.y, 0x01010101u, m);
        m = __dp4a(o.z + a0.z, 0x01010101u, m);
        m = __dp4a(o.w + a0.w, 0x01010101u, m);
        part[0] += x;
        part[1] += m;
    }
    block_atomic_add<2>(acc, part);
}

// set_random_spin (reference: set_random_spin_sub, src/ising3d_gpu_m.f90:91-100):
// s = (u < 0.5) with u = (U+1) 2^-32, U = R[lane & 3] of
// philox(ctr(p, draw, colour, lane >> 2), (seed, TAG_INIT)).
__global__ void __launch_bounds__(256)
ring_random_bits_kernel(uint4* own, int64_t nvec, int64_t H, int64_t p0, uint32_t seed,
                        uint64_t draw, uint32_t colour)
{
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvec) return;
    uint32_t w[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const uint4 r = philox4x32_10(mk_ctr((uint64_t)(p0 + v), draw, colour, g), make_uint2(seed, TAG_INIT));
        // (U+1) 2^-32 < 0.5  <=>  U < 2^31 - 1
        w[g] = (r.x < 0x7FFFFFFFu ? 1u : 0u) | (r.y < 0x7FFFFFFFu ? 0x100u : 0u) |
               (r.z < 0x7FFFFFFFu ? 0x10000u : 0u) | (r.w < 0x7FFFFFFFu ? 0x1000000u : 0u);
    }
    own[v + H] = make_uint4(w[0], w[1], w[2], w[3]);
}
