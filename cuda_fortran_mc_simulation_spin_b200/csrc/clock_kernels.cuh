// Device kernels for the q-state clock model on the folded-ring layout
// (reference: src/clock_gpu_m.f90, src/clock_gpu_multi_m.f90).
//
// States are int8 in the same 16-lane folded colour arrays as the Ising models
// (ring.cuh), so the four neighbour vectors are aligned 128-bit loads.  The
// update itself is per-site scalar work: a q^6 table lookup.  The reference
// keeps ws(up,down,left,right,before,after) as q^6 real64 in global memory
// (373 KB for q = 6, src/clock_gpu_m.f90:77,203); here the host builds the same
// real64 table, converts every entry to the integer threshold of the 32-bit
// uniform and compresses BY VALUE: a one-byte class id per index (q^6 bytes,
// staged in shared memory) + <= 256 thresholds.
#pragma once
#include "common.cuh"
#include "ising_kernels.cuh"  // RingPassArgs, philox_rk
#include "clock_word.cuh"

#define CLOCK_MAX_CLASSES 256        // class ids fit a byte up to here (thresholds staged in shared memory)
#define CLOCK_MAX_CLASSES16 65536    // beyond: 16-bit class ids, thresholds read through L1 / L2

// Philox4x32-10 with both round-key schedules supplied by the host
__device__ __forceinline__ uint4 philox_rk2(uint4 c, const uint32_t (&rk0)[10], const uint32_t (&rk1)[10])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t lo0, hi0, lo1, hi1;
        mulwide(PHILOX_M0, c.x, lo0, hi0);
        mulwide(PHILOX_M1, c.z, lo1, hi1);
        uint4 n;
        n.x = hi1 ^ c.y ^ rk0[r];
        n.y = lo1;
        n.z = hi0 ^ c.w ^ rk1[r];
        n.w = lo0;
        c = n;
    }
    return c;
}

struct ClockArgs {
    RingPassArgs r;
    const uint8_t* cls;        // q^6 class ids (one byte each, or two when cls16), index up + q(down + q(left + q(right + q(cur + q next))))
    int cls16;                 // more than 256 distinct thresholds (q >= 14): 16-bit class ids, thresholds in global memory
    const uint64_t* thr;       // per class: accept iff U < thr   (0 .. 2^32)
    uint32_t q;
    uint32_t tab_bytes;        // q^6
    uint32_t replica;
    uint32_t rk0[10];          // Philox round keys: seed + r W0
    uint32_t rk1[10];          //                    TAG_CLOCK + replica + r W1
    int cls_in_smem;           // class table fits in shared memory
    const uint16_t* thr16;     // direct lookup (q <= 6): T[next][F] = thr >> 17, F = up + q down + q^2 left + q^3 right + q^4 cur
};

// exact evaluation of the 16 sites of vector pglob (RNG contract v3, clock_word.cuh): both stages of the accept uniforms,
// class table, full 33-bit threshold.
// accept iff U_a < thr[class]; proposal next = floor(W_e q / 2^32) = the reference's min(floor(next_states q), q - 1)
// (src/clock_gpu_m.f90:211, with the u == 1 clamp of SURVEY Q4) on the contract's next_states
template <typename CLS>
__device__ __forceinline__ uint4 clock_vector_exact_body(const ClockArgs& a, const CLS* cls, const uint64_t* thr, uint64_t pglob,
                                                         const uint4& o, const uint4& nu, const uint4& nd, const uint4& nl, const uint4& nr)
{
    const RingPassArgs& r = a.r;
    const uint32_t q = a.q, q2 = q * q, q3 = q2 * q, q4 = q3 * q, q5 = q4 * q;
    uint32_t X[12], Y[12];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const uint4 R = philox_rk2(mk_ctr(pglob, r.draw, r.colour, (uint32_t)i), a.rk0, a.rk1);
        const uint4 R2 = philox_rk2(mk_ctr(pglob, r.draw, r.colour, 4u + (uint32_t)i), a.rk0, a.rk1);
        X[4 * i] = R.x; X[4 * i + 1] = R.y; X[4 * i + 2] = R.z; X[4 * i + 3] = R.w;
        Y[4 * i] = R2.x; Y[4 * i + 1] = R2.y; Y[4 * i + 2] = R2.z; Y[4 * i + 3] = R2.w;
    }
    const uint32_t ow[4] = {o.x, o.y, o.z, o.w}, lw[4] = {nl.x, nl.y, nl.z, nl.w}, rw[4] = {nr.x, nr.y, nr.z, nr.w},
                   uw[4] = {nu.x, nu.y, nu.z, nu.w}, dw[4] = {nd.x, nd.y, nd.z, nd.w};
    uint32_t res[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        uint32_t W = X[3 * w + 2], out = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const uint32_t Ua = clk_accept32(X[3 * w + (e & 1)], Y[3 * w + (e & 1)], e >> 1);
            uint32_t nxt;
            mulwide(W, q, W, nxt);
            const int sh = 8 * e;
            const uint32_t cur = (ow[w] >> sh) & 0xFFu;
            const uint32_t idx = ((uw[w] >> sh) & 0xFFu) + q * ((dw[w] >> sh) & 0xFFu) + q2 * ((lw[w] >> sh) & 0xFFu) +
                                 q3 * ((rw[w] >> sh) & 0xFFu) + q4 * cur + q5 * nxt;
            const uint64_t t = thr[cls[idx]];
            out |= (((unsigned long long)Ua < t) ? nxt : cur) << sh;
        }
        res[w] = out;
    }
    return make_uint4(res[0], res[1], res[2], res[3]);
}
// (arguments and result by value: registers, not local memory, on the caller's hot path)
__device__ __noinline__ uint4 clock_vector_exact(const ClockArgs& a, uint64_t pglob, uint4 o, uint4 nu, uint4 nd, uint4 nl, uint4 nr)
{
    return clock_vector_exact_body<uint8_t>(a, a.cls, a.thr, pglob, o, nu, nd, nl, nr);   // (direct path: q <= 6, at most 15 classes)
}

// class-table pass (any q): every site with its full 32-bit uniforms.  CLS = uint8_t: at most 256 distinct thresholds, staged
// in shared memory together with the class ids when they fit (q <= 7), else ids through L1 / L2; CLS = uint16_t (q >= 14):
// ids and thresholds through L1 / L2.
template <typename CLS>
__global__ void __launch_bounds__(256)
clock_pass_kernel(const __grid_constant__ ClockArgs a)
{
    extern __shared__ __align__(16) uint8_t sm[];
    uint64_t* sthr = reinterpret_cast<uint64_t*>(sm);               // CLOCK_MAX_CLASSES * 8 bytes
    uint8_t* scls = sm + CLOCK_MAX_CLASSES * sizeof(uint64_t);      // q^6 bytes (if cls_in_smem)
    const bool wide = sizeof(CLS) == 2;
    if (!wide) {
        for (int i = threadIdx.x; i < CLOCK_MAX_CLASSES; i += blockDim.x) sthr[i] = a.thr[i];
        if (a.cls_in_smem) {
            const uint4* src = reinterpret_cast<const uint4*>(a.cls);
            uint4* dst = reinterpret_cast<uint4*>(scls);
            for (uint32_t i = threadIdx.x; i < (a.tab_bytes + 15) / 16; i += blockDim.x) dst[i] = src[i];
        }
        __syncthreads();
    }
    const CLS* cls = (!wide && a.cls_in_smem) ? reinterpret_cast<const CLS*>(scls) : reinterpret_cast<const CLS*>(a.cls);
    const uint64_t* thr = wide ? a.thr : sthr;

    const RingPassArgs& r = a.r;
    uint4* own = r.own + r.H;
    const uint4* oth = r.oth + r.H;
    const int nvec = (int)r.nvec;
    const int stride = gridDim.x * blockDim.x;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        const uint4 o = own[v];
        const uint4 nl = ld_other(oth + v + (int)r.off[0]);   // i-1  left
        const uint4 nr = ld_other(oth + v + (int)r.off[1]);   // i+1  right
        const uint4 nu = ld_other(oth + v + (int)r.off[2]);   // i+nx up
        const uint4 nd = ld_other(oth + v + (int)r.off[3]);   // i-nx down
        own[v] = clock_vector_exact_body<CLS>(a, cls, thr, (uint64_t)(r.p0 + v), o, nu, nd, nl, nr);
    }
}

// Direct-table variant (q <= 6; one CLOCK_DIRECT_THREADS-thread block per SM): clock_word.cuh.  Shared memory: the table
// T[next][F] (q slabs of 2 q^5 bytes).  Q = q at compile time (0: run time).
#define CLOCK_DIRECT_THREADS 768
template <int Q, int THREADS = CLOCK_DIRECT_THREADS>
__global__ void __launch_bounds__(THREADS, 1)
clock_pass_direct_kernel(const __grid_constant__ ClockArgs a)
{
    extern __shared__ __align__(16) uint8_t tab[];
    const uint32_t q = Q ? (uint32_t)Q : a.q;
    {
        const uint4* src = reinterpret_cast<const uint4*>(a.thr16);
        uint4* dst = reinterpret_cast<uint4*>(tab);
        for (uint32_t i = threadIdx.x; i < (2u * a.tab_bytes + 15) / 16; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const RingPassArgs& r = a.r;
    uint4* own = r.own + r.H;
    const uint4* oth = r.oth + r.H;
    const int nvec = (int)r.nvec;
    const uint32_t kstride = Q ? 2u * Q * Q * Q * Q * Q : 2u * (a.tab_bytes / a.q);
    const int stride = gridDim.x * blockDim.x;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        const uint4 o = own[v];
        const uint4 nl = ld_other(oth + v + (int)r.off[0]);   // i-1  left
        const uint4 nr = ld_other(oth + v + (int)r.off[1]);   // i+1  right
        const uint4 nu = ld_other(oth + v + (int)r.off[2]);   // i+nx up
        const uint4 nd = ld_other(oth + v + (int)r.off[3]);   // i-nx down
        const uint32_t ow[4] = {o.x, o.y, o.z, o.w}, lw[4] = {nl.x, nl.y, nl.z, nl.w}, rw[4] = {nr.x, nr.y, nr.z, nr.w},
                       uw[4] = {nu.x, nu.y, nu.z, nu.w}, dw[4] = {nd.x, nd.y, nd.z, nd.w};
        const uint64_t pglob = (uint64_t)(r.p0 + v);
        uint32_t X[12];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const uint4 R = philox_rk2(mk_ctr(pglob, r.draw, r.colour, (uint32_t)i), a.rk0, a.rk1);
            X[4 * i] = R.x; X[4 * i + 1] = R.y; X[4 * i + 2] = R.z; X[4 * i + 3] = R.w;
        }
        uint32_t res[4], amin = 0x7FFF7FFFu;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            uint32_t Fe, Fo;
            clk_index_fields(uw[w], dw[w], lw[w], rw[w], ow[w], q, Fe, Fo);
            res[w] = clk_word_fast<false>(ow[w], Fe, Fo, X[3 * w], X[3 * w + 1], X[3 * w + 2], tab, kstride, q, q, amin);
        }
        uint4 out = make_uint4(res[0], res[1], res[2], res[3]);
        if (clk_accept_tie(amin)) out = clock_vector_exact(a, pglob, o, nu, nd, nl, nr);   // rare (2^-15 per site): redo the vector exactly
        own[v] = out;
    }
}

// reference-stream pass: uniforms from device arrays in the reference's index order, real64
// compare against the real64 table (clock_gpu_m: reject when r > w; multi: reject when r >= w)
__global__ void __launch_bounds__(256)
clock_pass_randoms_kernel(const __grid_constant__ ClockArgs a, const double* __restrict__ randoms,
                          const double* __restrict__ next_states, const double* __restrict__ ws,
                          int64_t L, int64_t Nc, int strict)
{
    const RingPassArgs& r = a.r;
    uint4* own = r.own + r.H;
    const uint4* oth = r.oth + r.H;
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= r.nvec) return;
    const uint32_t q = a.q;
    const uint8_t* ob = reinterpret_cast<const uint8_t*>(own + v);
    const uint8_t* lb = reinterpret_cast<const uint8_t*>(oth + v + r.off[0]);
    const uint8_t* rb = reinterpret_cast<const uint8_t*>(oth + v + r.off[1]);
    const uint8_t* ub = reinterpret_cast<const uint8_t*>(oth + v + r.off[2]);
    const uint8_t* db = reinterpret_cast<const uint8_t*>(oth + v + r.off[3]);
    uint8_t out[16];
    for (int b = 0; b < 16; ++b) {
        out[b] = ob[b];
        const int64_t k = (int64_t)b * L + (r.p0 + v);
        if (k >= Nc) continue;
        const int64_t i = 2 * k + r.colour;
        int nxt = (int)floor(next_states[i] * (double)q);
        if (nxt >= (int)q) nxt = (int)q - 1;
        const size_t idx = (size_t)ub[b] + q * ((size_t)db[b] + q * ((size_t)lb[b] + q * ((size_t)rb[b] + q * ((size_t)ob[b] + (size_t)q * nxt))));
        const double w = ws[idx], u = randoms[i];
        if (strict ? (u >= w) : (u > w)) continue;
        out[b] = (uint8_t)nxt;
    }
    uint8_t* dst = reinterpret_cast<uint8_t*>(own + v);
    for (int b = 0; b < 16; ++b) dst[b] = out[b];
}

// set_random_spin: s = min(floor(u q), q-1), u = (U+1) 2^-32, U = R[lane & 3] of
// philox(ctr(p, draw, colour, lane >> 2), (seed', TAG_INIT))   (src/clock_gpu_m.f90:94-104)
__global__ void __launch_bounds__(256)
clock_random_kernel(uint4* own, int64_t nvec, int64_t H, int64_t p0, uint32_t seed, uint64_t draw,
                    uint32_t colour, uint32_t q)
{
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvec) return;
    uint32_t w[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const uint4 R = philox4x32_10(mk_ctr((uint64_t)(p0 + v), draw, colour, g), make_uint2(seed, TAG_INIT));
        const uint32_t rr[4] = {R.x, R.y, R.z, R.w};
        uint32_t o = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t s = (uint32_t)((((unsigned long long)rr[j] + 1ull) * q) >> 32);
            o |= min(s, q - 1) << (8 * j);
        }
        w[g] = o;
    }
    own[v + H] = make_uint4(w[0], w[1], w[2], w[3]);
}

// Integer observables (exact): acc[c] += #{sites in state c};
// acc[64 + d] += #{bonds (i, i-1) with (s(i-1) - s(i)) mod q = d};
// acc[128 + d] += same for bonds (i, i-nx).          (src/clock_gpu_m.f90:245-280:
// E = sum_i etab(s(i-nx), s(i-1), s(i)), M = sum_i cos(psi s(i)); both only depend on these counts)
// q <= 8: bytes -> one-hot bytes with PRMT as LUT, then AND + POPC per state (SWAR, 8 sites per
// step); q > 8: shared-memory atomics.

// (a - b) mod q per byte, a, b in [0, q)
__device__ __forceinline__ uint32_t submod_q(uint32_t a, uint32_t b, uint32_t q)
{
    const uint32_t d = a + q * 0x01010101u - b;                   // 1 .. 2q-1, no borrows
    const uint32_t ge = ((d + (0x80u - q) * 0x01010101u) >> 7) & 0x01010101u;  // d >= q
    return d - ge * q;
}
// one-hot bytes (1 << value) for 8 values given as two byte-words; dead = 0xFF bytes where no site
__device__ __forceinline__ void onehot8(uint32_t w0, uint32_t w1, uint32_t keep0, uint32_t keep1,
                                        uint32_t& hA, uint32_t& hB)
{
    const uint32_t p = w0 + (w1 << 4);
    hA = prmt(0x08040201u, 0x80402010u, p);
    hB = prmt(0x08040201u, 0x80402010u, p >> 16);
    // hA holds lanes (0,4,1,5), hB lanes (2,6,3,7): permute the keep masks the same way
    hA &= prmt(keep0, keep1, 0x5140u);
    hB &= prmt(keep0, keep1, 0x7362u);
}

__global__ void __launch_bounds__(256)
clock_measure_kernel(const uint4* __restrict__ c0, const uint4* __restrict__ c1, int64_t nvec, int64_t H,
                     int64_t p0, int64_t off_left0, int64_t off_down0, int64_t off_left1, int64_t off_down1,
                     int64_t L, int64_t Nc, int64_t ptail, uint32_t q, unsigned long long* acc)
{
    __shared__ unsigned long long sacc[3 * 64];
    for (int i = threadIdx.x; i < 3 * 64; i += blockDim.x) sacc[i] = 0;
    __syncthreads();
    uint32_t cnt[3][8];
#pragma unroll
    for (int h = 0; h < 3; ++h)
#pragma unroll
        for (int c = 0; c < 8; ++c) cnt[h][c] = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        const int64_t p = p0 + v;
        uint32_t keep[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
        if (p >= ptail) {
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                keep[w] = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if ((int64_t)(4 * w + j) * L + p < Nc) keep[w] |= 0xFFu << (8 * j);
            }
        }
#pragma unroll
        for (int colour = 0; colour < 2; ++colour) {
            const uint4* own = (colour ? c1 : c0) + H + v;
            const uint4* oth = (colour ? c0 : c1) + H + v;
            const uint4 ov = *own, lv = *(oth + (colour ? off_left1 : off_left0)), dv = *(oth + (colour ? off_down1 : off_down0));
            const uint32_t ow[4] = {ov.x, ov.y, ov.z, ov.w}, lw[4] = {lv.x, lv.y, lv.z, lv.w}, dw[4] = {dv.x, dv.y, dv.z, dv.w};
            if (q <= 8) {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    uint32_t hs[2], hl[2], hd[2];
                    onehot8(ow[2 * g], ow[2 * g + 1], keep[2 * g], keep[2 * g + 1], hs[0], hs[1]);
                    onehot8(submod_q(lw[2 * g], ow[2 * g], q), submod_q(lw[2 * g + 1], ow[2 * g + 1], q),
                            keep[2 * g], keep[2 * g + 1], hl[0], hl[1]);
                    onehot8(submod_q(dw[2 * g], ow[2 * g], q), submod_q(dw[2 * g + 1], ow[2 * g + 1], q),
                            keep[2 * g], keep[2 * g + 1], hd[0], hd[1]);
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const uint32_t m = 0x01010101u << c;
                        cnt[0][c] += __popc(hs[0] & m) + __popc(hs[1] & m);
                        cnt[1][c] += __popc(hl[0] & m) + __popc(hl[1] & m);
                        cnt[2][c] += __popc(hd[0] & m) + __popc(hd[1] & m);
                    }
                }
            } else {
#pragma unroll
                for (int b = 0; b < 16; ++b) {
                    if (!((keep[b >> 2] >> (8 * (b & 3))) & 1u)) continue;
                    const uint32_t s = (ow[b >> 2] >> (8 * (b & 3))) & 0xFFu;
                    const uint32_t l = (lw[b >> 2] >> (8 * (b & 3))) & 0xFFu;
                    const uint32_t d = (dw[b >> 2] >> (8 * (b & 3))) & 0xFFu;
                    uint32_t dl = l + q - s; if (dl >= q) dl -= q;
                    uint32_t dd = d + q - s; if (dd >= q) dd -= q;
                    atomicAdd(&sacc[s], 1ull);
                    atomicAdd(&sacc[64 + dl], 1ull);
                    atomicAdd(&sacc[128 + dd], 1ull);
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
        if (sacc[i]) atomicAdd(&acc[i], sacc[i]);
}
