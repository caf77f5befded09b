// Shared device/host helpers for the B200 (sm_100a) spin Monte Carlo library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define B200MC_OK 0
#define B200MC_ERR_ARG 1      // invalid argument / invalid lattice shape
#define B200MC_ERR_CUDA 2     // a CUDA runtime call failed (see b200mc_last_error)
#define B200MC_ERR_STATE 3    // call not valid in the current state
#define B200MC_ERR_UNSUPPORTED 4

extern thread_local char g_b200mc_err[512];
extern unsigned long long g_b200mc_launches;  // kernels launched by this library (all handles)
#define COUNT_LAUNCH() (++g_b200mc_launches)

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            snprintf(g_b200mc_err, sizeof(g_b200mc_err), "%s:%d: %s -> %s", __FILE__,     \
                     __LINE__, #call, cudaGetErrorString(e_));                            \
            return B200MC_ERR_CUDA;                                                       \
        }                                                                                 \
    } while (0)

#define ARG_FAIL(...)                                                  \
    do {                                                               \
        snprintf(g_b200mc_err, sizeof(g_b200mc_err), __VA_ARGS__);     \
        return B200MC_ERR_ARG;                                         \
    } while (0)

// ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11), in registers.  Ten rounds of two
// 32x32->64 multiplies (IMAD.WIDE.U32) and two three-input XORs (LOP3); the
// round keys are uniform across the grid so their schedule costs nothing per
// thread.  Pinned to the Random123 known-answer vectors by
// tests/test_gpu_rng.py (through b200mc_debug_philox).
// ---------------------------------------------------------------------------
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

__device__ __forceinline__ void mulwide(uint32_t a, uint32_t b, uint32_t& lo, uint32_t& hi)
{
    asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0,%1}, t;\n\t}"
        : "=r"(lo), "=r"(hi)
        : "r"(a), "r"(b));
}

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t lo0, hi0, lo1, hi1;
        mulwide(PHILOX_M0, c.x, lo0, hi0);
        mulwide(PHILOX_M1, c.z, lo1, hi1);
        uint4 n;
        n.x = hi1 ^ c.y ^ k.x;
        n.y = lo1;
        n.z = hi0 ^ c.w ^ k.y;
        n.w = lo0;
        c = n;
        k.x += PHILOX_W0;
        k.y += PHILOX_W1;
    }
    return c;
}

// RNG contract (see DESIGN.md "RNG contract"; CPU restatement in
// oracle/rng_contract.c): key = (seed, TAG), counter =
// (block_lo, block_hi, draw_lo, draw[32..47] | colour << 16 | sub << 24).
#define TAG_ISING 0x49534E47u
#define TAG_INIT 0x494E4954u
#define TAG_CLOCK 0x434C4F4Bu
#define TAG_TORUS 0x544F5253u
#define TAG_XY 0x58593244u

__device__ __forceinline__ uint4 mk_ctr(uint64_t blk, uint64_t draw, uint32_t colour,
                                                 uint32_t sub)
{
    uint4 c;
    c.x = (uint32_t)blk;
    c.y = (uint32_t)(blk >> 32);
    c.z = (uint32_t)draw;
    c.w = (uint32_t)((draw >> 32) & 0xFFFFu) | (colour << 16) | (sub << 24);
    return c;
}

// ---------------------------------------------------------------------------
// 128-bit global memory access with cache hints.
//   ld_other : the colour that is read-only during this pass.  Neighbouring
//              threads / later rows re-read it, so it goes through L1 (nc path).
//   ld_own / st_own : streamed exactly once per pass -> no L1 allocation.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint4 ld_other(const uint4* p)
{
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
// the same colour when other blocks wrote it earlier in THIS launch (cooperative multi-pass kernel): the nc path may
// serve stale L1 lines, so read at L2
__device__ __forceinline__ uint4 ld_other_coherent(const uint4* p)
{
    uint4 r;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint4 ld_own(const uint4* p, uint64_t pol)
{
    uint4 r;
#ifdef OWN_PLAIN
    asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
#endif
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void st_own(uint4* p, uint4 v, uint64_t pol)
{
#ifdef OWN_PLAIN
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    return;
#endif
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p),
                 "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol)
                 : "memory");
}

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// does the word contain a zero byte?
__device__ __forceinline__ uint32_t zero_byte_mask(uint32_t v)
{
    return (v - 0x01010101u) & ~v & 0x80808080u;
}

// warp + block sum of up to NV 64-bit values, one atomic per block per value
template <int NV>
__device__ __forceinline__ void block_atomic_add(unsigned long long* dst, long long (&v)[NV])
{
    __shared__ long long sm[NV][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[j] += __shfl_down_sync(0xffffffffu, v[j], o);
        if (lane == 0) sm[j][warp] = v[j];
    }
    __syncthreads();
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            long long t = (lane < nw) ? sm[j][lane] : 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
            if (lane == 0 && t != 0) atomicAdd(dst + j, (unsigned long long)t);
        }
    }
}
template <int NV>
__device__ __forceinline__ void block_atomic_add_f64(double* dst, double (&v)[NV])
{
    __shared__ double smd[NV][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[j] += __shfl_down_sync(0xffffffffu, v[j], o);
        if (lane == 0) smd[j][warp] = v[j];
    }
    __syncthreads();
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            double t = (lane < nw) ? smd[j][lane] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
            if (lane == 0) atomicAdd(dst + j, t);
        }
    }
}

// ---------------------------------------------------------------------------
// XY models: angles are stored as fp32 turns in [0, 1]
// ---------------------------------------------------------------------------
// cos / sin of an angle known to lie in [0, 1] turns (every producer of stored angles keeps them there, and a
// candidate is a uniform in (0, 1]): sin.approx / cos.approx keep their 2^-20.5 absolute error on [-2 pi, 2 pi], so no
// range reduction is needed -- 1 FMUL + 2 MUFU (+ the FMUL.RZ inside the approximation).
// The argument MUFU sees is RZ(RN(t * c) * k) with c = float(2 pi) (+2.8e-8 relative), k = float(1 / 2 pi) (-4.0e-8) and a
// round-toward-zero product (-3e-8 on average): with the plain constant every angle is evaluated 5.4e-8 (relative) too
// small, a SYSTEMATIC rotation that adds up over the lattice (sum cos off by -4.9e-8 N; seen as a 4e-5 relative error of M
// at 1024 x 512).  turns_to_mufu_arg folds the compensation into the multiply: t * (c + 3.1e-7), one rounding (measured on the
// device over a uniform grid of 2^24 angles: mean error of cos -3.8e-8 -> -5e-10, tools/probes/sfu_bias.cu).
#define XY_TWO_PI_HI 6.283185307179586f
#define XY_TWO_PI_LO 3.1e-7f
__device__ __forceinline__ float turns_to_mufu_arg(float t) { return fmaf(t, XY_TWO_PI_HI, t * XY_TWO_PI_LO); }
// the same for an argument centred to [-1/2, 1/2] turns (the measurement kernels): the shrink towards zero is symmetric there
// and the residual is MUFU's own; measured on the device (tools/probes/sfu_bias.cu, profiles/r02c_sfu_bias.log): mean error of
// cos +1.8e-8 with the plain constant, -5.3e-8 with + 3.1e-7, zero crossing at + 7.7e-8
#define XY_TWO_PI_LO_CENTRED 7.7e-8f
__device__ __forceinline__ float turns_to_mufu_arg_centred(float t) { return fmaf(t, XY_TWO_PI_HI, t * XY_TWO_PI_LO_CENTRED); }
__device__ __forceinline__ void sincos_unit(float t, float& s, float& c)
{
    __sincosf(turns_to_mufu_arg(t), &s, &c);
}

// atan2(y, x) / (2 pi) in [-1/2, 1/2]: octant reduction, q = min / max by MUFU.RCP, odd polynomial q P(q^2)
// (Chebyshev fit of atan(q) / (2 pi q) on [0, 1], max error 3e-8 turns evaluated in fp32 -- half an ulp of the
// stored angle near 1/2), about half the instructions of atan2f.  atan2(0, 0) = 0 like atan2f.
__device__ __forceinline__ float atan2_turns(float y, float x)
{
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float q = __fdividef(mn, fmaxf(mx, 1e-30f));
    const float s2 = q * q;
    float p = -0.0007430583355017006f;
    p = fmaf(p, s2, 0.0038461685180664062f);
    p = fmaf(p, s2, -0.009448567405343056f);
    p = fmaf(p, s2, 0.015766043215990067f);
    p = fmaf(p, s2, -0.022308088839054108f);
    p = fmaf(p, s2, 0.03178202360868454f);
    p = fmaf(p, s2, -0.05304946005344391f);
    p = fmaf(p, s2, 0.15915492177009583f);
    float r = p * q;                       // [0, 1/8]
    r = ay > ax ? 0.25f - r : r;
    r = x < 0.0f ? 0.5f - r : r;
    return y < 0.0f ? -r : r;
}

// t - floor(t) for |t| < 2^22, result in [0, 1]: t - rint(t) with the 1.5 * 2^23 add/subtract pair (two FADD instead of
// FRND, which shares the XU pipe with MUFU), then one conditional add
__device__ __forceinline__ float frac_turns(float t)
{
    const float fr = t - __fsub_rn(__fadd_rn(t, 12582912.0f), 12582912.0f);   // [-1/2, 1/2]
    return fr < 0.0f ? fr + 1.0f : fr;
}

