// XY 2D with helical boundary (module xy2d_gpu_m, src/xy2d_gpu_m.f90): kernels, host-side handle and C ABI.
// Reference: type(xy2d_gpu) :12-43 -- the XY analogue of ising2d_gpu_m: one linear array
// spins(1-nx : nall+nx, 1:2) real64 (cos, sin) with a one-row halo, colour = parity of the linear index
// (valid iff nx odd, ny even), Metropolis (:138-174) and over-relaxation without
// renormalisation (:176-213), E and Mx sums (:259-291).
//
// Layout here: one fp32 angle in turns per site, the two colours in separate ring arrays of Nc = nall/2
// sites (colour-c site k = i >> 1, i = idx - 1).  The neighbours of colour-c site k are the other colour's
// sites k-1+c, k+c, k+h+c, k-h-1+c (mod Nc), nx = 2h+1; the ring wrap is index arithmetic (no halo).
// A thread owns 4 consecutive colour sites.  Uniforms: Philox in registers, contract in
// oracle/rng_contract.c (orc_xyh_uniforms).
#include <math.h>
#include <stdlib.h>
#include <new>
#include "../../include/b200mc.h"
#include "common.cuh"

namespace {

#define TWO_PI_F 6.283185307179586f
#define INV_TWO_PI_F 0.15915494309189535f
#define TAG_XYH 0x5859484Cu

struct XYHArgs {
    float* own;
    const float* oth;
    int nc, h, colour;
    float beta;
    uint64_t draw;
    uint32_t rk0[10];
};

__device__ __forceinline__ void sincos_turns(float t, float& s, float& c)
{
    const float r = t - rintf(t);
    __sincosf(turns_to_mufu_arg_centred(r), &s, &c);
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

__device__ __forceinline__ int wrap(int k, int nc) { return k < 0 ? k + nc : (k >= nc ? k - nc : k); }

// local field of colour-c site k: s(idx-1) + s(idx+1) + s(idx+nx) + s(idx-nx)   (:246-250, same order)
__device__ __forceinline__ void xyh_field(const XYHArgs& a, int k, float& hx, float& hy)
{
    const int c = a.colour, nc = a.nc;
    float s, co;
    // (stored angles lie in [0, 1] turns: no range reduction, common.cuh)
    sincos_unit(a.oth[wrap(k - 1 + c, nc)], s, co); hx = co; hy = s;
    sincos_unit(a.oth[wrap(k + c, nc)], s, co); hx += co; hy += s;
    sincos_unit(a.oth[wrap(k + a.h + c, nc)], s, co); hx += co; hy += s;
    sincos_unit(a.oth[wrap(k - a.h - 1 + c, nc)], s, co); hx += co; hy += s;
}

template <bool OVERRELAX>
__global__ void __launch_bounds__(256)
xyh_pass_kernel(const __grid_constant__ XYHArgs a)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int k0 = 4 * g;
    if (k0 >= a.nc) return;
    uint4 R[2];
    if (!OVERRELAX) {
        R[0] = philox_tag<TAG_XYH>(mk_ctr((uint64_t)g, a.draw, (uint32_t)a.colour, 0u), a.rk0);
        R[1] = philox_tag<TAG_XYH>(mk_ctr((uint64_t)g, a.draw, (uint32_t)a.colour, 1u), a.rk0);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int k = k0 + j;
        if (k >= a.nc) break;
        float hx, hy;
        xyh_field(a, k, hx, hy);
        const float t0 = a.own[k];
        if (OVERRELAX) {
            // over_relaxation_sub, :198-213: s <- 2 (h^ . s) h^ - s, i.e. theta <- 2 phi - theta
            const float t = 2.0f * atan2_turns(hy, hx) - t0;
            a.own[k] = frac_turns(t);
        } else {
            // update_sub, :157-174: accept iff r <= exp(-beta dE), dE = -(cand - s) . h
            const uint4 Rj = R[j >> 1];
            const uint32_t Ur = (j & 1) ? Rj.z : Rj.x, Uc = (j & 1) ? Rj.w : Rj.y;
            const float r = ((float)Ur + 1.0f) * 0x1p-32f;
            const float ct = ((float)Uc + 1.0f) * 0x1p-32f;
            float cs, cc, ss, sc;
            sincos_unit(ct, cs, cc);
            sincos_unit(t0, ss, sc);
            const float de = -((cc - sc) * hx + (cs - ss) * hy);
            if (!(r > __expf(-a.beta * de))) a.own[k] = ct;
        }
    }
}

// E = -sum_i s(i) . (s(i+1) + s(i+nx)) (:270-274), Mx = sum cos (:286-289); real64 accumulation
__global__ void __launch_bounds__(256)
xyh_measure_kernel(const float* __restrict__ c0, const float* __restrict__ c1, int nc, int h, double* acc)
{
    double part[2] = {0.0, 0.0};
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nc; k += gridDim.x * blockDim.x) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const float* own = c ? c1 : c0;
            const float* oth = c ? c0 : c1;
            float s, co, rs, rc, us, uc;
            sincos_turns(own[k], s, co);
            sincos_turns(oth[wrap(k + c, nc)], rs, rc);        // idx + 1
            sincos_turns(oth[wrap(k + h + c, nc)], us, uc);    // idx + nx
            part[0] -= (double)(co * (rc + uc) + s * (rs + us));
            part[1] += (double)co;
        }
    }
    block_atomic_add_f64<2>(acc, part);
}

// set_random_spin_sub (:97-105): theta = 2 pi u
__global__ void __launch_bounds__(256)
xyh_random_kernel(float* own, int nc, int colour, uint32_t seed, uint64_t draw)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (4 * g >= nc) return;
    const uint4 R = philox4x32_10(mk_ctr((uint64_t)g, draw, (uint32_t)colour, 0u), make_uint2(seed, TAG_INIT));
    const uint32_t rr[4] = {R.x, R.y, R.z, R.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (4 * g + j < nc) own[4 * g + j] = ((float)rr[j] + 1.0f) * 0x1p-32f;
}

__global__ void xyh_fill_kernel(float* a, float* b, int n, float v)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { a[i] = v; b[i] = v; }
}

// spins() in the reference layout spins(1-nx : nall+nx, 1:2) real64: out[j + len*(k-1)], j = idx - 1 + nx
__global__ void xyh_export_kernel(const float* c0, const float* c1, long long nall, int nx, double* out)
{
    const long long len = nall + 2LL * nx;
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= len) return;
    long long i = j - nx;                      // 0-based site index, halo rows wrap around the ring
    if (i < 0) i += nall;
    if (i >= nall) i -= nall;
    const float t = ((i & 1) ? c1 : c0)[i >> 1];
    double s, c;
    sincospi(2.0 * (double)t, &s, &c);
    out[j] = c;
    out[len + j] = s;
}
__global__ void xyh_export_turns_kernel(const float* c0, const float* c1, long long nall, float* out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nall) out[i] = ((i & 1) ? c1 : c0)[i >> 1];
}
__global__ void xyh_import_turns_kernel(float* c0, float* c1, long long nall, const float* in)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nall) { const float t = in[i]; ((i & 1) ? c1 : c0)[i >> 1] = t - floorf(t); }   // stored angles live in [0, 1] turns
}

struct XYH {
    int64_t nx, ny, nall;
    int nc, h;
    float* c[2];
    void* stage;
    double* d_acc;
    cudaStream_t stream;
    double beta;
    uint32_t seed;
    uint64_t draw;
    bool obs_valid;
    double obs[2];
    int sms;
};

void fill_args(XYH* m, int colour, XYHArgs* a)
{
    a->own = m->c[colour]; a->oth = m->c[colour ^ 1];
    a->nc = m->nc; a->h = m->h; a->colour = colour; a->beta = (float)m->beta; a->draw = m->draw;
    for (int r = 0; r < 10; ++r) a->rk0[r] = m->seed + (uint32_t)r * PHILOX_W0;
}

template <bool OR>
int pass_pair(XYH* m)
{
    m->obs_valid = false;
    const int groups = (m->nc + 3) / 4;
    for (int colour = 0; colour < 2; ++colour) {  // offset 1 = odd idx = even i = colour 0 first (:146,:151)
        XYHArgs a;
        fill_args(m, colour, &a);
        COUNT_LAUNCH();
        xyh_pass_kernel<OR><<<(groups + 255) / 256, 256, 0, m->stream>>>(a);
        CK(cudaGetLastError());
    }
    return B200MC_OK;
}

int measure(XYH* m)
{
    if (m->obs_valid) return B200MC_OK;
    CK(cudaMemsetAsync(m->d_acc, 0, 2 * sizeof(double), m->stream));
    COUNT_LAUNCH();
    xyh_measure_kernel<<<m->sms * 8, 256, 0, m->stream>>>(m->c[0], m->c[1], m->nc, m->h, m->d_acc);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(m->obs, m->d_acc, 2 * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    m->obs_valid = true;
    return B200MC_OK;
}

void destroy(XYH* m)
{
    cudaStreamSynchronize(m->stream);
    cudaFree(m->c[0]); cudaFree(m->c[1]); cudaFree(m->stage); cudaFree(m->d_acc);
    delete m;
}

int fill(XYH* m, float v)
{
    m->obs_valid = false;
    COUNT_LAUNCH();
    xyh_fill_kernel<<<(m->nc + 255) / 256, 256, 0, m->stream>>>(m->c[0], m->c[1], m->nc, v);
    CK(cudaGetLastError());
    return B200MC_OK;
}

int ensure_stage(XYH* m)
{
    if (!m->stage) CK(cudaMalloc(&m->stage, (size_t)2 * (m->nall + 2 * m->nx) * sizeof(double)));
    return B200MC_OK;
}

}  // namespace

#define HXH(h) (reinterpret_cast<XYH*>(h))
#define CHECK_XH(h) do { if (!(h)) ARG_FAIL("invalid handle"); } while (0)

extern "C" {

int b200mc_xy2dh_create(void** out, int64_t nx, int64_t ny, double kbt, int32_t iseed)
{
    if (!out) ARG_FAIL("null handle pointer");
    *out = nullptr;
    if (!(kbt > 0.0)) ARG_FAIL("kbt must be > 0");
    // colour = parity of the linear index: a checkerboard only for nx odd, ny even
    if (nx < 3 || ny < 2 || (nx & 1) == 0 || (ny & 1)) ARG_FAIL("xy2d helical: need nx odd >= 3 and ny even >= 2 (got %lld x %lld)", (long long)nx, (long long)ny);
    if (nx * ny >= (int64_t)0x7FFFFFF0) ARG_FAIL("xy2d helical: lattice too large");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        snprintf(g_b200mc_err, sizeof(g_b200mc_err), "no CUDA device: this library has no CPU fallback");
        return B200MC_ERR_CUDA;
    }
    XYH* m = new (std::nothrow) XYH();
    if (!m) ARG_FAIL("out of host memory");
    m->nx = nx; m->ny = ny; m->nall = nx * ny; m->nc = (int)(m->nall / 2); m->h = (int)((nx - 1) / 2);
    m->stream = 0; m->beta = 1 / kbt; m->seed = (uint32_t)iseed; m->draw = 0; m->obs_valid = false;
    m->c[0] = m->c[1] = nullptr; m->stage = nullptr; m->d_acc = nullptr;
    int dev = 0; m->sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&m->sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaMalloc(&m->c[0], (size_t)m->nc * sizeof(float)) != cudaSuccess || cudaMalloc(&m->c[1], (size_t)m->nc * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&m->d_acc, 2 * sizeof(double)) != cudaSuccess) {
        snprintf(g_b200mc_err, sizeof(g_b200mc_err), "cudaMalloc failed");
        destroy(m); return B200MC_ERR_CUDA;
    }
    int rc = fill(m, 0.0f);  // set_allup_spin
    if (rc) { destroy(m); return rc; }
    *out = m;
    return B200MC_OK;
}
int b200mc_xy2dh_destroy(void* h) { if (h) destroy(HXH(h)); return B200MC_OK; }
int b200mc_xy2dh_set_stream(void* h, void* s) { CHECK_XH(h); HXH(h)->stream = (cudaStream_t)s; return B200MC_OK; }
int b200mc_xy2dh_skip_curand(void* h, int64_t n)
{
    CHECK_XH(h);
    if (n < 0) ARG_FAIL("n_skip < 0");
    HXH(h)->draw += (uint64_t)((n + HXH(h)->nall - 1) / HXH(h)->nall);
    return B200MC_OK;
}
int b200mc_xy2dh_set_allup_spin(void* h) { CHECK_XH(h); return fill(HXH(h), 0.0f); }
int b200mc_xy2dh_set_random_spin(void* h)
{
    CHECK_XH(h);
    XYH* m = HXH(h);
    m->obs_valid = false;
    const int groups = (m->nc + 3) / 4;
    for (int c = 0; c < 2; ++c) {
        COUNT_LAUNCH();
        xyh_random_kernel<<<(groups + 255) / 256, 256, 0, m->stream>>>(m->c[c], m->nc, c, m->seed, m->draw);
        CK(cudaGetLastError());
    }
    m->draw += 1;
    return B200MC_OK;
}
int b200mc_xy2dh_set_kbt(void* h, double kbt) { CHECK_XH(h); if (!(kbt > 0.0)) ARG_FAIL("kbt must be > 0"); HXH(h)->beta = 1 / kbt; return B200MC_OK; }
int b200mc_xy2dh_set_beta(void* h, double beta) { CHECK_XH(h); HXH(h)->beta = beta; return B200MC_OK; }
int b200mc_xy2dh_update(void* h) { CHECK_XH(h); int rc = pass_pair<false>(HXH(h)); HXH(h)->draw += 1; return rc; }
int b200mc_xy2dh_update_n(void* h, int32_t n) { CHECK_XH(h); for (int i = 0; i < n; ++i) { int rc = b200mc_xy2dh_update(h); if (rc) return rc; } return B200MC_OK; }
int b200mc_xy2dh_update_over_relaxation(void* h, int32_t n_steps) { CHECK_XH(h); for (int i = 0; i < n_steps; ++i) { int rc = pass_pair<true>(HXH(h)); if (rc) return rc; } return B200MC_OK; }
int b200mc_xy2dh_calc_energy_sum(void* h, double* e) { CHECK_XH(h); int rc = measure(HXH(h)); if (rc) return rc; *e = HXH(h)->obs[0]; return B200MC_OK; }
int b200mc_xy2dh_calc_magne_sum(void* h, double* m) { CHECK_XH(h); int rc = measure(HXH(h)); if (rc) return rc; *m = HXH(h)->obs[1]; return B200MC_OK; }
int b200mc_xy2dh_get_spins(void* h, double* out)
{
    CHECK_XH(h);
    if (!out) ARG_FAIL("null output");
    XYH* m = HXH(h);
    int rc = ensure_stage(m);
    if (rc) return rc;
    const long long len = m->nall + 2 * m->nx;
    COUNT_LAUNCH();
    xyh_export_kernel<<<(unsigned)((len + 255) / 256), 256, 0, m->stream>>>(m->c[0], m->c[1], m->nall, (int)m->nx, reinterpret_cast<double*>(m->stage));
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, m->stage, (size_t)2 * len * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    return B200MC_OK;
}
int b200mc_xy2dh_get_angles(void* h, float* out)
{
    CHECK_XH(h);
    if (!out) ARG_FAIL("null output");
    XYH* m = HXH(h);
    int rc = ensure_stage(m);
    if (rc) return rc;
    COUNT_LAUNCH();
    xyh_export_turns_kernel<<<(unsigned)((m->nall + 255) / 256), 256, 0, m->stream>>>(m->c[0], m->c[1], m->nall, reinterpret_cast<float*>(m->stage));
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, m->stage, (size_t)m->nall * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    return B200MC_OK;
}
int b200mc_xy2dh_set_angles(void* h, const float* in)
{
    CHECK_XH(h);
    if (!in) ARG_FAIL("null input");
    XYH* m = HXH(h);
    int rc = ensure_stage(m);
    if (rc) return rc;
    m->obs_valid = false;
    CK(cudaMemcpyAsync(m->stage, in, (size_t)m->nall * sizeof(float), cudaMemcpyHostToDevice, m->stream));
    COUNT_LAUNCH();
    xyh_import_turns_kernel<<<(unsigned)((m->nall + 255) / 256), 256, 0, m->stream>>>(m->c[0], m->c[1], m->nall, reinterpret_cast<const float*>(m->stage));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(m->stream));
    return B200MC_OK;
}
int64_t b200mc_xy2dh_nx(void* h) { return h ? HXH(h)->nx : -1; }
int64_t b200mc_xy2dh_ny(void* h) { return h ? HXH(h)->ny : -1; }
int64_t b200mc_xy2dh_nall(void* h) { return h ? HXH(h)->nall : -1; }
double b200mc_xy2dh_kbt(void* h) { return h ? 1 / HXH(h)->beta : 0.0; }
double b200mc_xy2dh_beta(void* h) { return h ? HXH(h)->beta : 0.0; }
int b200mc_xy2dh_sync(void* h) { CHECK_XH(h); CK(cudaStreamSynchronize(HXH(h)->stream)); return B200MC_OK; }

}  // extern "C"
