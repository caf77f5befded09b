// q-state clock, helical (clock_gpu_m / clock_gpu_multi_m): host-side handle and C ABI.
// Reference: type(clock_gpu) src/clock_gpu_m.f90:13-47; batched twin src/clock_gpu_multi_m.f90:13-48.
#include <math.h>
#include <stdlib.h>
#include <map>
#include <new>
#include <vector>
#include "../../include/b200mc.h"
#include "clock_kernels.cuh"
#include "ring.cuh"

namespace {

struct Clock {
    int64_t nx, ny;
    int32_t q, n_multi;
    bool multi;  // clock_gpu_multi_m semantics: strict comparator, array-valued observables
    std::vector<RingStore> st;
    cudaStream_t stream;
    double beta;
    uint32_t seed;
    uint64_t draw;
    std::vector<double> magne, etab, ws;   // host tables exactly as the reference builds them
    uint8_t* d_cls;        // q^6 class ids: bytes, or 16-bit when cls16
    int cls16, n_classes;
    uint64_t* d_thr;
    uint16_t* d_thr16;     // direct lookup table (q <= 6)
    int direct, grid_direct, smem_direct, threads_direct;
    int sample0;           // this handle holds samples sample0 .. sample0 + n_multi - 1 of the job (batch split across GPUs)
    double* d_ws;
    double* d_rand;
    double* d_next;
    unsigned long long* d_acc;  // per replica 3 x 64 counters
    int grid, smem_bytes, cls_in_smem;
    bool obs_valid;
    std::vector<long long> obs;  // n_multi x 192
};

int build_tables(Clock* m)
{
    const int q = m->q;
    const double pi = 4 * atan(1.0);
    const double psi = 2 * pi / q;  // this%pi_state_inv_, src/clock_gpu_m.f90:60
    const size_t q3 = (size_t)q * q * q, q6 = q3 * q3;
    m->magne.resize(q); m->etab.resize(q3); m->ws.resize(q6);
    for (int i = 0; i < q; ++i) m->magne[i] = cos(psi * i);  // :69
    // update_ws_clock_gpu, :105-146 (same loop nest, same expression order)
    for (int c = 0; c < q; ++c)
        for (int j = 0; j < q; ++j)
            for (int i = 0; i < q; ++i) {
                const double a = cos(psi * (i - c)), b = cos(psi * (j - c));
                m->etab[i + q * (j + q * c)] = -(a + b);
            }
#define ET(i, j, c) m->etab[(i) + q * ((j) + q * (c))]
    std::map<uint64_t, int> classes;
    std::vector<uint16_t> cls(q6);
    std::vector<uint64_t> thr;
    for (int ca = 0; ca < q; ++ca)
        for (int cb = 0; cb < q; ++cb)
            for (int l = 0; l < q; ++l)
                for (int k = 0; k < q; ++k)
                    for (int j = 0; j < q; ++j)
                        for (int i = 0; i < q; ++i) {
                            const double de = (ET(i, j, ca) + ET(k, l, ca)) - (ET(i, j, cb) + ET(k, l, cb));
                            const double w = (de <= 0.0) ? 1.0 : exp(-m->beta * de);
                            const size_t at = (size_t)i + (size_t)q * (j + (size_t)q * (k + (size_t)q * (l + (size_t)q * (cb + (size_t)q * ca))));
                            m->ws[at] = w;
                            // u = (U+1) 2^-32.  clock_gpu_m accepts iff u <= w  <=> U < floor(w 2^32);
                            // clock_gpu_multi_m accepts iff u < w  <=> U < ceil(w 2^32) - 1
                            const double x = w * 4294967296.0;
                            uint64_t t;
                            if (!m->multi) t = (uint64_t)floor(x);
                            else { const double c = ceil(x); t = c >= 1.0 ? (uint64_t)c - 1 : 0; }
                            auto it = classes.find(t);
                            int id;
                            if (it == classes.end()) {
                                id = (int)thr.size();
                                if (id >= CLOCK_MAX_CLASSES16) {
                                    snprintf(g_b200mc_err, sizeof(g_b200mc_err), "clock: more than %d distinct acceptance thresholds (q = %d)", CLOCK_MAX_CLASSES16, q);
                                    return B200MC_ERR_UNSUPPORTED;
                                }
                                classes[t] = id; thr.push_back(t);
                            } else id = it->second;
                            cls[at] = (uint16_t)id;
                        }
#undef ET
    m->n_classes = (int)thr.size();
    m->cls16 = m->n_classes > CLOCK_MAX_CLASSES ? 1 : 0;      // q >= 14: more than 256 distinct thresholds
    thr.resize(m->cls16 ? CLOCK_MAX_CLASSES16 : CLOCK_MAX_CLASSES, 0);
    if (m->d_thr16) {
        // direct table T[next][F] = thr >> 17 (0 .. 32768), F = up + q down + q^2 left + q^3 right + q^4 cur: the same order as
        // the class table (index = F + q^5 next)
        std::vector<uint16_t> t16(q6);
        for (size_t i = 0; i < q6; ++i) t16[i] = (uint16_t)(thr[cls[i]] >> 17);
        CK(cudaMemcpyAsync(m->d_thr16, t16.data(), q6 * sizeof(uint16_t), cudaMemcpyHostToDevice, m->stream));
        CK(cudaStreamSynchronize(m->stream));
    }
    std::vector<uint8_t> cls8;
    if (!m->cls16) {
        cls8.resize(q6);
        for (size_t i = 0; i < q6; ++i) cls8[i] = (uint8_t)cls[i];
        CK(cudaMemcpyAsync(m->d_cls, cls8.data(), q6, cudaMemcpyHostToDevice, m->stream));
    } else CK(cudaMemcpyAsync(m->d_cls, cls.data(), q6 * sizeof(uint16_t), cudaMemcpyHostToDevice, m->stream));
    CK(cudaMemcpyAsync(m->d_thr, thr.data(), thr.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, m->stream));
    if (m->d_ws) CK(cudaMemcpyAsync(m->d_ws, m->ws.data(), q6 * sizeof(double), cudaMemcpyHostToDevice, m->stream));
    CK(cudaStreamSynchronize(m->stream));  // host vectors go out of scope
    return B200MC_OK;
}

void fill_args(Clock* m, int j, int colour, ClockArgs* a)
{
    const RingGeom& g = m->st[j].g;
    a->r.own = m->st[j].vec[colour];
    a->r.oth = m->st[j].vec[colour ^ 1];
    a->r.nvec = g.L; a->r.H = g.H; a->r.p0 = 0;
    for (int t = 0; t < 6; ++t) a->r.off[t] = g.off[colour][t];
    a->r.seed = m->seed; a->r.colour = (uint32_t)colour; a->r.draw = m->draw;
    a->r.ticket = nullptr; a->r.chunk = 128;
    a->cls = m->d_cls; a->cls16 = m->cls16; a->thr = m->d_thr; a->q = (uint32_t)m->q;
    a->tab_bytes = (uint32_t)((size_t)m->q * m->q * m->q * m->q * m->q * m->q);
    a->replica = (uint32_t)(m->sample0 + j);
    for (int r = 0; r < 10; ++r) {
        a->rk0[r] = m->seed + (uint32_t)r * PHILOX_W0;
        a->rk1[r] = TAG_CLOCK + (uint32_t)(m->sample0 + j) + (uint32_t)r * PHILOX_W1;
    }
    a->cls_in_smem = m->cls_in_smem;
    a->thr16 = m->d_thr16;
}

int sweep(Clock* m)
{
    m->obs_valid = false;
    for (int colour = 0; colour < 2; ++colour) {
        for (int j = 0; j < m->n_multi; ++j) {
            ClockArgs a;
            fill_args(m, j, colour, &a);
            COUNT_LAUNCH();
            if (m->direct && m->q == 6 && m->threads_direct == 1024) clock_pass_direct_kernel<6, 1024><<<m->grid_direct, 1024, m->smem_direct, m->stream>>>(a);
            else if (m->direct && m->q == 6) clock_pass_direct_kernel<6><<<m->grid_direct, CLOCK_DIRECT_THREADS, m->smem_direct, m->stream>>>(a);
            else if (m->direct) clock_pass_direct_kernel<0><<<m->grid_direct, CLOCK_DIRECT_THREADS, m->smem_direct, m->stream>>>(a);
            else if (m->cls16) clock_pass_kernel<uint16_t><<<m->grid, 256, m->smem_bytes, m->stream>>>(a);
            else clock_pass_kernel<uint8_t><<<m->grid, 256, m->smem_bytes, m->stream>>>(a);
            CK(cudaGetLastError());
            int rc = ring_halo(&m->st[j], colour, m->stream);
            if (rc) return rc;
        }
    }
    m->draw += 1;
    return B200MC_OK;
}

int update_with_randoms(Clock* m, const double* randoms, const double* next_states)
{
    if (!randoms || !next_states) ARG_FAIL("null uniforms");
    const int64_t N = m->st[0].g.N;
    const size_t q6 = m->ws.size();
    if (!m->d_ws) {
        CK(cudaMalloc(&m->d_ws, q6 * sizeof(double)));
        CK(cudaMemcpy(m->d_ws, m->ws.data(), q6 * sizeof(double), cudaMemcpyHostToDevice));
        CK(cudaMalloc(&m->d_rand, (size_t)N * sizeof(double)));
        CK(cudaMalloc(&m->d_next, (size_t)N * sizeof(double)));
    }
    m->obs_valid = false;
    for (int j = 0; j < m->n_multi; ++j) {
        CK(cudaMemcpyAsync(m->d_rand, randoms + (size_t)j * N, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, m->stream));
        CK(cudaMemcpyAsync(m->d_next, next_states + (size_t)j * N, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, m->stream));
        const RingGeom& g = m->st[j].g;
        for (int colour = 0; colour < 2; ++colour) {
            ClockArgs a;
            fill_args(m, j, colour, &a);
            COUNT_LAUNCH();
            clock_pass_randoms_kernel<<<(unsigned)((g.L + 255) / 256), 256, 0, m->stream>>>(a, m->d_rand, m->d_next, m->d_ws, g.L, g.Nc, m->multi ? 1 : 0);
            CK(cudaGetLastError());
            int rc = ring_halo(&m->st[j], colour, m->stream);
            if (rc) return rc;
        }
        CK(cudaStreamSynchronize(m->stream));
    }
    return B200MC_OK;
}

int measure(Clock* m)
{
    if (m->obs_valid) return B200MC_OK;
    CK(cudaMemsetAsync(m->d_acc, 0, (size_t)m->n_multi * 192 * sizeof(unsigned long long), m->stream));
    for (int j = 0; j < m->n_multi; ++j) {
        const RingGeom& g = m->st[j].g;
        COUNT_LAUNCH();
        clock_measure_kernel<<<m->grid, 256, 0, m->stream>>>(m->st[j].vec[0], m->st[j].vec[1], g.L, g.H, 0, g.off[0][0], g.off[0][3],
                                                             g.off[1][0], g.off[1][3], g.L, g.Nc, g.ptail, (uint32_t)m->q, m->d_acc + (size_t)j * 192);
        CK(cudaGetLastError());
    }
    m->obs.resize((size_t)m->n_multi * 192);
    CK(cudaMemcpyAsync(m->obs.data(), m->d_acc, m->obs.size() * sizeof(long long), cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    m->obs_valid = true;
    return B200MC_OK;
}

void destroy(Clock* m)
{
    cudaStreamSynchronize(m->stream);
    for (auto& s : m->st) ring_free(&s);
    cudaFree(m->d_cls); cudaFree(m->d_thr); cudaFree(m->d_thr16); cudaFree(m->d_ws); cudaFree(m->d_rand); cudaFree(m->d_next); cudaFree(m->d_acc);
    delete m;
}

int create(void** out, int64_t nx, int64_t ny, double kbt, int32_t q, int32_t n_multi, bool multi, int32_t iseed)
{
    if (!out) ARG_FAIL("null handle pointer");
    *out = nullptr;
    if (!(kbt > 0.0)) ARG_FAIL("kbt must be > 0");
    if (q < 2) ARG_FAIL("state must be >= 2");
    // (the reference's nominal limit is 50, src/clock_gpu_m.f90:10, with a q^6 real64 table -- 125 GB at q = 50; here the host builds
    // the same table: 1.5 GB and a few seconds at q = 24)
    if (q > 24) { snprintf(g_b200mc_err, sizeof(g_b200mc_err), "clock: state = %d not supported (q^6 table built on the host; max 24)", q); return B200MC_ERR_UNSUPPORTED; }
    if (n_multi < 1) ARG_FAIL("n_multi must be >= 1");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        snprintf(g_b200mc_err, sizeof(g_b200mc_err), "no CUDA device: this library has no CPU fallback");
        return B200MC_ERR_CUDA;
    }
    Clock* m = new (std::nothrow) Clock();
    if (!m) ARG_FAIL("out of host memory");
    m->nx = nx; m->ny = ny; m->q = q; m->n_multi = n_multi; m->multi = multi; m->stream = 0;
    m->seed = (uint32_t)iseed; m->draw = 0; m->beta = 1 / kbt; m->obs_valid = false;
    m->d_cls = nullptr; m->d_thr = nullptr; m->d_thr16 = nullptr; m->direct = 0; m->sample0 = 0; m->d_ws = nullptr; m->d_rand = nullptr; m->d_next = nullptr; m->d_acc = nullptr;
    RingGeom g;
    int rc = ring_geom_init(&g, nx, ny, 0);
    if (rc) { delete m; return rc; }
    m->st.resize(n_multi);
    for (int j = 0; j < n_multi; ++j) {
        m->st[j].g = g; m->st[j].vec[0] = m->st[j].vec[1] = nullptr; m->st[j].stage = nullptr;
        rc = ring_alloc(&m->st[j]);
        if (rc) { destroy(m); return rc; }
        rc = ring_fill(&m->st[j], 0, m->stream);
        if (rc) { destroy(m); return rc; }
    }
    const size_t q6 = (size_t)q * q * q * q * q * q;
    if (cudaMalloc(&m->d_cls, (2 * q6 + 15) / 16 * 16) != cudaSuccess || cudaMalloc(&m->d_thr, CLOCK_MAX_CLASSES16 * sizeof(uint64_t)) != cudaSuccess ||
        cudaMalloc(&m->d_acc, (size_t)n_multi * 192 * sizeof(unsigned long long)) != cudaSuccess) {
        snprintf(g_b200mc_err, sizeof(g_b200mc_err), "cudaMalloc failed");
        destroy(m); return B200MC_ERR_CUDA;
    }
    // class table in shared memory when it fits (q <= 7), else read through L1/L2
    const size_t want = CLOCK_MAX_CLASSES * sizeof(uint64_t) + (q6 + 15) / 16 * 16;
    int dev = 0, sms = 148, maxsm = 0, occ = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    m->cls_in_smem = want <= (size_t)maxsm ? 1 : 0;
    m->smem_bytes = (int)(m->cls_in_smem ? want : CLOCK_MAX_CLASSES * sizeof(uint64_t));
    if (cudaFuncSetAttribute(clock_pass_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, m->smem_bytes) != cudaSuccess ||
        cudaFuncSetAttribute(clock_pass_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(CLOCK_MAX_CLASSES * sizeof(uint64_t))) != cudaSuccess) {
        snprintf(g_b200mc_err, sizeof(g_b200mc_err), "cudaFuncSetAttribute(smem) failed");
        destroy(m); return B200MC_ERR_CUDA;
    }
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, clock_pass_kernel<uint8_t>, 256, m->smem_bytes);
    if (occ < 1) occ = 1;
    const int64_t need = (g.L + 255) / 256;
    m->grid = (int)(need < (int64_t)sms * occ ? need : (int64_t)sms * occ);
    {   // direct lookup (q <= 6: q^3 < 256 for the byte-parallel index): 2 q^6 bytes of thresholds in shared memory
        const size_t wantd = (2 * q6 + 15) / 16 * 16;
        const char* t = getenv("B200MC_CLOCK_DIRECT");
        int occd = 0;
        if (q <= 6 && wantd <= (size_t)maxsm && !(t && atoi(t) == 0) &&
            cudaFuncSetAttribute(clock_pass_direct_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wantd) == cudaSuccess &&
            cudaFuncSetAttribute(clock_pass_direct_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wantd) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occd, clock_pass_direct_kernel<0>, CLOCK_DIRECT_THREADS, wantd) == cudaSuccess && occd >= 1 &&
            cudaMalloc(&m->d_thr16, (2 * q6 + 15) / 16 * 16) == cudaSuccess) {
            m->direct = 1; m->smem_direct = (int)wantd; m->threads_direct = CLOCK_DIRECT_THREADS;
            const char* tt = getenv("B200MC_CLOCK_THREADS");   // A/B: 1024-thread blocks (64 registers) for q = 6
            if (tt && atoi(tt) == 1024 && q == 6 &&
                cudaFuncSetAttribute(clock_pass_direct_kernel<6, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wantd) == cudaSuccess) m->threads_direct = 1024;
            const int64_t needd = (g.L + m->threads_direct - 1) / m->threads_direct;
            m->grid_direct = (int)(needd < (int64_t)sms * occd ? needd : (int64_t)sms * occd);
        } else cudaGetLastError();
    }
    rc = build_tables(m);
    if (rc) { destroy(m); return rc; }
    *out = m;
    return B200MC_OK;
}

int set_random(Clock* m)
{
    m->obs_valid = false;
    for (int j = 0; j < m->n_multi; ++j) {
        const RingGeom& g = m->st[j].g;
        for (int c = 0; c < 2; ++c) {
            COUNT_LAUNCH();
            clock_random_kernel<<<(unsigned)((g.L + 255) / 256), 256, 0, m->stream>>>(m->st[j].vec[c], g.L, g.H, 0, m->seed + 0x9E3779B9u * (uint32_t)(m->sample0 + j), m->draw, (uint32_t)c, (uint32_t)m->q);
            CK(cudaGetLastError());
        }
        int rc = ring_halo(&m->st[j], 0, m->stream);
        if (rc) return rc;
        rc = ring_halo(&m->st[j], 1, m->stream);
        if (rc) return rc;
    }
    m->draw += 1;
    return B200MC_OK;
}

}  // namespace

#define HC(h) (reinterpret_cast<Clock*>(h))
#define CHECK_C(h) do { if (!(h)) ARG_FAIL("invalid handle"); } while (0)

extern "C" {

int b200mc_clock_create(void** h, int64_t nx, int64_t ny, double kbt, int32_t state, int32_t iseed)
{
    return create(h, nx, ny, kbt, state, 1, false, iseed);
}
int b200mc_clock_multi_create(void** h, int64_t nx, int64_t ny, double kbt, int32_t state, int32_t n_multi, int32_t iseed)
{
    return create(h, nx, ny, kbt, state, n_multi, true, iseed);
}
int b200mc_clock_destroy(void* h) { if (h) destroy(HC(h)); return B200MC_OK; }
int b200mc_clock_set_stream(void* h, void* s) { CHECK_C(h); HC(h)->stream = (cudaStream_t)s; return B200MC_OK; }
int b200mc_clock_skip_curand(void* h, int64_t n)
{
    CHECK_C(h);
    if (n < 0) ARG_FAIL("n_skip < 0");
    const int64_t per = 2 * HC(h)->st[0].g.N * HC(h)->n_multi;  // uniforms drawn per update (src/clock_gpu_m.f90:188-189)
    HC(h)->draw += (uint64_t)((n + per - 1) / per);
    return B200MC_OK;
}
int b200mc_clock_set_allup_spin(void* h)
{
    CHECK_C(h);
    HC(h)->obs_valid = false;
    for (auto& s : HC(h)->st) { int rc = ring_fill(&s, 0, HC(h)->stream); if (rc) return rc; }
    return B200MC_OK;
}
int b200mc_clock_set_random_spin(void* h) { CHECK_C(h); return set_random(HC(h)); }
int b200mc_clock_set_beta(void* h, double beta) { CHECK_C(h); if (!(beta >= 0.0)) ARG_FAIL("beta must be >= 0"); HC(h)->beta = beta; return build_tables(HC(h)); }
int b200mc_clock_set_kbt(void* h, double kbt) { CHECK_C(h); if (!(kbt > 0.0)) ARG_FAIL("kbt must be > 0"); HC(h)->beta = 1 / kbt; return build_tables(HC(h)); }
int b200mc_clock_update(void* h) { CHECK_C(h); return sweep(HC(h)); }
int b200mc_clock_update_n(void* h, int32_t n) { CHECK_C(h); for (int i = 0; i < n; ++i) { int rc = sweep(HC(h)); if (rc) return rc; } return B200MC_OK; }
int b200mc_clock_update_with_randoms(void* h, const double* randoms, const double* next_states) { CHECK_C(h); return update_with_randoms(HC(h), randoms, next_states); }
int b200mc_clock_get_histograms(void* h, int64_t* hist, int64_t* bond_left, int64_t* bond_down)
{
    CHECK_C(h);
    Clock* m = HC(h);
    int rc = measure(m);
    if (rc) return rc;
    for (int j = 0; j < m->n_multi; ++j)
        for (int c = 0; c < m->q; ++c) {
            if (hist) hist[j * m->q + c] = m->obs[(size_t)j * 192 + c];
            if (bond_left) bond_left[j * m->q + c] = m->obs[(size_t)j * 192 + 64 + c];
            if (bond_down) bond_down[j * m->q + c] = m->obs[(size_t)j * 192 + 128 + c];
        }
    return B200MC_OK;
}
int b200mc_clock_calc_energy_sum(void* h, double* res)
{
    CHECK_C(h);
    Clock* m = HC(h);
    if (!res) ARG_FAIL("null output");
    int rc = measure(m);
    if (rc) return rc;
    const double pi = 4 * atan(1.0), psi = 2 * pi / m->q;
    for (int j = 0; j < m->n_multi; ++j) {
        double e = 0.0;
        for (int d = 0; d < m->q; ++d)
            e -= (double)(m->obs[(size_t)j * 192 + 64 + d] + m->obs[(size_t)j * 192 + 128 + d]) * cos(psi * d);
        res[j] = e;
    }
    return B200MC_OK;
}
int b200mc_clock_calc_magne_sum(void* h, double* res)
{
    CHECK_C(h);
    Clock* m = HC(h);
    if (!res) ARG_FAIL("null output");
    int rc = measure(m);
    if (rc) return rc;
    for (int j = 0; j < m->n_multi; ++j) {
        double s = 0.0;
        for (int c = 0; c < m->q; ++c) s += (double)m->obs[(size_t)j * 192 + c] * m->magne[c];
        res[j] = s;
    }
    return B200MC_OK;
}
int b200mc_clock_get_spins(void* h, int32_t* out)
{
    CHECK_C(h);
    if (!out) ARG_FAIL("null output");
    Clock* m = HC(h);
    const int64_t per = m->st[0].g.N + 2 * m->st[0].g.P;
    for (int j = 0; j < m->n_multi; ++j) { int rc = ring_export_i32(&m->st[j], out + (size_t)j * per, RING_MAP_IDENTITY, m->stream); if (rc) return rc; }
    return B200MC_OK;
}
int b200mc_clock_set_spins(void* h, const int32_t* in)
{
    CHECK_C(h);
    if (!in) ARG_FAIL("null input");
    Clock* m = HC(h);
    m->obs_valid = false;
    const int64_t per = m->st[0].g.N + 2 * m->st[0].g.P;
    for (int j = 0; j < m->n_multi; ++j) { int rc = ring_import_i32(&m->st[j], in + (size_t)j * per, RING_MAP_IDENTITY, m->stream, 0, m->q); if (rc) return rc; }
    return B200MC_OK;
}
int b200mc_clock_get_ws(void* h, double* out)
{
    CHECK_C(h);
    for (size_t i = 0; i < HC(h)->ws.size(); ++i) out[i] = HC(h)->ws[i];
    return B200MC_OK;
}
int64_t b200mc_clock_nx(void* h) { return h ? HC(h)->nx : -1; }
int64_t b200mc_clock_ny(void* h) { return h ? HC(h)->ny : -1; }
int64_t b200mc_clock_nall(void* h) { return h ? HC(h)->st[0].g.N : -1; }
int32_t b200mc_clock_state(void* h) { return h ? HC(h)->q : -1; }
int b200mc_clock_set_sample_offset(void* h, int32_t first_sample)
{
    CHECK_C(h);
    if (first_sample < 0) ARG_FAIL("first_sample must be >= 0");
    HC(h)->sample0 = first_sample;
    return B200MC_OK;
}
int32_t b200mc_clock_n_multi(void* h) { return h ? HC(h)->n_multi : -1; }
double b200mc_clock_kbt(void* h) { return h ? 1 / HC(h)->beta : 0.0; }
double b200mc_clock_beta(void* h) { return h ? HC(h)->beta : 0.0; }
int b200mc_clock_sync(void* h) { CHECK_C(h); CK(cudaStreamSynchronize(HC(h)->stream)); return B200MC_OK; }

}  // extern "C"
