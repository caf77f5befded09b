// Acceptance tables of the Ising models, built on the host exactly as the reference builds them, and their
// integer-threshold form for the kernels (shared by the helical modules, ising.cu, and the periodic one, ising_periodic.cu).
#pragma once
#include <math.h>
#include "ising_kernels.cuh"

struct IsingHostTables {
    double w[16];      // w[s*8 + S]: acceptance probability exactly as the reference builds it
    double exparr[17]; // 2D: exparr(-8:8)
    double ws3[14];    // 3D: ws(0:6, 0:1)
    IsingTab tab;
    IsingTabF64 tabf;
};

static inline int ising_build_host_tables(int ndim, int method, double beta, uint32_t seed, IsingHostTables* m)
{
    for (int i = 0; i < 16; ++i) m->w[i] = 0.0;
    if (ndim == 3) {
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
                m->ws3[s1 + s2 + 7 * 0] = fmin(1.0, exp(-beta * (double)(e2 - e1)));
                m->ws3[s1 + s2 + 7 * 1] = fmin(1.0, exp(-beta * (double)(e1 - e2)));
            }
        for (int s = 0; s < 2; ++s)
            for (int S = 0; S <= 6; ++S) m->w[s * 8 + S] = m->ws3[S + 7 * s];
    } else {
        // update_exparr_ising2d_gpu, src/ising2d_gpu_m.f90:122-131
        for (int i = 0; i < 17; ++i) m->exparr[i] = 1.0;
        for (int diff = 1; diff <= 8; ++diff) m->exparr[diff + 8] = exp(-beta * diff);
        // calc_delta_energy :195 with sigma = 2s-1 and sum(sigma_nb) = 2S-4
        for (int s = 0; s < 2; ++s)
            for (int S = 0; S <= 4; ++S) {
                const int de = 2 * (2 * s - 1) * (2 * S - 4);
                m->w[s * 8 + S] = m->exparr[de + 8];
            }
    }
    if (method == METHOD_HEATBATH) {
        // SURVEY Q10 definition: p_up(S) = 1/(1+exp(-2 beta (2S - z))), new spin = up iff u <= p_up
        const int z = ndim == 3 ? 6 : 4;
        for (int i = 0; i < 16; ++i) m->w[i] = 0.0;
        for (int S = 0; S <= z; ++S) {
            const double p = 1.0 / (1.0 + exp(-2.0 * beta * (double)(2 * S - z)));
            m->w[S] = p;
            m->w[8 + S] = p;
        }
    }
    // thresholds: u <= w  <=>  U < thr, thr = floor(w 2^32)
    // kernel table index: Metropolis k' = number of aligned neighbours (w(S, s) = w(k') for the
    // zero-field model: verified below), heat-bath S
    const int z = ndim == 3 ? 6 : 4;
    uint8_t tb[8];
    for (int idx = 0; idx < 8; ++idx) {
        double w = 0.0;
        if (idx <= z) {
            if (method == METHOD_METROPOLIS) {
                w = m->w[1 * 8 + idx];                      // s = 1: k' = S
                if (m->w[0 * 8 + (z - idx)] != w) ARG_FAIL("internal: acceptance table is not symmetric");
            } else {
                w = m->w[idx];
            }
        }
        uint64_t thr = (uint64_t)floor(w * 4294967296.0);
        if (thr > 4294967296ull) thr = 4294967296ull;
        const uint32_t t7 = (uint32_t)(thr >> 25);  // 0..128
        tb[idx] = (uint8_t)(128u - t7);
        m->tab.low25[idx] = (uint32_t)(thr & 0x1FFFFFFu);
    }
    for (int i = 0; i < 16; ++i) m->tabf.w[i] = m->w[i];
    for (int r = 0; r < 10; ++r) m->tab.rk0[r] = seed + (uint32_t)r * PHILOX_W0;
    m->tab.tlo = tb[0] | (tb[1] << 8) | (tb[2] << 16) | ((uint32_t)tb[3] << 24);
    m->tab.thi = tb[4] | (tb[5] << 8) | (tb[6] << 16) | ((uint32_t)tb[7] << 24);
    return B200MC_OK;
}

