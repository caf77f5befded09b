// Ising 2D / 3D (helical): host-side handle and C ABI.
// Reference: type(ising2d_gpu) src/ising2d_gpu_m.f90:12-42,
//            type(ising3d_gpu) src/ising3d_gpu_m.f90:15-48.
#include <math.h>
#include <stdlib.h>
#include <new>
#include <vector>
#include "../../include/b200mc.h"
#include "ising_kernels.cuh"
#include "ising_tables.cuh"
#include "ring.cuh"

namespace {

struct Ising {
    int ndim;  // 2 or 3
    int64_t nx, ny, nz;
    RingStore st;
    cudaStream_t stream;
    cudaStream_t comm_stream;          // slab mode: halo exchange overlapped with the interior launch
    cudaEvent_t ev_boundary, ev_halo;  // boundary vectors written / halo blocks received
    double beta;
    uint32_t seed;
    uint64_t draw;
    int method;
    double w[16];      // w[s*8 + S]: acceptance probability exactly as the reference builds it
    double exparr[17]; // 2D: exparr(-8:8)
    double ws3[14];    // 3D: ws(0:6, 0:1)
    IsingTab tab;
    IsingTabF64 tabf;
    unsigned long long* d_acc;  // [X, sum s]
    unsigned long long* h_acc;  // pinned host copy (device -> host read of every measurement)
    bool h_acc_pending;         // the cooperative sweep kernel stores its fused sums in h_acc itself: measure() only synchronises
    unsigned long long* acc_target;  // where the fused pass / the measure kernel add their sums (d_acc, or a slot of d_series)
    unsigned long long* d_series;    // run_relaxation: [mcs][2] sums, one slot per MCS
    int64_t series_cap;
    double* d_stats;                 // run_relaxation_stats: [mcs][10] Kahan sums + compensations
    int64_t stats_cap;
    int64_t* d_off1;            // colour-1 offsets for the measure kernel
    double* d_randoms;
    unsigned int* d_ticket;
    int tune;  // debug knobs from env B200MC_TUNE: bit0 = static round-robin (no ticket)
    int chunk; // vectors per ticket (env B200MC_CHUNK, default 128)
    int grid;
    int coop_grid;  // resident grid of the cooperative small-lattice sweep kernel (0: not available)
    int slab_nb;    // slab mode: blocks of the one-launch pass that start on the boundary tickets (env B200MC_SLAB_NB; 0 = from the boundary's share of the slab)
    int grid_push;  // resident grid of the fused update + halo-push kernel (its register budget differs)
    bool self_clean;   // ticket launches clear the other pass's counters / the fused sums themselves and store the sums to pinned host memory
                       // (env B200MC_SELF_CLEAN=0: memsets + copy, the round-1 form, for A/B)
    bool acc_zeroed;   // the first pass of this sweep has cleared d_acc for the fused second pass
    bool use_tma;   // single-GPU launches go through the copy-engine staged kernel
    int tma_grid;
    bool alive;
    // observables cache: valid until the configuration changes
    bool obs_valid;
    int64_t obs_e, obs_m;        // sample 0
    int n_multi;                 // batch of independent samples updated by the same launches (single GPU)
    std::vector<int64_t> obs_ev, obs_mv;   // all samples
    // fused measurement: when the caller measures after every update (the drivers' loop), the second
    // colour pass of the next sweep accumulates X and sum(s) itself and measure() only reads them back
    bool fuse_ok;        // layout allows it (no site-less tail positions)
    bool want_fused;     // the last sweep was followed by a measurement
    bool fused_pending;  // d_acc holds the sums of the current configuration
    // optional per-launch timing of the pass kernel (CUDA events on the handle's stream)
    bool timing;
    std::vector<cudaEvent_t> evs;  // event pool: pair (2i, 2i+1) brackets the i-th timed launch
    size_t ev_used;
};

int build_tables(Ising* m)
{
    IsingHostTables t;
    const int rc = ising_build_host_tables(m->ndim, m->method, m->beta, m->seed, &t);
    if (rc) return rc;
    for (int i = 0; i < 16; ++i) m->w[i] = t.w[i];
    for (int i = 0; i < 17; ++i) m->exparr[i] = t.exparr[i];
    for (int i = 0; i < 14; ++i) m->ws3[i] = t.ws3[i];
    m->tab = t.tab;
    m->tabf = t.tabf;
    return B200MC_OK;
}

template <int NNB>
int launch_range(Ising* m, int colour, int64_t vbeg, int64_t n, bool ordered, bool fuse, int fewer_blocks = 0, bool tickets_ready = false,
                 unsigned long long* acc_reset = nullptr, unsigned long long* host_out = nullptr)
{
    const RingGeom& g = m->st.g;
    RingPassArgs a = RingPassArgs();
    a.own = m->st.vec[colour] + vbeg;
    a.oth = m->st.vec[colour ^ 1] + vbeg;
    a.nvec = n;
    a.H = g.H;
    a.p0 = g.p0 + vbeg;
    for (int j = 0; j < 6; ++j) a.off[j] = g.off[colour][j];
    a.seed = m->seed;
    a.colour = (uint32_t)colour;
    a.draw = m->draw;
    a.ticket = nullptr;
    a.chunk = m->chunk;
    a.acc = m->acc_target;
    a.rstride = m->st.rstride;
    a.Lfold = g.L; a.Nc = g.Nc;
    a.mask_from = g.ptail < g.L ? (int)(g.ptail - a.p0 > 0 ? g.ptail - a.p0 : 0) : 0x7FFFFFFF;
    a.nopush = (m->tune & 32) ? 2 : 0;  // debug bit 5: no L2 prefetch
    int64_t need = (n + 255) / 256;
    int grid = (int)(need < (int64_t)m->grid ? need : (int64_t)m->grid);
    if (ordered && !(m->tune & 1) && n > (int64_t)m->grid * 256 * 4 && m->n_multi == 1) {
        a.ticket = m->d_ticket;
        if (m->self_clean && !tickets_ready && fewer_blocks == 0 && !m->use_tma) {
            // self-cleaning counters: set `colour` for this pass, set `colour ^ 1` cleared by the kernel (both clean at create)
            a.ticket = m->d_ticket + colour * TK_NCNT * 64;
            a.ticket_reset = m->d_ticket + (colour ^ 1) * TK_NCNT * 64;
            if (fuse && host_out) { a.host_out = host_out; a.done_warps = m->d_ticket + 2 * TK_NCNT * 64; m->h_acc_pending = true; }
            if (!fuse && acc_reset) { a.acc_reset = acc_reset; m->acc_zeroed = true; }
        } else if (!tickets_ready) CK(cudaMemsetAsync(m->d_ticket, 0, TK_NCNT * 64 * sizeof(unsigned int), m->stream));
        if (grid - fewer_blocks >= 1) grid -= fewer_blocks;  // slab mode: the other blocks of the resident set run ising_slab_kernel on the same tickets
    }
    COUNT_LAUNCH();
#define PASS(METHOD, ORD, MEAS) ising_pass_kernel<NNB, METHOD, ORD, false, MEAS><<<grid, 256, 0, m->stream>>>(a, m->tab)
#define BPASS(METHOD, MEAS) ising_pass_kernel<NNB, METHOD, false, false, MEAS, true><<<dim3(grid, m->n_multi), 256, 0, m->stream>>>(a, m->tab)
    if (m->n_multi > 1) {
        if (m->method == METHOD_METROPOLIS) { if (fuse) BPASS(METHOD_METROPOLIS, true); else BPASS(METHOD_METROPOLIS, false); }
        else { if (fuse) BPASS(METHOD_HEATBATH, true); else BPASS(METHOD_HEATBATH, false); }
    } else if (a.ticket && m->use_tma) {
        // copy-engine staging (ising_pass_tma_kernel): full tickets through cp.async.bulk, 2 blocks per SM
        const size_t smem = (size_t)TMA_STAGES * NNB * TMA_SLOT;
        const int tgrid = (int)(need < (int64_t)m->tma_grid ? need : (int64_t)m->tma_grid);  // 256 vectors per block and tile
#define TPASS(METHOD, MEAS) ising_pass_tma_kernel<NNB, METHOD, MEAS><<<tgrid, 288, smem, m->stream>>>(a, m->tab)
        if (m->method == METHOD_METROPOLIS) { if (fuse) TPASS(METHOD_METROPOLIS, true); else TPASS(METHOD_METROPOLIS, false); }
        else { if (fuse) TPASS(METHOD_HEATBATH, true); else TPASS(METHOD_HEATBATH, false); }
#undef TPASS
    } else if (m->method == METHOD_METROPOLIS) {
        if (a.ticket) { if (fuse) PASS(METHOD_METROPOLIS, true, true); else PASS(METHOD_METROPOLIS, true, false); }
        else { if (fuse) PASS(METHOD_METROPOLIS, false, true); else PASS(METHOD_METROPOLIS, false, false); }
    } else {
        if (a.ticket) { if (fuse) PASS(METHOD_HEATBATH, true, true); else PASS(METHOD_HEATBATH, true, false); }
        else { if (fuse) PASS(METHOD_HEATBATH, false, true); else PASS(METHOD_HEATBATH, false, false); }
    }
#undef BPASS
#undef PASS
    CK(cudaGetLastError());
    return B200MC_OK;
}

// slab mode with the direct transport: ONE launch per colour pass (update + halo push fused)
static void push_args(Ising* m, int colour, bool boundary_only, unsigned int* ticket, RingPassArgs& a)
{
    RingStore& st = m->st;
    const RingGeom& g = st.g;
    a.own = st.vec[colour];
    a.oth = st.vec[colour ^ 1];
    a.nvec = g.Lloc;
    a.H = g.H;
    a.p0 = g.p0;
    for (int j = 0; j < 6; ++j) a.off[j] = g.off[colour][j];
    a.seed = m->seed;
    a.colour = (uint32_t)colour;
    a.draw = m->draw;
    a.ticket = ticket;
    a.chunk = m->chunk;
    a.peer_lo = st.peer_vec[0][colour] + g.H + st.Lloc_prev;  // rank-1's high halo
    a.peer_hi = st.peer_vec[1][colour];                       // rank+1's low halo
    a.rot_lo = g.rank == 0 ? -1 : 0;              // the ring closes between rank 0 and rank P-1:
    a.rot_hi = g.rank == g.nranks - 1 ? +1 : 0;   // crossing the end of the fold moves a site to the next lane
    a.nb = (int)g.H;
    a.hi_start = (int)(g.Lloc - g.H);
    const int64_t nchunks = (g.Lloc + TK_CHUNK - 1) / TK_CHUNK;
    const int64_t blo = (g.H + TK_CHUNK - 1) / TK_CHUNK;      // chunks [0, blo) hold the first H owned vectors
    const int64_t jhi = (g.Lloc - g.H) / TK_CHUNK;            // chunks [jhi, nchunks) the last H
    a.nbchunks = (int)(blo + (nchunks - jhi));
    a.hi_tickets = (int)((nchunks - jhi) * TK_CHUNK);
    a.hi_first = (int)(jhi * TK_CHUNK);
    a.lo_end = (int)(blo * TK_CHUNK);
    a.nopush = ((m->tune & 4) ? 1 : 0) | ((m->tune & 32) ? 2 : 0);  // debug: bit 0 skip the NVLink stores (wrong results, timing only), bit 1 no L2 prefetch
    // boundary_only: just the tickets of the two boundary blocks (the interior follows as a plain launch)
    a.q_total = boundary_only ? a.nbchunks : (int)nchunks;
    a.dbg_wait = reinterpret_cast<unsigned long long*>(st.flags + 48);
    a.done = st.flags + 32;
    a.sig_prev = st.peer_flags[0] + 16;  // I am rank-1's "next"
    a.sig_next = st.peer_flags[1] + 0;   // and rank+1's "prev"
    a.wait_prev = st.flags + 0;
    a.wait_next = st.flags + 16;
    a.wait_seq = st.push_seq;
    a.sig_seq = ++st.push_seq;
    a.acc = m->acc_target;
    a.rstride = 0;
    a.Lfold = g.L; a.Nc = g.Nc; a.mask_from = 0x7FFFFFFF;   // slabs need Nc % 16 == 0: no tail
}

template <int NNB>
int launch_push(Ising* m, int colour, bool fuse, bool boundary_only, cudaStream_t stream, unsigned int* ticket)
{
    RingPassArgs a = RingPassArgs();
    push_args(m, colour, boundary_only, ticket, a);
    CK(cudaMemsetAsync(ticket, 0, TK_NCNT * 64 * sizeof(unsigned int), stream));
    COUNT_LAUNCH();
    if (m->method == METHOD_METROPOLIS) {
        if (fuse) ising_pass_kernel<NNB, METHOD_METROPOLIS, true, true, true><<<m->grid_push, 256, 0, stream>>>(a, m->tab);
        else ising_pass_kernel<NNB, METHOD_METROPOLIS, true, true, false><<<m->grid_push, 256, 0, stream>>>(a, m->tab);
    } else {
        if (fuse) ising_pass_kernel<NNB, METHOD_HEATBATH, true, true, true><<<m->grid_push, 256, 0, stream>>>(a, m->tab);
        else ising_pass_kernel<NNB, METHOD_HEATBATH, true, true, false><<<m->grid_push, 256, 0, stream>>>(a, m->tab);
    }
    CK(cudaGetLastError());
    return B200MC_OK;
}

// slab mode, direct transport, ONE launch per colour pass: the first blocks of the grid take the boundary tickets
// (update + NVLink push + flags) and then join the others on the interior, which runs the plain body
template <int NNB>
int launch_slab(Ising* m, int colour, bool fuse, int64_t vbeg, int64_t n)
{
    const RingGeom& g = m->st.g;
    RingPassArgs ab = RingPassArgs(), ai;
    push_args(m, colour, true, m->d_ticket + TK_NCNT * 64, ab);
    ai = ab;
    ai.own = m->st.vec[colour] + vbeg;
    ai.oth = m->st.vec[colour ^ 1] + vbeg;
    ai.nvec = n;
    ai.p0 = g.p0 + vbeg;
    ai.ticket = m->d_ticket;
    ai.nopush = (m->tune & 32) ? 2 : 0;
    ai.rstride = m->st.rstride;
    ai.mask_from = 0x7FFFFFFF;
    const int full = m->grid_push < m->grid ? m->grid_push : m->grid;
    // blocks that start on the boundary tickets: enough of them that the boundary (a slower code path with system-scope
    // fences, ~2.5x the time of an interior ticket) is done at about 40 % of the pass
    int nbb = m->slab_nb;
    if (nbb <= 0) {
        const double frac = (double)ab.nbchunks / (double)((g.Lloc + TK_CHUNK - 1) / TK_CHUNK);
        nbb = (int)(full * frac * 6.0 + 0.5);
        if (nbb < 8) nbb = 8;
        if (nbb > full / 2) nbb = full / 2;
    }
    if (nbb > full) nbb = full;
    if (nbb < 1) nbb = 1;
    const bool share = !(m->tune & (1024 | 1)) && m->comm_stream && n > (int64_t)m->grid * 256 * 4 && nbb < full;
    CK(cudaMemsetAsync(m->d_ticket, 0, 2 * TK_NCNT * 64 * sizeof(unsigned int), m->stream));
    // share: nbb blocks run the slab kernel (second stream), the rest of the resident set the plain kernel, both
    // kinds of block fit on the SMs together (same footprint) and take the interior tickets from the same counters
    cudaStream_t sb = m->stream;
    if (share) {
        CK(cudaEventRecord(m->ev_boundary, m->stream));
        CK(cudaStreamWaitEvent(m->comm_stream, m->ev_boundary, 0));
        sb = m->comm_stream;
    }
    const int grid = share ? nbb : full;
    COUNT_LAUNCH();
    if (m->method == METHOD_METROPOLIS) {
        if (fuse) ising_slab_kernel<NNB, METHOD_METROPOLIS, true><<<grid, 256, 0, sb>>>(ab, ai, m->tab, nbb);
        else ising_slab_kernel<NNB, METHOD_METROPOLIS, false><<<grid, 256, 0, sb>>>(ab, ai, m->tab, nbb);
    } else {
        if (fuse) ising_slab_kernel<NNB, METHOD_HEATBATH, true><<<grid, 256, 0, sb>>>(ab, ai, m->tab, nbb);
        else ising_slab_kernel<NNB, METHOD_HEATBATH, false><<<grid, 256, 0, sb>>>(ab, ai, m->tab, nbb);
    }
    CK(cudaGetLastError());
    if (share) {
        CK(cudaEventRecord(m->ev_halo, m->comm_stream));
        int rc = launch_range<NNB>(m, colour, vbeg, n, true, fuse, nbb, true);
        CK(cudaStreamWaitEvent(m->stream, m->ev_halo, 0));
        return rc;
    }
    return B200MC_OK;
}

// One colour pass + halo refresh.  Single GPU: one launch over the whole fold, then the halo kernel.
// Slab mode: the first and last H owned vectors (what the neighbouring ranks need) are updated first,
// their exchange runs on the comm stream while the interior launch runs on the compute stream.
template <int NNB>
int launch_pass(Ising* m, int colour, bool fuse, bool fuse_next = false)
{
    const RingGeom& g = m->st.g;
    m->obs_valid = false;
    m->fused_pending = false;
    m->h_acc_pending = false;
    if (colour == 0) m->acc_zeroed = false;
    if (fuse && m->acc_target == m->d_acc && !m->acc_zeroed) CK(cudaMemsetAsync(m->d_acc, 0, 2 * sizeof(unsigned long long) * m->n_multi, m->stream));
    m->acc_zeroed = false;
    // single GPU, one sample, sums in d_acc: the first pass clears them for the fused second pass, which stores the totals in h_acc
    const bool own_sums = g.nranks == 1 && !m->st.p2p && m->n_multi == 1 && m->acc_target == m->d_acc;
    unsigned long long* acc_reset = (own_sums && colour == 0 && fuse_next) ? m->d_acc : nullptr;
    unsigned long long* host_out = (own_sums && fuse) ? m->h_acc : nullptr;
    if (m->timing) {
        while (m->evs.size() < m->ev_used + 2) { cudaEvent_t e; CK(cudaEventCreate(&e)); m->evs.push_back(e); }
        CK(cudaEventRecord(m->evs[m->ev_used], m->stream));
    }
    int rc;
    if (g.nranks == 1 && !m->st.p2p) {
        rc = launch_range<NNB>(m, colour, 0, g.Lloc, true, fuse, 0, false, acc_reset, host_out);
        if (m->timing) { CK(cudaEventRecord(m->evs[m->ev_used + 1], m->stream)); m->ev_used += 2; }
        if (rc) return rc;
        return ring_halo(&m->st, colour, m->stream);
    }
    // (decided from the SMALLEST slab of the job, L / nranks, so that every rank takes the same transport even when
    // L % nranks != 0 makes the slabs differ by one vector)
    const bool split = g.L / g.nranks >= 4 * g.H && !(m->tune & 2);
    if (!split && g.nranks == 1) m->st.p2p = false;
    if (!split) {
        rc = launch_range<NNB>(m, colour, 0, g.Lloc, true, fuse);
        if (m->timing) { CK(cudaEventRecord(m->evs[m->ev_used + 1], m->stream)); m->ev_used += 2; }
        if (rc) return rc;
        return ring_halo(&m->st, colour, m->stream);
    }
    if (m->st.p2p) {
        // direct transport: one launch updates the whole slab, boundary chunks first, and stores their
        // results straight into the neighbours' halos over NVLink; the next pass waits (in the
        // kernel) for the neighbours' flags
        // The boundary tickets (first / last H owned vectors, ~3 % of a 1023 x 1023 x 1024 slab) run in the fused
        // update + push kernel; the interior -- which reads no halo cell -- runs in the plain kernel, whose
        // register budget and unrolled body are tuned for exactly that (B200MC_TUNE bit 8: everything in the fused kernel).
        const int64_t blo = (g.H + TK_CHUNK - 1) / TK_CHUNK, jhi = (g.Lloc - g.H) / TK_CHUNK;
        const bool two = jhi > blo + 64 && !(m->tune & 256) && m->comm_stream;
        if (two && !(m->tune & 512)) {
            rc = launch_slab<NNB>(m, colour, fuse, blo * TK_CHUNK, (jhi - blo) * TK_CHUNK);
        } else if (two) {
            // (B200MC_TUNE bit 9: the earlier two-launch form)
            // the two launches are independent of each other: the small boundary launch runs on the second stream,
            // concurrently with the interior launch; both see everything enqueued before this pass, and the compute
            // stream joins at the end
            CK(cudaEventRecord(m->ev_boundary, m->stream));
            CK(cudaStreamWaitEvent(m->comm_stream, m->ev_boundary, 0));
            rc = launch_push<NNB>(m, colour, fuse, true, m->comm_stream, m->d_ticket + TK_NCNT * 64);
            if (rc) return rc;
            CK(cudaEventRecord(m->ev_halo, m->comm_stream));
            rc = launch_range<NNB>(m, colour, blo * TK_CHUNK, (jhi - blo) * TK_CHUNK, true, fuse);
            CK(cudaStreamWaitEvent(m->stream, m->ev_halo, 0));
        } else {
            rc = launch_push<NNB>(m, colour, fuse, false, m->stream, m->d_ticket);
        }
        if (m->timing) { CK(cudaEventRecord(m->evs[m->ev_used + 1], m->stream)); m->ev_used += 2; }
        return rc;
    }
    if ((rc = launch_range<NNB>(m, colour, 0, g.H, false, fuse))) return rc;
    if ((rc = launch_range<NNB>(m, colour, g.Lloc - g.H, g.H, false, fuse))) return rc;
    CK(cudaEventRecord(m->ev_boundary, m->stream));
    CK(cudaStreamWaitEvent(m->comm_stream, m->ev_boundary, 0));
    if ((rc = ring_halo(&m->st, colour, m->comm_stream))) return rc;
    CK(cudaEventRecord(m->ev_halo, m->comm_stream));
    rc = launch_range<NNB>(m, colour, g.H, g.Lloc - 2 * g.H, true, fuse);
    if (m->timing) { CK(cudaEventRecord(m->evs[m->ev_used + 1], m->stream)); m->ev_used += 2; }
    if (rc) return rc;
    CK(cudaStreamWaitEvent(m->stream, m->ev_halo, 0));
    return B200MC_OK;
}

// Small lattices: n sweeps in one cooperative launch (ising_coop_kernel).  series: per-sweep sums (run_relaxation);
// fuse_last: the last sweep's sums go to acc_target.
static bool coop_usable(const Ising* m)
{
    const RingGeom& g = m->st.g;
    return m->coop_grid > 0 && g.nranks == 1 && m->n_multi == 1 && !m->st.p2p && !m->use_tma && !m->timing && !(m->tune & 2048) &&
           g.Lloc <= (int64_t)m->coop_grid * 256 * 2;
}

int coop_sweeps(Ising* m, int n, bool fuse_last, unsigned long long* series)
{
    const RingGeom& g = m->st.g;
    IsingCoopArgs s;
    for (int c = 0; c < 2; ++c) {
        RingPassArgs& a = s.a[c];
        a = RingPassArgs();
        a.own = m->st.vec[c];
        a.oth = m->st.vec[c ^ 1];
        a.nvec = g.Lloc;
        a.H = g.H;
        a.p0 = g.p0;
        for (int j = 0; j < 6; ++j) a.off[j] = g.off[c][j];
        a.seed = m->seed;
        a.colour = (uint32_t)c;
        a.draw = m->draw;
        a.ticket = nullptr;
        a.chunk = 32;
        a.acc = m->acc_target;
        a.rstride = m->st.rstride;
        a.Lfold = g.L; a.Nc = g.Nc;
        a.mask_from = g.ptail < g.L ? (int)(g.ptail - a.p0 > 0 ? g.ptail - a.p0 : 0) : 0x7FFFFFFF;
        a.nopush = 2;
        // narrow halo (2D: H = nx/2 + 1 vectors of L): no refresh between the passes, the few threads next to the ends of the
        // fold rebuild their wrapped neighbours; wide halo (3D: a plane): refresh after every pass
        a.wrap_mode = (8 * g.H > g.L || (m->tune & 4096)) ? 0 : ((g.H <= g.L && g.ptail >= g.H) ? 1 : 2);
    }
    s.L = g.L; s.H = g.H; s.Nc = g.Nc; s.ptail = g.ptail;
    s.halo_fast = (g.H <= g.L && g.ptail >= g.H) ? 1 : 0;
    s.n_sweeps = n; s.fuse_last = fuse_last ? 1 : 0; s.series = series;
    m->obs_valid = false;
    m->fused_pending = false;
    s.host_out = (fuse_last && !series && m->acc_target == m->d_acc) ? m->h_acc : nullptr;
    m->h_acc_pending = false;
    const int64_t need = (g.Lloc + 255) / 256;
    const int grid = (int)(need < (int64_t)m->coop_grid ? need : (int64_t)m->coop_grid);
    void* args[2] = {&s, &m->tab};
    const void* fn = m->ndim == 3
        ? (m->method == METHOD_METROPOLIS ? (const void*)ising_coop_kernel<6, METHOD_METROPOLIS> : (const void*)ising_coop_kernel<6, METHOD_HEATBATH>)
        : (m->method == METHOD_METROPOLIS ? (const void*)ising_coop_kernel<4, METHOD_METROPOLIS> : (const void*)ising_coop_kernel<4, METHOD_HEATBATH>);
    COUNT_LAUNCH();
    CK(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(256), args, 0, m->stream));
    m->fused_pending = fuse_last && !series;
    m->h_acc_pending = s.host_out != nullptr;
    m->draw += (uint64_t)n;
    return B200MC_OK;
}

// One MCS.  allow_fuse: this is the last sweep before control returns to the caller.
int sweep(Ising* m, bool allow_fuse = true, bool force_fuse = false)
{
    int rc;
    if (m->fused_pending) m->want_fused = false;  // the sums of the previous sweep were never asked for
    const bool fuse = force_fuse || (allow_fuse && m->want_fused && m->fuse_ok && !(m->tune & 8));
    if (coop_usable(m)) return coop_sweeps(m, 1, fuse, nullptr);
    for (int colour = 0; colour < 2; ++colour) {
        rc = m->ndim == 3 ? launch_pass<6>(m, colour, fuse && colour == 1, fuse) : launch_pass<4>(m, colour, fuse && colour == 1, fuse);
        if (rc) return rc;
    }
    m->fused_pending = fuse;
    m->draw += 1;
    return B200MC_OK;
}

// n MCS (update_n): small lattices in one cooperative launch, else sweep by sweep
int sweeps_n(Ising* m, int32_t n)
{
    if (n <= 0) return B200MC_OK;
    if (coop_usable(m)) {
        if (m->fused_pending) m->want_fused = false;
        const bool fuse = m->want_fused && m->fuse_ok && !(m->tune & 8);
        return coop_sweeps(m, n, fuse, nullptr);
    }
    for (int i = 0; i < n; ++i) { int rc = sweep(m, i == n - 1); if (rc) return rc; }
    return B200MC_OK;
}

template <int NNB>
int launch_pass_randoms(Ising* m, int colour)
{
    const RingGeom& g = m->st.g;
    RingPassArgs a = RingPassArgs();
    a.own = m->st.vec[colour];
    a.oth = m->st.vec[colour ^ 1];
    a.nvec = g.Lloc;
    a.H = g.H;
    a.p0 = g.p0;
    for (int j = 0; j < 6; ++j) a.off[j] = g.off[colour][j];
    a.seed = m->seed;
    a.colour = (uint32_t)colour;
    a.draw = m->draw;
    a.ticket = nullptr;
    a.chunk = 128;
    const unsigned grid = (unsigned)((g.Lloc + 255) / 256);
    m->obs_valid = false; m->fused_pending = false;
    COUNT_LAUNCH();
    if (m->method == METHOD_METROPOLIS)
        ising_pass_randoms_kernel<NNB, METHOD_METROPOLIS><<<grid, 256, 0, m->stream>>>(a, m->tabf, m->d_randoms, g.L, g.Nc);
    else
        ising_pass_randoms_kernel<NNB, METHOD_HEATBATH><<<grid, 256, 0, m->stream>>>(a, m->tabf, m->d_randoms, g.L, g.Nc);
    CK(cudaGetLastError());
    return ring_halo(&m->st, colour, m->stream);
}

int launch_measure(Ising* m, unsigned long long* acc);
int set_random(Ising* m);

int measure(Ising* m, int64_t* e, int64_t* mag)
{
    const RingGeom& g = m->st.g;
    if (m->obs_valid) {  // update -> calc_magne_sum -> calc_energy_sum (the drivers' loop) costs one pass
        if (e) *e = m->obs_e;
        if (mag) *mag = m->obs_m;
        return B200MC_OK;
    }
    const bool direct = m->fused_pending && m->h_acc_pending && g.nranks == 1 && m->n_multi == 1;  // already stored in h_acc by the kernel
    m->h_acc_pending = false;
    if (!m->fused_pending) {
        { int rcq = ring_p2p_quiesce(&m->st, m->stream); if (rcq) return rcq; }
        CK(cudaMemsetAsync(m->d_acc, 0, 2 * sizeof(unsigned long long) * m->n_multi, m->stream));
        { int rcm = launch_measure(m, m->d_acc); if (rcm) return rcm; }
    }
    m->fused_pending = false;  // the all-reduce below turns d_acc into global sums; they are cached in obs_*
    m->want_fused = true;
    bool exchanged = false;
    if (g.nranks > 1) {  // every rank gets the global sums (allreduce of {X, sum s})
        if (m->st.sums_p2p && !(m->tune & 8192)) {
            // direct transport: partial sums stored into every rank's mailbox over NVLink and added up by a one-warp
            // kernel that also writes the totals to pinned host memory -- no collective, no copy (B200MC_TUNE bit 13: NCCL)
            int rc = ring_sum_exchange(&m->st, m->d_acc, 2, m->h_acc, m->stream);
            if (rc) return rc;
            exchanged = true;
        } else {
            int rc = dist_allreduce_u64(m->st.comm, m->d_acc, 2, m->stream);
            if (rc) return rc;
        }
    }
    unsigned long long* acc = m->h_acc;
    if (!direct && !exchanged) CK(cudaMemcpyAsync(acc, m->d_acc, 2 * sizeof(unsigned long long) * m->n_multi, cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    m->obs_ev.resize(m->n_multi); m->obs_mv.resize(m->n_multi);
    for (int j = 0; j < m->n_multi; ++j) {
        const int64_t X = (int64_t)acc[2 * j], sum = (int64_t)acc[2 * j + 1];
        // E = -(bonds) + 2 X with bonds = (nnb/2) N;   M = 2 sum(s) - N
        m->obs_ev[j] = -(int64_t)(g.nnb / 2) * g.N + 2 * X;
        m->obs_mv[j] = 2 * sum - g.N;
    }
    m->obs_e = m->obs_ev[0];
    m->obs_m = m->obs_mv[0];
    m->obs_valid = true;
    if (e) *e = m->obs_e;
    if (mag) *mag = m->obs_m;
    return B200MC_OK;
}

int destroy(struct Ising* m);

int launch_measure(Ising* m, unsigned long long* acc)
{
    const RingGeom& g = m->st.g;
    COUNT_LAUNCH();
    if (m->ndim == 3)
        ising_measure_kernel<6><<<dim3(m->grid, m->n_multi), 256, 0, m->stream>>>(m->st.vec[0], m->st.vec[1], g.Lloc, g.H, g.p0, m->d_off1, g.L, g.Nc, g.ptail, acc, m->st.rstride);
    else
        ising_measure_kernel<4><<<dim3(m->grid, m->n_multi), 256, 0, m->stream>>>(m->st.vec[0], m->st.vec[1], g.Lloc, g.H, g.p0, m->d_off1, g.L, g.Nc, g.ptail, acc, m->st.rstride);
    CK(cudaGetLastError());
    return B200MC_OK;
}

// The drivers' inner loop on the device: mcs x [update; calc_magne_sum; calc_energy_sum]
// (app/ising3d_gpu_relaxation.f90:40-46) without a host round trip per MCS.  Every sweep adds its sums to its
// own slot of a device series (fused into the second colour pass where the layout allows, else by the measure
// kernel); one all-reduce (slab mode), one copy and one synchronisation at the end.
// fills d_series[mcs][n_multi][2] with the per-MCS sums {X, sum s} (all-reduced over the ranks in slab mode); no host copy
int relaxation_series_device(Ising* m, int32_t mcs)
{
    const RingGeom& g = m->st.g;
    const size_t per = 2 * (size_t)m->n_multi;   // sums per MCS: [sample][X, sum s]
    if (m->series_cap < (int64_t)mcs) {
        cudaFree(m->d_series);
        m->d_series = nullptr; m->series_cap = 0;
        CK(cudaMalloc(&m->d_series, (size_t)mcs * per * sizeof(unsigned long long)));
        m->series_cap = mcs;
    }
    CK(cudaMemsetAsync(m->d_series, 0, (size_t)mcs * per * sizeof(unsigned long long), m->stream));
    const bool fuse = m->fuse_ok && !(m->tune & 8);
    int rc = B200MC_OK;
    const bool coop = fuse && coop_usable(m);
    if (coop) {  // small lattice: the whole relaxation is one cooperative launch
        if (m->fused_pending) m->want_fused = false;
        rc = coop_sweeps(m, mcs, false, m->d_series);
    }
    for (int32_t i = 0; i < mcs && !rc && !coop; ++i) {
        m->acc_target = m->d_series + per * (size_t)i;
        m->fused_pending = false;
        rc = sweep(m, true, fuse);
        if (!rc && !fuse) {
            rc = ring_p2p_quiesce(&m->st, m->stream);
            if (!rc) rc = launch_measure(m, m->acc_target);
        }
    }
    m->acc_target = m->d_acc;
    m->fused_pending = false;
    m->want_fused = false;
    m->obs_valid = false;
    if (rc) return rc;
    if (g.nranks > 1 && (rc = dist_allreduce_u64(m->st.comm, m->d_series, (int)(per * (size_t)mcs), m->stream))) return rc;
    return B200MC_OK;
}

int run_relaxation(Ising* m, int32_t mcs, int64_t* e, int64_t* mag)
{
    const RingGeom& g = m->st.g;
    if (mcs < 0) ARG_FAIL("mcs < 0");
    if (mcs == 0) return B200MC_OK;
    const size_t per = 2 * (size_t)m->n_multi;
    int rc = relaxation_series_device(m, mcs);
    if (rc) return rc;
    std::vector<unsigned long long> host((size_t)mcs * per);
    CK(cudaMemcpyAsync(host.data(), m->d_series, host.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    m->obs_ev.resize(m->n_multi); m->obs_mv.resize(m->n_multi);
    for (int32_t i = 0; i < mcs; ++i)
        for (int j = 0; j < m->n_multi; ++j) {
            const int64_t X = (int64_t)host[per * (size_t)i + 2 * j], sum = (int64_t)host[per * (size_t)i + 2 * j + 1];
            const int64_t ei = -(int64_t)(g.nnb / 2) * g.N + 2 * X, mi = 2 * sum - g.N;
            if (e) e[(size_t)j * mcs + i] = ei;       // sample-major: sample j, MCS i
            if (mag) mag[(size_t)j * mcs + i] = mi;
            if (i == mcs - 1) { m->obs_ev[j] = ei; m->obs_mv[j] = mi; }
        }
    m->obs_e = m->obs_ev[0]; m->obs_m = m->obs_mv[0]; m->obs_valid = true;
    return B200MC_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// The drivers' whole measurement on the device (app/ising3d_gpu_relaxation.f90:37-55,
// app/ising2d_gpu_relaxation.f90:33-52):
//     do sample = 1, tot_sample:  set_allup_spin | set_random_spin
//        do i = 1, mcs:  update; m = calc_magne_sum; e = calc_energy_sum; order_parameter(i)%add_data(m / N, e / N)
// `variance_covariance_kahan` is the drivers' external accumulator (fpm.toml:14, osada-yum/Numerical_utilities, not
// vendored, no version pinned).  Restated here from its use (add_data / num_sample / mean1,2 / square_mean1,2 / var1,2 /
// cov, app/ising3d_gpu_relaxation.f90:46-55): Kahan-compensated running sums of v1, v2, v1^2, v2^2, v1 v2;
// mean = sum / n, square_mean = sum of squares / n, var = n / (n - 1) (square_mean - mean^2) (unbiased; 0 for n = 1),
// cov = n / (n - 1) (mean_v1v2 - mean1 mean2).  One thread per MCS adds the samples of a batch in sample order, so the
// result does not depend on the batch size; explicit __dadd_rn / __dmul_rn keep the compiler from contracting the
// compensation away (the CPU oracle is compiled with -ffp-contract=off) -- the two agree bit for bit.
// ---------------------------------------------------------------------------------------------------------------
__global__ void ising_stats_kernel(const unsigned long long* series, int mcs, int n_multi, long long N, int nnb, double n_inv, double* stats)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= mcs) return;
    double* st = stats + (size_t)i * 10;
    double s[5], c[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) { s[k] = st[k]; c[k] = st[5 + k]; }
    for (int j = 0; j < n_multi; ++j) {
        const long long X = (long long)series[((size_t)i * n_multi + j) * 2], sum = (long long)series[((size_t)i * n_multi + j) * 2 + 1];
        const long long e = -(long long)(nnb / 2) * N + 2 * X, mg = 2 * sum - N;
        const double v1 = __dmul_rn((double)mg, n_inv), v2 = __dmul_rn((double)e, n_inv);   // m * n_inv_r64, e * n_inv_r64 (:46)
        const double x[5] = {v1, v2, __dmul_rn(v1, v1), __dmul_rn(v2, v2), __dmul_rn(v1, v2)};
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const double y = __dadd_rn(x[k], -c[k]);
            const double t = __dadd_rn(s[k], y);
            c[k] = __dadd_rn(__dadd_rn(t, -s[k]), -y);
            s[k] = t;
        }
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) { st[k] = s[k]; st[5 + k] = c[k]; }
}

int run_relaxation_stats(Ising* m, int32_t mcs, int32_t tot_sample, int32_t random_start, double* out)
{
    const RingGeom& g = m->st.g;
    if (mcs <= 0 || tot_sample <= 0) ARG_FAIL("mcs and tot_sample must be > 0");
    if (!out) ARG_FAIL("null output");
    if (tot_sample % m->n_multi) ARG_FAIL("tot_sample (%d) must be a multiple of the batch size n_multi (%d)", tot_sample, m->n_multi);
    if (m->stats_cap < (int64_t)mcs) {
        cudaFree(m->d_stats);
        m->d_stats = nullptr; m->stats_cap = 0;
        CK(cudaMalloc(&m->d_stats, (size_t)mcs * 10 * sizeof(double)));
        m->stats_cap = mcs;
    }
    CK(cudaMemsetAsync(m->d_stats, 0, (size_t)mcs * 10 * sizeof(double), m->stream));
    const double n_inv = 1.0 / (double)g.N;   // n_inv_r64 = 1 / real(nx * ny * nz, real64), app/ising3d_gpu_relaxation.f90:11
    for (int32_t done = 0; done < tot_sample; done += m->n_multi) {
        int rc = random_start ? set_random(m) : ring_fill(&m->st, 1, m->stream);
        if (rc) return rc;
        m->obs_valid = false; m->fused_pending = false;
        if ((rc = relaxation_series_device(m, mcs))) return rc;
        COUNT_LAUNCH();
        ising_stats_kernel<<<(mcs + 127) / 128, 128, 0, m->stream>>>(m->d_series, mcs, m->n_multi, (long long)g.N, g.nnb, n_inv, m->d_stats);
        CK(cudaGetLastError());
    }
    std::vector<double> host((size_t)mcs * 10);
    CK(cudaMemcpyAsync(host.data(), m->d_stats, host.size() * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    const double n = (double)tot_sample;
    for (int32_t i = 0; i < mcs; ++i) {
        const double* s = &host[(size_t)i * 10];
        double* o = out + (size_t)i * 8;
        const double mean1 = s[0] / n, mean2 = s[1] / n, sq1 = s[2] / n, sq2 = s[3] / n, m12 = s[4] / n;
        const double f = tot_sample > 1 ? n / (n - 1.0) : 0.0;
        o[0] = n; o[1] = mean1; o[2] = mean2; o[3] = sq1; o[4] = sq2;
        o[5] = f * (sq1 - mean1 * mean1); o[6] = f * (sq2 - mean2 * mean2); o[7] = f * (m12 - mean1 * mean2);
    }
    return B200MC_OK;
}

int create(void** out, int ndim, int64_t nx, int64_t ny, int64_t nz, double kbt, int32_t iseed,
           int rank = 0, int nranks = 1, const char* nccl_id = nullptr, int n_multi = 1)
{
    if (n_multi < 1) ARG_FAIL("n_multi must be >= 1");
    if (n_multi > 1 && nranks > 1) ARG_FAIL("a batch of samples and a slab decomposition cannot be combined");
    if (n_multi > 65535) ARG_FAIL("n_multi too large");
    if (!out) ARG_FAIL("null handle pointer");
    *out = nullptr;
    if (!(kbt > 0.0)) ARG_FAIL("kbt must be > 0");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        snprintf(g_b200mc_err, sizeof(g_b200mc_err), "no CUDA device: this library has no CPU fallback");
        return B200MC_ERR_CUDA;
    }
    Ising* m = new (std::nothrow) Ising();
    if (!m) ARG_FAIL("out of host memory");
    m->ndim = ndim; m->nx = nx; m->ny = ny; m->nz = ndim == 3 ? nz : 0;
    m->stream = 0; m->d_acc = nullptr; m->h_acc = nullptr; m->h_acc_pending = false; m->acc_target = nullptr; m->d_series = nullptr; m->series_cap = 0; m->d_stats = nullptr; m->stats_cap = 0; m->d_off1 = nullptr; m->d_randoms = nullptr; m->d_ticket = nullptr;
    { const char* t = getenv("B200MC_TUNE"); m->tune = t ? atoi(t) : 0; t = getenv("B200MC_CHUNK"); m->chunk = t ? atoi(t) : 128; m->chunk = TK_CHUNK;  /* compile-time now */
    }
    m->method = METHOD_METROPOLIS; m->seed = (uint32_t)iseed; m->draw = 0; m->alive = true;
    m->obs_valid = false; m->timing = false; m->ev_used = 0;
    m->fuse_ok = false; m->want_fused = false; m->fused_pending = false;
    m->comm_stream = nullptr; m->ev_boundary = m->ev_halo = nullptr; m->st.comm = nullptr;
    int rc = ring_geom_init(&m->st.g, nx, ny, m->nz);
    if (rc) { delete m; return rc; }
    rc = ring_geom_set_slab(&m->st.g, rank, nranks);
    if (rc) { delete m; return rc; }
    if (nranks == 1 && n_multi == 1 && (m->tune & 16)) {
        if (cudaStreamCreateWithFlags(&m->comm_stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&m->ev_boundary, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&m->ev_halo, cudaEventDisableTiming) != cudaSuccess) {
            snprintf(g_b200mc_err, sizeof(g_b200mc_err), "cannot create the second stream / events");
            delete m; return B200MC_ERR_CUDA;
        }
    }
    if (nranks > 1) {
        if (!nccl_id) { delete m; ARG_FAIL("slab mode needs the NCCL unique id of the job (b200mc_dist_unique_id on rank 0, broadcast by the caller)"); }
        rc = dist_comm_init(&m->st.comm, rank, nranks, nccl_id);
        if (rc) { delete m; return rc; }
        if (cudaStreamCreateWithFlags(&m->comm_stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&m->ev_boundary, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&m->ev_halo, cudaEventDisableTiming) != cudaSuccess) {
            snprintf(g_b200mc_err, sizeof(g_b200mc_err), "cannot create the comm stream / events");
            dist_comm_destroy(m->st.comm); delete m; return B200MC_ERR_CUDA;
        }
    }
    m->n_multi = n_multi;
    m->st.n_rep = n_multi;
    rc = ring_alloc(&m->st);
    if (rc) { destroy(m); return rc; }
    if (cudaMalloc(&m->d_ticket, (2 * TK_NCNT * 64 + 64) * sizeof(unsigned int)) != cudaSuccess ||   // two counter sets + the finished-warps counter
        cudaMemset(m->d_ticket, 0, (2 * TK_NCNT * 64 + 64) * sizeof(unsigned int)) != cudaSuccess ||
        cudaMalloc(&m->d_acc, 2 * sizeof(unsigned long long) * n_multi) != cudaSuccess ||
        cudaHostAlloc(&m->h_acc, 2 * sizeof(unsigned long long) * n_multi, cudaHostAllocDefault) != cudaSuccess ||
        cudaMalloc(&m->d_off1, 6 * sizeof(int64_t)) != cudaSuccess) {
        destroy(m);
        snprintf(g_b200mc_err, sizeof(g_b200mc_err), "cudaMalloc failed");
        return B200MC_ERR_CUDA;
    }
    m->acc_target = m->d_acc;
    cudaMemcpy(m->d_off1, m->st.g.off[1], 6 * sizeof(int64_t), cudaMemcpyHostToDevice);
    // persistent-style grid: SMs x resident blocks, grid-stride over the vectors
    int dev = 0, sms = 148, occ = 4;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (ndim == 3) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ising_pass_kernel<6, METHOD_METROPOLIS, true>, 256, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ising_pass_kernel<4, METHOD_METROPOLIS, true>, 256, 0);
    if (occ < 1) occ = 1;
    int64_t need = (m->st.g.Lloc + 255) / 256;
    m->grid = (int)(need < (int64_t)sms * occ ? need : (int64_t)sms * occ);
    int occp = occ;
    if (ndim == 3) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occp, ising_pass_kernel<6, METHOD_METROPOLIS, true, true>, 256, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occp, ising_pass_kernel<4, METHOD_METROPOLIS, true, true>, 256, 0);
    if (occp < 1) occp = 1;
    m->grid_push = (int)(need < (int64_t)sms * occp ? need : (int64_t)sms * occp);
    {   // cooperative small-lattice sweep kernel: all blocks must be resident
        int coop_ok = 0, occc = 0;
        cudaDeviceGetAttribute(&coop_ok, cudaDevAttrCooperativeLaunch, dev);
        if (ndim == 3) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occc, ising_coop_kernel<6, METHOD_METROPOLIS>, 256, 0);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occc, ising_coop_kernel<4, METHOD_METROPOLIS>, 256, 0);
        m->coop_grid = (coop_ok && occc >= 1) ? sms * (occc > 2 ? 2 : occc) : 0;
        cudaGetLastError();
    }
    { const char* t = getenv("B200MC_SLAB_NB"); m->slab_nb = t ? atoi(t) : 0; }
    { const char* t = getenv("B200MC_SELF_CLEAN"); m->self_clean = !(t && atoi(t) == 0); m->acc_zeroed = false; }
    m->use_tma = false; m->tma_grid = 0;
    if (nranks == 1 && n_multi == 1 && (m->tune & 128)) {
        // opt in to the maximum dynamic shared memory of the staged kernels and size their grid
        const int smem = TMA_STAGES * m->st.g.nnb * TMA_SLOT;
        cudaError_t e;
        int tocc = 0;
        if (ndim == 3) {
            e = cudaFuncSetAttribute(ising_pass_tma_kernel<6, METHOD_METROPOLIS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(ising_pass_tma_kernel<6, METHOD_METROPOLIS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(ising_pass_tma_kernel<6, METHOD_HEATBATH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(ising_pass_tma_kernel<6, METHOD_HEATBATH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e == cudaSuccess) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&tocc, ising_pass_tma_kernel<6, METHOD_METROPOLIS, false>, 288, smem);
        } else {
            e = cudaFuncSetAttribute(ising_pass_tma_kernel<4, METHOD_METROPOLIS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(ising_pass_tma_kernel<4, METHOD_METROPOLIS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(ising_pass_tma_kernel<4, METHOD_HEATBATH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(ising_pass_tma_kernel<4, METHOD_HEATBATH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e == cudaSuccess) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&tocc, ising_pass_tma_kernel<4, METHOD_METROPOLIS, false>, 288, smem);
        }
        if (e == cudaSuccess && tocc >= 1) { m->use_tma = true; m->tma_grid = sms * tocc; }
        else cudaGetLastError();
    }
    m->beta = 1 / kbt;
    m->fuse_ok = m->st.g.off[1][0] == 0;
    build_tables(m);
    rc = ring_fill(&m->st, 1, m->stream);  // set_allup_spin
    if (rc) { destroy(m); return rc; }
    if (nranks == 1 && n_multi == 1 && (m->tune & 16) && m->st.g.Nc % 16 == 0 && m->st.g.Lloc >= 4 * m->st.g.H) {
        // experiment (B200MC_TUNE bit 4): one GPU runs the fused update + halo kernel against its own arrays
        rc = ring_p2p_connect_self(&m->st);
        if (rc) { destroy(m); return rc; }
    }
    *out = m;
    return B200MC_OK;
}

int destroy(Ising* m)
{
    if (!m) return B200MC_OK;
    ring_p2p_quiesce(&m->st, m->stream);  // the neighbours' last pushes have landed before the arrays go away
    cudaStreamSynchronize(m->stream);
    ring_free(&m->st);
    cudaFree(m->d_acc);
    cudaFreeHost(m->h_acc);
    cudaFree(m->d_series);
    cudaFree(m->d_stats);
    cudaFree(m->d_off1);
    cudaFree(m->d_randoms);
    cudaFree(m->d_ticket);
    for (cudaEvent_t e : m->evs) cudaEventDestroy(e);
    if (m->comm_stream) { cudaStreamSynchronize(m->comm_stream); cudaStreamDestroy(m->comm_stream); }
    if (m->ev_boundary) cudaEventDestroy(m->ev_boundary);
    if (m->ev_halo) cudaEventDestroy(m->ev_halo);
    dist_comm_destroy(m->st.comm);
    delete m;
    return B200MC_OK;
}

int set_random(Ising* m)
{
    const RingGeom& g = m->st.g;
    m->obs_valid = false; m->fused_pending = false;
    { int rcq = ring_p2p_quiesce(&m->st, m->stream); if (rcq) return rcq; }
    for (int c = 0; c < 2; ++c) {
        COUNT_LAUNCH();
        ring_random_bits_kernel<<<dim3((unsigned)((g.Lloc + 255) / 256), (unsigned)m->n_multi), 256, 0, m->stream>>>(m->st.vec[c], g.Lloc, g.H, g.p0, m->seed, m->draw, (uint32_t)c, m->st.rstride);
        CK(cudaGetLastError());
    }
    m->draw += 1;
    int rc = ring_halo(&m->st, 0, m->stream);
    if (rc) return rc;
    return ring_halo(&m->st, 1, m->stream);
}

int update_with_randoms(Ising* m, const double* randoms)
{
    if (!randoms) ARG_FAIL("null randoms");
    if (m->n_multi > 1) { snprintf(g_b200mc_err, sizeof(g_b200mc_err), "update_with_randoms: not available for a batch of samples"); return B200MC_ERR_UNSUPPORTED; }
    const RingGeom& g = m->st.g;
    { int rcq = ring_p2p_quiesce(&m->st, m->stream); if (rcq) return rcq; }
    if (!m->d_randoms) CK(cudaMalloc(&m->d_randoms, (size_t)g.N * sizeof(double)));
    CK(cudaMemcpyAsync(m->d_randoms, randoms, (size_t)g.N * sizeof(double), cudaMemcpyHostToDevice, m->stream));
    for (int colour = 0; colour < 2; ++colour) {
        int rc = m->ndim == 3 ? launch_pass_randoms<6>(m, colour) : launch_pass_randoms<4>(m, colour);
        if (rc) return rc;
    }
    CK(cudaStreamSynchronize(m->stream));  // the host array may be reused by the caller
    return B200MC_OK;
}

int get_timing(Ising* m, int64_t* launches, double* total_ms)
{
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

int skip(Ising* m, int64_t n_skip)
{
    if (n_skip < 0) ARG_FAIL("n_skip < 0");
    // the reference offsets its XORWOW stream by n_skip uniforms (src/ising3d_gpu_m.f90:72-77);
    // one generate call draws nall of them, so advance the draw counter by ceil(n_skip / nall)
    m->draw += (uint64_t)((n_skip + m->st.g.N - 1) / m->st.g.N);
    return B200MC_OK;
}

}  // namespace

#define H(h) (reinterpret_cast<Ising*>(h))
#define CHECK_H(h, nd)                                                              \
    do {                                                                            \
        if (!(h) || H(h)->ndim != (nd) || !H(h)->alive) ARG_FAIL("invalid handle"); \
    } while (0)

extern "C" {

const char* b200mc_last_error(void) { return g_b200mc_err; }
void b200mc_print_last_error(void) { fprintf(stderr, "b200mc: %s\n", g_b200mc_err); }
int b200mc_version(void) { return 100; }
// one row of the drivers' output table (app/ising3d_gpu_relaxation.f90:49-55): nall, num_sample, i, mean1, mean2,
// square_mean1, square_mean2, nall * var1, nall * var2, nall * cov -- blank-separated like the list-directed '(*(g0, 1x))'
int b200mc_format_relaxation_row(int64_t nall, int32_t i, const double row[8], char* buf, int32_t buflen)
{
    if (!row || !buf || buflen <= 0) ARG_FAIL("null argument");
    const int n = snprintf(buf, (size_t)buflen, "%lld %lld %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g", (long long)nall, (long long)row[0], (int)i,
                           row[1], row[2], row[3], row[4], (double)nall * row[5], (double)nall * row[6], (double)nall * row[7]);
    if (n < 0 || n >= buflen) ARG_FAIL("buffer too small (%d bytes needed)", n + 1);
    return B200MC_OK;
}
unsigned long long b200mc_launch_count(void) { return g_b200mc_launches; }

__global__ void philox_debug_kernel(uint4 c, uint2 k, uint4* out) { *out = philox4x32_10(c, k); }
int b200mc_debug_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint4* d = nullptr;
    CK(cudaMalloc(&d, sizeof(uint4)));
    philox_debug_kernel<<<1, 1>>>(make_uint4(ctr[0], ctr[1], ctr[2], ctr[3]), make_uint2(key[0], key[1]), d);
    uint4 r;
    cudaError_t e = cudaMemcpy(&r, d, sizeof(r), cudaMemcpyDeviceToHost);
    cudaFree(d);
    CK(e);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
    return B200MC_OK;
}

#define DEFINE_COMMON(PFX, ND)                                                                    \
    int PFX##_destroy(void* h) { if (!h) return B200MC_OK; CHECK_H(h, ND); return destroy(H(h)); } \
    int PFX##_set_stream(void* h, void* s) { CHECK_H(h, ND); H(h)->stream = (cudaStream_t)s; return B200MC_OK; } \
    int PFX##_skip_curand(void* h, int64_t n) { CHECK_H(h, ND); return skip(H(h), n); }           \
    int PFX##_set_allup_spin(void* h) { CHECK_H(h, ND); H(h)->obs_valid = false; H(h)->fused_pending = false; return ring_fill(&H(h)->st, 1, H(h)->stream); } \
    int PFX##_set_random_spin(void* h) { CHECK_H(h, ND); return set_random(H(h)); }               \
    int PFX##_set_beta(void* h, double beta) { CHECK_H(h, ND); if (!(beta >= 0.0)) ARG_FAIL("beta must be >= 0"); H(h)->beta = beta; return build_tables(H(h)); } \
    int PFX##_set_kbt(void* h, double kbt) { CHECK_H(h, ND); if (!(kbt > 0.0)) ARG_FAIL("kbt must be > 0"); H(h)->beta = 1 / kbt; return build_tables(H(h)); } \
    int PFX##_set_method(void* h, int32_t method) { CHECK_H(h, ND); if (method != METHOD_METROPOLIS && method != METHOD_HEATBATH) ARG_FAIL("unknown method %d", method); H(h)->method = method; return build_tables(H(h)); } \
    int PFX##_update(void* h) { CHECK_H(h, ND); return sweep(H(h)); }                             \
    int PFX##_update_n(void* h, int32_t n) { CHECK_H(h, ND); return sweeps_n(H(h), n); } \
    int PFX##_update_with_randoms(void* h, const double* r) { CHECK_H(h, ND); return update_with_randoms(H(h), r); } \
    int PFX##_calc_energy_sum(void* h, int64_t* e) { CHECK_H(h, ND); return measure(H(h), e, nullptr); } \
    int PFX##_calc_magne_sum(void* h, int64_t* m) { CHECK_H(h, ND); return measure(H(h), nullptr, m); } \
    int PFX##_measure(void* h, int64_t* e, int64_t* m) { CHECK_H(h, ND); return measure(H(h), e, m); } \
    int PFX##_run_relaxation(void* h, int32_t mcs, int64_t* e, int64_t* m) { CHECK_H(h, ND); return run_relaxation(H(h), mcs, e, m); } \
    int PFX##_run_relaxation_stats(void* h, int32_t mcs, int32_t tot_sample, int32_t random_start, double* out) { CHECK_H(h, ND); return run_relaxation_stats(H(h), mcs, tot_sample, random_start, out); } \
    int32_t PFX##_n_multi(void* h) { return h ? H(h)->n_multi : -1; }                              \
    int PFX##_measure_multi(void* h, int64_t* e, int64_t* m) { CHECK_H(h, ND); int rc = measure(H(h), nullptr, nullptr); if (rc) return rc; \
        for (int j = 0; j < H(h)->n_multi; ++j) { if (e) e[j] = H(h)->obs_ev[j]; if (m) m[j] = H(h)->obs_mv[j]; } return B200MC_OK; } \
    int PFX##_get_spins_multi(void* h, int32_t sample, int32_t* out) { CHECK_H(h, ND); if (!out) ARG_FAIL("null output"); return ring_export_i32(&H(h)->st, out, ND == 2 ? RING_MAP_PM1 : RING_MAP_IDENTITY, H(h)->stream, sample); } \
    int PFX##_set_spins_multi(void* h, int32_t sample, const int32_t* in) { CHECK_H(h, ND); if (!in) ARG_FAIL("null input"); H(h)->obs_valid = false; H(h)->fused_pending = false; return ring_import_i32(&H(h)->st, in, ND == 2 ? RING_MAP_PM1 : RING_MAP_IDENTITY, H(h)->stream, sample); } \
    int PFX##_get_spins(void* h, int32_t* out) { CHECK_H(h, ND); if (!out) ARG_FAIL("null output"); return ring_export_i32(&H(h)->st, out, ND == 2 ? RING_MAP_PM1 : RING_MAP_IDENTITY, H(h)->stream); } \
    int PFX##_set_spins(void* h, const int32_t* in) { CHECK_H(h, ND); if (!in) ARG_FAIL("null input"); H(h)->obs_valid = false; H(h)->fused_pending = false; return ring_import_i32(&H(h)->st, in, ND == 2 ? RING_MAP_PM1 : RING_MAP_IDENTITY, H(h)->stream); } \
    int64_t PFX##_nx(void* h) { return h ? H(h)->nx : -1; }                                       \
    int64_t PFX##_ny(void* h) { return h ? H(h)->ny : -1; }                                       \
    int64_t PFX##_nall(void* h) { return h ? H(h)->st.g.N : -1; }                                 \
    double PFX##_kbt(void* h) { return h ? 1 / H(h)->beta : 0.0; }                                \
    double PFX##_beta(void* h) { return h ? H(h)->beta : 0.0; }                                   \
    int PFX##_set_timing(void* h, int32_t on) { CHECK_H(h, ND); H(h)->timing = on != 0; H(h)->ev_used = 0; return B200MC_OK; } \
    int PFX##_get_timing(void* h, int64_t* launches, double* total_ms) { CHECK_H(h, ND); return get_timing(H(h), launches, total_ms); } \
    int PFX##_sync(void* h) { CHECK_H(h, ND); CK(cudaStreamSynchronize(H(h)->stream)); return B200MC_OK; }

DEFINE_COMMON(b200mc_ising3d, 3)
DEFINE_COMMON(b200mc_ising2d, 2)

int b200mc_ising3d_create(void** h, int64_t nx, int64_t ny, int64_t nz, double kbt, int32_t iseed)
{
    if (nz <= 0) ARG_FAIL("nz must be > 0");
    return create(h, 3, nx, ny, nz, kbt, iseed);
}
int b200mc_ising2d_create(void** h, int64_t nx, int64_t ny, double kbt, int32_t iseed)
{
    return create(h, 2, nx, ny, 0, kbt, iseed);
}
// batch of independent samples ("multi-sample batch": the drivers' tot_sample loop, n_multi samples at a time)
int b200mc_ising3d_create_multi(void** h, int64_t nx, int64_t ny, int64_t nz, double kbt, int32_t iseed, int32_t n_multi)
{
    if (nz <= 0) ARG_FAIL("nz must be > 0");
    return create(h, 3, nx, ny, nz, kbt, iseed, 0, 1, nullptr, n_multi);
}
int b200mc_ising2d_create_multi(void** h, int64_t nx, int64_t ny, double kbt, int32_t iseed, int32_t n_multi)
{
    return create(h, 2, nx, ny, 0, kbt, iseed, 0, 1, nullptr, n_multi);
}
int b200mc_dist_unique_id(char out[128]) { return dist_unique_id(out); }
int b200mc_ising3d_create_slab(void** h, int64_t nx, int64_t ny, int64_t nz, double kbt, int32_t iseed,
                               int32_t rank, int32_t nranks, const char nccl_id[128])
{
    if (nz <= 0) ARG_FAIL("nz must be > 0");
    return create(h, 3, nx, ny, nz, kbt, iseed, rank, nranks, nccl_id);
}
int b200mc_ising2d_create_slab(void** h, int64_t nx, int64_t ny, double kbt, int32_t iseed,
                               int32_t rank, int32_t nranks, const char nccl_id[128])
{
    return create(h, 2, nx, ny, 0, kbt, iseed, rank, nranks, nccl_id);
}
int b200mc_ring_slab_geometry(int64_t nx, int64_t ny, int64_t nz, int32_t rank, int32_t nranks, int64_t out[6])
{
    RingGeom g;
    int rc = ring_geom_init(&g, nx, ny, nz);
    if (rc) return rc;
    rc = ring_geom_set_slab(&g, rank, nranks);
    if (rc) return rc;
    out[0] = g.Nc; out[1] = g.L; out[2] = g.H; out[3] = g.p0; out[4] = g.Lloc; out[5] = g.ptail;
    return B200MC_OK;
}
int b200mc_ising3d_p2p_handles(void* h, char out[192]) { CHECK_H(h, 3); return ring_p2p_export(&H(h)->st, out); }
int b200mc_ising2d_p2p_handles(void* h, char out[192]) { CHECK_H(h, 2); return ring_p2p_export(&H(h)->st, out); }
int b200mc_ising3d_p2p_connect(void* h, const char prev[192], const char next[192]) { CHECK_H(h, 3); return ring_p2p_connect(&H(h)->st, prev, next); }
int b200mc_ising2d_p2p_connect(void* h, const char prev[192], const char next[192]) { CHECK_H(h, 2); return ring_p2p_connect(&H(h)->st, prev, next); }
int b200mc_ising3d_p2p_connect_sums(void* h, const char* handles) { CHECK_H(h, 3); return ring_p2p_connect_sums(&H(h)->st, handles); }
int b200mc_ising2d_p2p_connect_sums(void* h, const char* handles) { CHECK_H(h, 2); return ring_p2p_connect_sums(&H(h)->st, handles); }
int b200mc_ising3d_rank_info(void* h, int32_t* rank, int32_t* nranks) { CHECK_H(h, 3); *rank = H(h)->st.g.rank; *nranks = H(h)->st.g.nranks; return B200MC_OK; }
int b200mc_ising2d_rank_info(void* h, int32_t* rank, int32_t* nranks) { CHECK_H(h, 2); *rank = H(h)->st.g.rank; *nranks = H(h)->st.g.nranks; return B200MC_OK; }
int64_t b200mc_ising3d_nz(void* h) { return h ? H(h)->nz : -1; }
// debug: ns block 0 of the fused colour-pass kernels has spent waiting for the neighbours' flags so far
unsigned long long b200mc_debug_slab_wait_ns(void* h)
{
    if (!h || !H(h)->st.flags) return 0;
    unsigned long long v = 0;
    cudaStreamSynchronize(H(h)->stream);
    cudaMemcpy(&v, H(h)->st.flags + 48, sizeof(v), cudaMemcpyDeviceToHost);
    return v;
}
int b200mc_ising3d_get_ws(void* h, double out[14])
{
    CHECK_H(h, 3);
    for (int i = 0; i < 14; ++i) out[i] = H(h)->ws3[i];
    return B200MC_OK;
}
int b200mc_ising2d_get_exparr(void* h, double out[17])
{
    CHECK_H(h, 2);
    for (int i = 0; i < 17; ++i) out[i] = H(h)->exparr[i];
    return B200MC_OK;
}

}  // extern "C"
