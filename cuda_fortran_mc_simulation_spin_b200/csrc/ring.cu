// Folded-ring storage: geometry, halo ("norishiro") refresh, import/export.
// See ring.cuh for the layout.
#include "ring.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>
#include <limits.h>

thread_local char g_b200mc_err[512] = {0};
unsigned long long g_b200mc_launches = 0;

int ring_geom_init(RingGeom* g, int64_t nx, int64_t ny, int64_t nz)
{
    const bool is3d = nz > 0;
    if (nx < 3 || ny < 2 || (is3d && nz < 2)) ARG_FAIL("lattice too small: %lld x %lld x %lld", (long long)nx, (long long)ny, (long long)nz);
    // the reference colours by linear-index parity, which is a valid
    // checkerboard only for nx odd and (2D) ny even / (3D) ny odd, nz even.
    if ((nx & 1) == 0) ARG_FAIL("helical checkerboard needs odd nx (got %lld)", (long long)nx);
    if (!is3d && (ny & 1)) ARG_FAIL("helical 2D checkerboard needs even ny (got %lld)", (long long)ny);
    if (is3d && ((ny & 1) == 0 || (nz & 1))) ARG_FAIL("helical 3D checkerboard needs odd ny and even nz (got ny=%lld nz=%lld)", (long long)ny, (long long)nz);
    const int64_t nxy = nx * ny;
    g->N = is3d ? nxy * nz : nxy;
    g->Nc = g->N / 2;
    g->L = (g->Nc + 15) / 16;
    g->P = is3d ? nxy : nx;
    g->nnb = is3d ? 6 : 4;
    const int64_t h = (nx - 1) / 2, gg = (nxy - 1) / 2;
    g->H = (is3d ? gg : h) + 1;
    for (int c = 0; c < 2; ++c) {
        g->off[c][0] = -1 + c;      // i-1
        g->off[c][1] = c;           // i+1
        g->off[c][2] = h + c;       // i+nx
        g->off[c][3] = -h - 1 + c;  // i-nx
        g->off[c][4] = is3d ? gg + c : 0;       // i+nxy
        g->off[c][5] = is3d ? -gg - 1 + c : 0;  // i-nxy
    }
    // lanes 0..14 are full; lane 15 holds Nc - 15 L sites (may be <= 0 for tiny rings)
    int64_t l15 = g->Nc - 15 * g->L;
    g->ptail = l15 < 0 ? 0 : l15;
    g->p0 = 0; g->Lloc = g->L; g->rank = 0; g->nranks = 1;
    if (g->N / g->P < 2) ARG_FAIL("lattice too small");
    if (g->L + 2 * g->H >= (int64_t)0x7C000000) ARG_FAIL("lattice too large for 32-bit vector indices (%lld vectors per colour)", (long long)g->L);
    return B200MC_OK;
}

int ring_geom_set_slab(RingGeom* g, int rank, int nranks)
{
    if (nranks < 1 || rank < 0 || rank >= nranks) ARG_FAIL("bad rank %d / %d", rank, nranks);
    if (nranks == 1) return B200MC_OK;
    if (g->Nc % 16) ARG_FAIL("slab decomposition needs the number of sites per colour (%lld) to be a multiple of 16", (long long)g->Nc);
    const int64_t base = g->L / nranks, rem = g->L % nranks;
    g->Lloc = base + (rank < rem ? 1 : 0);
    g->p0 = rank * base + (rank < rem ? rank : rem);
    g->rank = rank; g->nranks = nranks;
    if (base < g->H) ARG_FAIL("slab too thin: %lld positions per rank < halo %lld", (long long)base, (long long)g->H);
    return B200MC_OK;
}

int ring_alloc(RingStore* s)
{
    if (s->n_rep < 1) s->n_rep = 1;
    s->rstride = s->g.Lloc + 2 * s->g.H;
    const size_t nv = (size_t)s->rstride * (size_t)s->n_rep;
    s->vec[0] = s->vec[1] = nullptr;
    s->stage = nullptr;
    s->p2p = false; s->flags = nullptr; s->n_peer_maps = 0; s->push_seq = 0; s->Lloc_prev = 0;
    s->n_sum_maps = 0; s->sums_p2p = false; s->sum_seq = 0;
    CK(cudaMalloc(&s->vec[0], nv * sizeof(uint4)));
    CK(cudaMalloc(&s->vec[1], nv * sizeof(uint4)));
    s->stage_elems = 1 << 24;
    if (s->stage_elems > s->g.N + 2 * s->g.P) s->stage_elems = s->g.N + 2 * s->g.P;
    CK(cudaMalloc(&s->stage, (size_t)s->stage_elems * sizeof(int32_t)));
    return B200MC_OK;
}

void ring_free(RingStore* s)
{
    ring_p2p_close(s);
    cudaFree(s->flags);
    s->flags = nullptr;
    cudaFree(s->vec[0]);
    cudaFree(s->vec[1]);
    cudaFree(s->stage);
    s->vec[0] = s->vec[1] = nullptr;
    s->stage = nullptr;
}

int ring_fill(RingStore* s, uint8_t value, cudaStream_t st)
{
    int rcq = ring_p2p_quiesce(s, st);
    if (rcq) return rcq;
    const size_t nv = (size_t)s->rstride * (size_t)s->n_rep;
    CK(cudaMemsetAsync(s->vec[0], value, nv * sizeof(uint4), st));
    CK(cudaMemsetAsync(s->vec[1], value, nv * sizeof(uint4), st));
    // Direct transport between ranks: the memset covers the halo vectors, which the neighbours store into from their next
    // colour pass on.  A rank that finishes its fill early must not push before every rank has finished filling (its
    // push would be overwritten by the slower neighbour's memset, whose flag already shows the new sequence number):
    // a one-word all-reduce on the stream is the barrier -- it completes on a rank only after every rank's stream has
    // reached it, i.e. after every memset.  (set_random_spin / set_spins refresh their halos through ncclSend/Recv,
    // which orders the ranks pairwise.)
    if (s->p2p && s->g.nranks > 1 && s->comm && s->flags) {
        unsigned long long* scratch = reinterpret_cast<unsigned long long*>(s->flags + 56);
        int rc = dist_allreduce_u64(s->comm, scratch, 1, st);
        if (rc) return rc;
    }
    return B200MC_OK;
}

// ---------------------------------------------------------------------------
// halo refresh
// ---------------------------------------------------------------------------
// (pos_mod, ring_site and the per-item halo functions live in ring.cuh: the cooperative small-lattice sweep kernel
// of ising_kernels.cuh refreshes the halo between its colour passes with the same code)

// generic, byte-granular: one thread per (dirty vector, lane).  Dirty vectors
// are the 2H halo vectors and the tail positions [ptail, L).
__global__ void ring_halo_generic_kernel(uint8_t* base, int64_t L, int64_t H, int64_t Nc,
                                         int64_t ptail, int64_t v_begin, int64_t n_items, int64_t rstride = 0)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_items) return;
    base += (size_t)blockIdx.y * (size_t)rstride * 16;   // sample of the batch
    ring_halo_generic_item(base, L, H, Nc, ptail, v_begin, t);
}

// fast path (needs H <= L): one thread per halo vector; a halo vector is the
// source vector with its lanes rotated by one, plus one or two patched lanes.
__global__ void ring_halo_fast_kernel(uint4* vec, int64_t L, int64_t H, int64_t Nc, int64_t rstride = 0)
{
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= 2 * H) return;
    vec += (size_t)blockIdx.y * (size_t)rstride;         // sample of the batch
    ring_halo_fast_item(vec, L, H, Nc, v);
}

// ---- NCCL through dlsym -------------------------------------------------------------------
namespace {
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
} g_nccl;

int nccl_load()
{
    if (g_nccl.lib) return B200MC_OK;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { snprintf(g_b200mc_err, sizeof(g_b200mc_err), "cannot load libnccl.so.2: %s", dlerror()); return B200MC_ERR_UNSUPPORTED; }
#define SYM(field, name) *(void**)(&g_nccl.field) = dlsym(lib, name); if (!g_nccl.field) { snprintf(g_b200mc_err, sizeof(g_b200mc_err), "libnccl: missing %s", name); return B200MC_ERR_UNSUPPORTED; }
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(Send, "ncclSend") SYM(Recv, "ncclRecv")
    SYM(AllReduce, "ncclAllReduce") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.lib = lib;
    return B200MC_OK;
}
#define NK(call)                                                                                      \
    do {                                                                                              \
        ncclResult_t r_ = (call);                                                                     \
        if (r_ != ncclSuccess) {                                                                      \
            snprintf(g_b200mc_err, sizeof(g_b200mc_err), "%s:%d: %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); \
            return B200MC_ERR_CUDA;                                                                   \
        }                                                                                             \
    } while (0)
}  // namespace

int dist_unique_id(char out[128])
{
    int rc = nccl_load();
    if (rc) return rc;
    ncclUniqueId id;
    NK(g_nccl.GetUniqueId(&id));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    memcpy(out, &id, 128);
    return B200MC_OK;
}
int dist_comm_init(void** comm, int rank, int nranks, const char idb[128])
{
    int rc = nccl_load();
    if (rc) return rc;
    ncclUniqueId id;
    memcpy(&id, idb, 128);
    ncclComm_t c;
    NK(g_nccl.CommInitRank(&c, nranks, id, rank));
    *comm = c;
    return B200MC_OK;
}
void dist_comm_destroy(void* comm) { if (comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)comm); }
int dist_allreduce_u64(void* comm, unsigned long long* buf, int n, cudaStream_t st)
{
    NK(g_nccl.AllReduce(buf, buf, (size_t)n, ncclUint64, ncclSum, (ncclComm_t)comm, st));
    return B200MC_OK;
}

int dist_allreduce_f64(void* comm, double* buf, int n, cudaStream_t st)
{
    NK(g_nccl.AllReduce(buf, buf, (size_t)n, ncclDouble, ncclSum, (ncclComm_t)comm, st));
    return B200MC_OK;
}
// One block each way around the ring of ranks: `first` -> rank-1 (which receives it as its `high`), `last` -> rank+1
// (its `low`).  The receives are posted in the order the two peers send (with two ranks both neighbours are the same
// peer and NCCL matches its sends in order: first, then last).
int dist_exchange_ring(void* comm, int rank, int nranks, const void* first, const void* last, void* low, void* high, size_t bytes, cudaStream_t st)
{
    const int prev = (rank + nranks - 1) % nranks, next = (rank + 1) % nranks;
    ncclComm_t c = (ncclComm_t)comm;
    NK(g_nccl.GroupStart());
    NK(g_nccl.Send(first, bytes, ncclUint8, prev, c, st));
    NK(g_nccl.Send(last, bytes, ncclUint8, next, c, st));
    NK(g_nccl.Recv(high, bytes, ncclUint8, next, c, st));   // next's first
    NK(g_nccl.Recv(low, bytes, ncclUint8, prev, c, st));    // prev's last
    NK(g_nccl.GroupEnd());
    return B200MC_OK;
}

// rotate every vector of a halo block by one byte-lane: dir = +1: lane b <- lane b-1 (lane 0 <- 15),
// dir = -1: lane b <- lane b+1 (lane 15 <- 0)
__global__ void ring_rotate_kernel(uint4* v, int64_t n, int dir)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 s = v[i];
    uint4 o;
    if (dir > 0) {
        o.x = __funnelshift_l(s.w, s.x, 8); o.y = __funnelshift_l(s.x, s.y, 8);
        o.z = __funnelshift_l(s.y, s.z, 8); o.w = __funnelshift_l(s.z, s.w, 8);
    } else {
        o.x = __funnelshift_r(s.x, s.y, 8); o.y = __funnelshift_r(s.y, s.z, 8);
        o.z = __funnelshift_r(s.z, s.w, 8); o.w = __funnelshift_r(s.w, s.x, 8);
    }
    v[i] = o;
}

// Slab halo exchange (one per colour pass): my first H owned vectors become the HIGH halo of rank-1,
// my last H owned vectors the LOW halo of rank+1 (ring of ranks).  Rank 0 / rank P-1 then rotate the
// received block by one lane: crossing the end of the fold moves a site to the next byte-lane.
static int ring_halo_dist(RingStore* s, int colour, cudaStream_t st)
{
    const RingGeom& g = s->g;
    ncclComm_t comm = (ncclComm_t)s->comm;
    const int prev = (g.rank + g.nranks - 1) % g.nranks, next = (g.rank + 1) % g.nranks;
    uint4* v = s->vec[colour];
    const size_t bytes = (size_t)g.H * sizeof(uint4);
    NK(g_nccl.GroupStart());
    NK(g_nccl.Send(v + g.H, bytes, ncclUint8, prev, comm, st));                 // first owned -> prev's high halo
    NK(g_nccl.Send(v + g.Lloc, bytes, ncclUint8, next, comm, st));              // last owned  -> next's low halo
    NK(g_nccl.Recv(v + g.H + g.Lloc, bytes, ncclUint8, next, comm, st));        // high halo <- next's first
    NK(g_nccl.Recv(v, bytes, ncclUint8, prev, comm, st));                       // low halo  <- prev's last
    NK(g_nccl.GroupEnd());
    if (g.rank == 0) {
        ring_rotate_kernel<<<(unsigned)((g.H + 255) / 256), 256, 0, st>>>(v, g.H, +1);
        COUNT_LAUNCH();
    }
    if (g.rank == g.nranks - 1) {
        ring_rotate_kernel<<<(unsigned)((g.H + 255) / 256), 256, 0, st>>>(v + g.H + g.Lloc, g.H, -1);
        COUNT_LAUNCH();
    }
    CK(cudaGetLastError());
    return B200MC_OK;
}

// ---- direct NVLink transport: IPC-mapped neighbour arrays ------------------------------------
int ring_p2p_export(RingStore* s, char out[RING_IPC_BYTES])
{
    static_assert(sizeof(cudaIpcMemHandle_t) * 3 == RING_IPC_BYTES, "cudaIpcMemHandle_t size");
    if (!s->flags) {
        CK(cudaMalloc(&s->flags, RING_FLAG_WORDS * sizeof(unsigned int)));
        CK(cudaMemset(s->flags, 0, RING_FLAG_WORDS * sizeof(unsigned int)));
    }
    cudaIpcMemHandle_t h[3];
    CK(cudaIpcGetMemHandle(&h[0], s->vec[0]));
    CK(cudaIpcGetMemHandle(&h[1], s->vec[1]));
    CK(cudaIpcGetMemHandle(&h[2], s->flags));
    memcpy(out, h, RING_IPC_BYTES);
    return B200MC_OK;
}

int ring_p2p_connect(RingStore* s, const char prev[RING_IPC_BYTES], const char next[RING_IPC_BYTES])
{
    const RingGeom& g = s->g;
    if (g.nranks < 2) ARG_FAIL("p2p_connect: not a slab handle");
    if (!s->flags) ARG_FAIL("p2p_connect: call p2p_handles first");
    const char* src[2] = {prev, next};
    const int nopen = (g.nranks == 2) ? 1 : 2;  // with two ranks both neighbours are the same process
    for (int side = 0; side < nopen; ++side) {
        cudaIpcMemHandle_t h[3];
        memcpy(h, src[side], RING_IPC_BYTES);
        void* ptr[3] = {nullptr, nullptr, nullptr};
        for (int j = 0; j < 3; ++j) {
            cudaError_t e = cudaIpcOpenMemHandle(&ptr[j], h[j], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                snprintf(g_b200mc_err, sizeof(g_b200mc_err), "cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
                cudaGetLastError();
                ring_p2p_close(s);
                return B200MC_ERR_UNSUPPORTED;
            }
            s->peer_maps[s->n_peer_maps++] = ptr[j];
        }
        s->peer_vec[side][0] = (uint4*)ptr[0];
        s->peer_vec[side][1] = (uint4*)ptr[1];
        s->peer_flags[side] = (unsigned int*)ptr[2];
    }
    if (nopen == 1) {
        s->peer_vec[1][0] = s->peer_vec[0][0];
        s->peer_vec[1][1] = s->peer_vec[0][1];
        s->peer_flags[1] = s->peer_flags[0];
    }
    const int prev_rank = (g.rank + g.nranks - 1) % g.nranks;
    const int64_t base = g.L / g.nranks, rem = g.L % g.nranks;
    s->Lloc_prev = base + (prev_rank < rem ? 1 : 0);
    s->p2p = true;
    return B200MC_OK;
}

// single GPU: the ring closes on itself, so the "neighbours" are this rank's own arrays -- the fused
// update + halo kernel then replaces the separate halo-refresh launch
int ring_p2p_connect_self(RingStore* s)
{
    const RingGeom& g = s->g;
    if (g.nranks != 1) ARG_FAIL("connect_self: slab handle");
    if (g.Nc % 16) ARG_FAIL("connect_self: sites per colour must be a multiple of 16");
    if (!s->flags) {
        CK(cudaMalloc(&s->flags, RING_FLAG_WORDS * sizeof(unsigned int)));
        CK(cudaMemset(s->flags, 0, RING_FLAG_WORDS * sizeof(unsigned int)));
    }
    for (int side = 0; side < 2; ++side) {
        s->peer_vec[side][0] = s->vec[0];
        s->peer_vec[side][1] = s->vec[1];
        s->peer_flags[side] = s->flags;
    }
    s->Lloc_prev = g.Lloc;
    s->p2p = true;
    return B200MC_OK;
}

void ring_p2p_close(RingStore* s)
{
    for (int i = 0; i < s->n_peer_maps; ++i) cudaIpcCloseMemHandle(s->peer_maps[i]);
    s->n_peer_maps = 0;
    for (int i = 0; i < s->n_sum_maps; ++i) cudaIpcCloseMemHandle(s->sum_maps[i]);
    s->n_sum_maps = 0;
    s->sums_p2p = false;
    s->p2p = false;
}

// ---- observable sums across the ranks over peer memory (no collective, no host round trip) --------------------
int ring_p2p_connect_sums(RingStore* s, const char* handles)
{
    const RingGeom& g = s->g;
    if (!s->p2p || g.nranks < 2) ARG_FAIL("p2p_connect_sums: call p2p_connect first");
    if (g.nranks > RING_MAX_SUM_RANKS) { snprintf(g_b200mc_err, sizeof(g_b200mc_err), "p2p_connect_sums: more than %d ranks", RING_MAX_SUM_RANKS); return B200MC_ERR_UNSUPPORTED; }
    const int prev = (g.rank + g.nranks - 1) % g.nranks, next = (g.rank + 1) % g.nranks;
    for (int r = 0; r < g.nranks; ++r) {
        if (r == g.rank) { s->all_flags[r] = s->flags; continue; }
        if (r == prev) { s->all_flags[r] = s->peer_flags[0]; continue; }   // already mapped by ring_p2p_connect
        if (r == next) { s->all_flags[r] = s->peer_flags[1]; continue; }
        cudaIpcMemHandle_t h[3];
        memcpy(h, handles + (size_t)r * RING_IPC_BYTES, RING_IPC_BYTES);
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h[2], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            snprintf(g_b200mc_err, sizeof(g_b200mc_err), "cudaIpcOpenMemHandle (flags of rank %d) failed: %s", r, cudaGetErrorString(e));
            cudaGetLastError();
            return B200MC_ERR_UNSUPPORTED;
        }
        s->sum_maps[s->n_sum_maps++] = ptr;
        s->all_flags[r] = (unsigned int*)ptr;
    }
    s->sums_p2p = true;
    return B200MC_OK;
}

struct RingSumArgs {
    unsigned int* box[RING_MAX_SUM_RANKS];   // every rank's flag buffer
    int rank, nranks, n;
    unsigned int seq;                        // this exchange (1, 2, ...): mailbox parity seq & 1
};

// One warp.  Lane r stores this rank's partial sums into rank r's mailbox entry [parity][my rank] (values, fence, then
// the sequence number with release semantics), then waits for entry [parity][r] of its OWN mailbox and reads rank r's
// partial sums; a warp reduction gives the totals.  Two parities suffice: a rank can only start exchange k + 2 after
// every rank has contributed to k + 1, i.e. after every rank has finished reading k.
__global__ void ring_sum_exchange_kernel(const __grid_constant__ RingSumArgs a, unsigned long long* buf, unsigned long long* host_out)
{
    const int lane = threadIdx.x;
    unsigned long long v[3] = {0ull, 0ull, 0ull};
    if (lane < a.nranks) {
        unsigned int* dst = a.box[lane] + RING_SUM_BASE + ((a.seq & 1u) * RING_MAX_SUM_RANKS + a.rank) * RING_SUM_SLOT;
        for (int j = 0; j < a.n; ++j) reinterpret_cast<volatile unsigned long long*>(dst)[j] = __ldcg(buf + j);
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst + 6), "r"(a.seq) : "memory");
        const unsigned int* src = a.box[a.rank] + RING_SUM_BASE + ((a.seq & 1u) * RING_MAX_SUM_RANKS + lane) * RING_SUM_SLOT;
        unsigned int got;
        do { asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(got) : "l"(src + 6) : "memory"); } while (got != a.seq);
        for (int j = 0; j < a.n; ++j) v[j] = reinterpret_cast<const volatile unsigned long long*>(src)[j];
    }
    for (int j = 0; j < 3; ++j)
        for (int o = 16; o > 0; o >>= 1) v[j] += __shfl_down_sync(0xffffffffu, v[j], o);
    if (lane == 0)
        for (int j = 0; j < a.n; ++j) { buf[j] = v[j]; if (host_out) host_out[j] = v[j]; }
}

int ring_sum_exchange(RingStore* s, unsigned long long* buf, int n, unsigned long long* host_out, cudaStream_t st)
{
    if (!s->sums_p2p) ARG_FAIL("sum exchange: peers not mapped");
    if (n < 1 || n > 3) ARG_FAIL("sum exchange: 1..3 values");
    RingSumArgs a;
    for (int r = 0; r < RING_MAX_SUM_RANKS; ++r) a.box[r] = r < s->g.nranks ? s->all_flags[r] : nullptr;
    a.rank = s->g.rank; a.nranks = s->g.nranks; a.n = n;
    a.seq = ++s->sum_seq;
    ring_sum_exchange_kernel<<<1, 32, 0, st>>>(a, buf, host_out);
    COUNT_LAUNCH();
    CK(cudaGetLastError());
    return B200MC_OK;
}

__global__ void ring_wait_flags_kernel(const unsigned int* flags, unsigned int seq)
{
    unsigned int v;
    do { asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags) : "memory"); } while ((int)(v - seq) < 0);
    do { asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + 16) : "memory"); } while ((int)(v - seq) < 0);
}

int ring_p2p_quiesce(RingStore* s, cudaStream_t st)
{
    if (!s->p2p || s->push_seq == 0) return B200MC_OK;
    ring_wait_flags_kernel<<<1, 1, 0, st>>>(s->flags, s->push_seq);
    COUNT_LAUNCH();
    CK(cudaGetLastError());
    return B200MC_OK;
}

int ring_halo(RingStore* s, int colour, cudaStream_t st)
{
    const RingGeom& g = s->g;
    if (g.nranks > 1) return ring_halo_dist(s, colour, st);
    uint8_t* base = reinterpret_cast<uint8_t*>(s->vec[colour]);
    const int64_t ntail = g.L - g.ptail;
    if (g.H <= g.L && g.ptail >= g.H) {
        const int64_t nv = 2 * g.H;
        ring_halo_fast_kernel<<<dim3((unsigned)((nv + 255) / 256), (unsigned)s->n_rep), 256, 0, st>>>(s->vec[colour], g.L, g.H, g.Nc, s->rstride);
        COUNT_LAUNCH();
        if (ntail > 0) {
            const int64_t n_items = ntail * 16;  // tail vectors are ordinals [2H, 2H + ntail)
            ring_halo_generic_kernel<<<dim3((unsigned)((n_items + 255) / 256), (unsigned)s->n_rep), 256, 0, st>>>(
                base, g.L, g.H, g.Nc, g.ptail, 2 * g.H, n_items, s->rstride);
            COUNT_LAUNCH();
        }
    } else {
        const int64_t n_items = (2 * g.H + ntail) * 16;
        ring_halo_generic_kernel<<<dim3((unsigned)((n_items + 255) / 256), (unsigned)s->n_rep), 256, 0, st>>>(base, g.L, g.H, g.Nc, g.ptail, 0, n_items, s->rstride);
        COUNT_LAUNCH();
    }
    CK(cudaGetLastError());
    return B200MC_OK;
}

// ---------------------------------------------------------------------------
// import / export in the reference layout spins(1-P : N+P), int32
// (src/ising3d_gpu_m.f90:232-236 `spins()` returns the raw array, halo included)
// ---------------------------------------------------------------------------
__global__ void ring_export_kernel(const uint8_t* a, const uint8_t* b, int64_t N, int64_t L,
                                   int64_t H, int64_t P, int64_t j0, int64_t n, int32_t* out,
                                   int map, int64_t p0, int64_t Lloc)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int64_t j = j0 + t;           // element of spins(1-P : N+P), 0-based
    const int64_t i = pos_mod(j - P, N);  // ring site
    const int64_t k = i >> 1;
    const int64_t lane = k / L, p = k - lane * L;
    if (p < p0 || p >= p0 + Lloc) { out[t] = INT32_MIN; return; }  // owned by another rank
    const uint8_t v = ((i & 1) ? b : a)[(p - p0 + H) * 16 + lane];
    out[t] = map == RING_MAP_PM1 ? 2 * (int32_t)v - 1 : (int32_t)v;
}

__global__ void ring_import_kernel(uint8_t* a, uint8_t* b, int64_t N, int64_t L, int64_t H,
                                   int64_t i0, int64_t n, const int32_t* in, int map, int64_t p0, int64_t Lloc)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int64_t i = i0 + t;  // ring site; in[] holds sites i0 .. i0+n-1
    const int64_t k = i >> 1;
    const int64_t lane = k / L, p = k - lane * L;
    if (p < p0 || p >= p0 + Lloc) return;  // owned by another rank
    int32_t v = in[t];
    if (map == RING_MAP_PM1) v = (v + 1) >> 1;
    ((i & 1) ? b : a)[(p - p0 + H) * 16 + lane] = (uint8_t)v;
}

int ring_import_i32(RingStore* s, const int32_t* host, RingValueMap map, cudaStream_t st, int rep, int32_t n_states)
{
    const RingGeom& g = s->g;
    if (rep < 0 || rep >= s->n_rep) ARG_FAIL("sample %d outside the batch of %d", rep, s->n_rep);
    // the byte-parallel kernels assume every stored value is a valid state (carries would cross lanes, table indices
    // would leave the table): reject anything else here.  Halo cells of the host array are ignored (rebuilt below).
    {
        const int32_t* v = host + g.P;
        int64_t bad = -1;
        if (map == RING_MAP_PM1) { for (int64_t i = 0; i < g.N; ++i) if (v[i] != 1 && v[i] != -1) { bad = i; break; } }
        else { for (int64_t i = 0; i < g.N; ++i) if ((uint32_t)v[i] >= (uint32_t)n_states) { bad = i; break; } }
        if (bad >= 0) {
            if (map == RING_MAP_PM1) ARG_FAIL("set_spins: site %lld holds %d (allowed: -1, +1)", (long long)(bad + 1), (int)v[bad]);
            ARG_FAIL("set_spins: site %lld holds %d (allowed: 0 .. %d)", (long long)(bad + 1), (int)v[bad], (int)n_states - 1);
        }
    }
    uint4* const v0 = s->vec[0] + (size_t)rep * s->rstride;
    uint4* const v1 = s->vec[1] + (size_t)rep * s->rstride;
    int rcq = ring_p2p_quiesce(s, st);
    if (rcq) return rcq;
    // interior only (the halo cells of the host array are ignored and rebuilt)
    for (int64_t i0 = 0; i0 < g.N; i0 += s->stage_elems) {
        const int64_t n = (g.N - i0 < s->stage_elems) ? g.N - i0 : s->stage_elems;
        CK(cudaMemcpyAsync(s->stage, host + g.P + i0, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        ring_import_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
            reinterpret_cast<uint8_t*>(v0), reinterpret_cast<uint8_t*>(v1), g.N, g.L,
            g.H, i0, n, s->stage, (int)map, g.p0, g.Lloc);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(st));
    }
    int rc = ring_halo(s, 0, st);
    if (rc) return rc;
    return ring_halo(s, 1, st);
}

int ring_export_i32(RingStore* s, int32_t* host, RingValueMap map, cudaStream_t st, int rep)
{
    const RingGeom& g = s->g;
    if (rep < 0 || rep >= s->n_rep) ARG_FAIL("sample %d outside the batch of %d", rep, s->n_rep);
    const uint4* const v0 = s->vec[0] + (size_t)rep * s->rstride;
    const uint4* const v1 = s->vec[1] + (size_t)rep * s->rstride;
    const int64_t total = g.N + 2 * g.P;
    for (int64_t j0 = 0; j0 < total; j0 += s->stage_elems) {
        const int64_t n = (total - j0 < s->stage_elems) ? total - j0 : s->stage_elems;
        ring_export_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
            reinterpret_cast<const uint8_t*>(v0), reinterpret_cast<const uint8_t*>(v1),
            g.N, g.L, g.H, g.P, j0, n, s->stage, (int)map, g.p0, g.Lloc);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(host + j0, s->stage, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    return B200MC_OK;
}
