// Folded-ring storage shared by the helical models (Ising 2D, Ising 3D, clock).
//
// The reference stores a helical lattice as ONE linear array spins(1-P : N+P)
// (P = nx in 2D, nx*ny in 3D) and colours sites by the parity of the linear
// index (src/ising3d_gpu_m.f90:60-62,196; src/ising2d_gpu_m.f90:52-54,155).
// With 0-based i = idx-1 the lattice is a ring of N sites, colour = i & 1,
// colour-site index k = i >> 1 (Nc = N/2 per colour), and with nx = 2h+1,
// nx*ny = 2g+1 the neighbours of colour-c site k are the OTHER colour's sites
//      k-1+c, k+c, k+h+c, k-h-1+c, (k+g+c, k-g-1+c)           (mod Nc)
//
// B200 layout: one int8 per site, the two colours in separate arrays, and each
// colour ring FOLDED into 16 byte-lanes of L = ceil(Nc/16) positions:
//      site k  ->  lane b = k / L,  position p = k % L,
// stored as byte b of the 128-bit vector p.  Every neighbour offset is then a
// whole-vector offset: all loads are aligned LDG.128 and the byte-parallel
// arithmetic never shifts data across lanes.  Lane wrap-around (position p+d
// beyond L belongs to lane b+1, and lane 15 wraps to lane 0) is materialised
// once per colour pass in H = max|offset| halo vectors on each side -- the
// counterpart of the reference's "norishiro" cells (src/ising3d_gpu_m.f90:
// 111-122).  If 16 does not divide Nc the last positions of the high lanes
// hold no site; they are kept equal to the ring continuation (so they behave
// as halo) and are masked out of the reductions.
#pragma once
#include "common.cuh"

#define RING_MAX_SUM_RANKS 16
#define RING_FLAG_WORDS 2048      // flag buffer: [0, 64) the push flags, [RING_SUM_BASE, ...) the sum mailboxes
#define RING_SUM_BASE 64
#define RING_SUM_SLOT 8           // words per mailbox entry: 3 x u64 values + seq (u32) + pad
struct RingGeom {
    int64_t N;       // sites on the ring
    int64_t Nc;      // sites per colour
    int64_t L;       // fold length = vectors per colour
    int64_t H;       // halo vectors on each side
    int64_t P;       // reference halo width (nx or nx*ny), for import/export
    int64_t ptail;   // first position that holds a non-site lane (== L if none)
    int64_t off[2][6];  // neighbour vector offsets per colour: x-,x+,y+,y-,z+,z-
    int nnb;         // 4 (2D) or 6 (3D)
    // slab decomposition: this rank owns positions [p0, p0 + Lloc) of every lane (single GPU: 0, L)
    int64_t p0, Lloc;
    int rank, nranks;
};

struct RingStore {
    RingGeom g;
    uint4* vec[2];   // [colour] -> (L + 2H) vectors; position p lives at index p + H
    int32_t* stage;  // staging buffer for import/export
    int64_t stage_elems;
    void* comm;      // ncclComm_t when nranks > 1
    // batch of independent samples (single GPU only): sample j of colour c starts at vec[c] + j * rstride
    int n_rep;
    int64_t rstride;  // vectors per sample and colour (Lloc + 2H)
    // direct NVLink transport (slab mode): the neighbours' colour arrays and flag words mapped with
    // cudaIpcOpenMemHandle; the boundary launch of a colour pass stores into them (ising_kernels.cuh, PUSH)
    bool p2p;
    unsigned int* flags;         // mine: [0] pushes received from rank-1, [16] from rank+1, [32] CTA counter, [48] debug wait ns (u64), [56] barrier scratch (u64)
    uint4* peer_vec[2][2];       // [0 = rank-1, 1 = rank+1][colour]
    unsigned int* peer_flags[2];
    void* peer_maps[6];          // what cudaIpcCloseMemHandle must be called on
    int n_peer_maps;
    int64_t Lloc_prev;           // owned vectors of rank-1 (its high halo starts at H + Lloc_prev)
    unsigned int push_seq;       // boundary pushes issued so far (same on every rank: SPMD)
    // observable sums across the ranks without a collective (slab mode, direct transport): every rank's flag buffer is
    // mapped on every rank; a one-warp kernel stores this rank's partial sums into each rank's mailbox and adds up the
    // nranks entries of its own (ring_sum_exchange)
    unsigned int* all_flags[RING_MAX_SUM_RANKS];   // [rank] -> that rank's flag buffer (own entry: flags)
    void* sum_maps[RING_MAX_SUM_RANKS];            // non-neighbour mappings to close
    int n_sum_maps;
    bool sums_p2p;
    unsigned int sum_seq;        // exchanges issued so far (same on every rank)
};
#define RING_IPC_BYTES 192       // three cudaIpcMemHandle_t: colour 0, colour 1, flags

enum RingValueMap { RING_MAP_IDENTITY = 0, RING_MAP_PM1 = 1 };  // PM1: stored 0/1 <-> -1/+1

// ---------------------------------------------------------------------------
// halo refresh, per work item (kernels: ring.cu; also called between the colour passes of the cooperative
// small-lattice sweep kernel, ising_kernels.cuh)
// ---------------------------------------------------------------------------
__device__ __forceinline__ int64_t pos_mod(int64_t a, int64_t m)
{
    if (m < (int64_t)0x40000000 && a > -(int64_t)0x40000000 && a < (int64_t)0x40000000) {   // small rings: 32-bit division
        const int r32 = (int)a % (int)m;
        return r32 < 0 ? r32 + (int)m : r32;
    }
    int64_t r = a % m;
    return r < 0 ? r + m : r;
}

// value of ring site k (of this colour) read from its owning (lane, position)
__device__ __forceinline__ uint8_t ring_site(const uint8_t* base, int64_t L, int64_t H, int64_t k)
{
    int64_t b, p;
    if (k < (int64_t)0x7FFFFFFF) { const unsigned b32 = (unsigned)k / (unsigned)L; b = b32; p = (unsigned)k - b32 * (unsigned)L; }
    else { b = k / L; p = k - b * L; }
    return __ldcg(base + (p + H) * 16 + b);   // L2: also called between the passes of the cooperative sweep kernel
}

// generic, byte-granular: item t = (dirty vector, lane).  Dirty vectors are the 2H halo vectors and the tail
// positions [ptail, L).
__device__ __forceinline__ void ring_halo_generic_item(uint8_t* base, int64_t L, int64_t H, int64_t Nc, int64_t ptail, int64_t v_begin, int64_t t)
{
    const int b = (int)(t & 15);
    int64_t v = v_begin + (t >> 4);  // dirty-vector ordinal
    int64_t p;
    if (v < H) p = v - H;
    else if (v < 2 * H) p = L + (v - H);
    else p = ptail + (v - 2 * H);
    const int64_t kraw = (int64_t)b * L + p;
    if (p >= 0 && p < L && kraw < Nc) return;  // a real site: owned, not a copy
    base[(p + H) * 16 + b] = ring_site(base, L, H, pos_mod(kraw, Nc));
}

// The vector a halo position holds (fast path: needs H <= L and ptail >= H): the source vector with its lanes rotated
// by one, plus one or two patched lanes.  Reads real sites only.  vec: index 0 = position -H.
__device__ __forceinline__ uint4 ring_wrapped_low(const uint4* vec, int64_t L, int64_t H, int64_t Nc, int64_t p)
{
    // p < 0: lane b <- lane b-1 at p + L; lane 0 <- site Nc + p
    const uint8_t* base = reinterpret_cast<const uint8_t*>(vec);
    const uint4 s = __ldcg(vec + (p + L + H));
    uint4 o;
    o.w = __funnelshift_l(s.z, s.w, 8);
    o.z = __funnelshift_l(s.y, s.z, 8);
    o.y = __funnelshift_l(s.x, s.y, 8);
    o.x = (s.x << 8) | ring_site(base, L, H, pos_mod(p, Nc));
    return o;
}
__device__ __forceinline__ uint4 ring_wrapped_high(const uint4* vec, int64_t L, int64_t H, int64_t Nc, int64_t p)
{
    // p >= L: lane b <- lane b+1 at p - L; lanes 14, 15 patched
    const uint8_t* base = reinterpret_cast<const uint8_t*>(vec);
    const uint4 s = __ldcg(vec + (p - L + H));
    uint4 o;
    o.x = __funnelshift_r(s.x, s.y, 8);
    o.y = __funnelshift_r(s.y, s.z, 8);
    o.z = __funnelshift_r(s.z, s.w, 8);
    const uint32_t b14 = ring_site(base, L, H, pos_mod(14 * L + p, Nc));
    const uint32_t b15 = ring_site(base, L, H, pos_mod(15 * L + p, Nc));
    o.w = ((s.w >> 8) & 0x0000FFFFu) | (b14 << 16) | (b15 << 24);
    return o;
}
__device__ __forceinline__ void ring_halo_fast_item(uint4* vec, int64_t L, int64_t H, int64_t Nc, int64_t v)
{
    if (v < H) vec[v] = ring_wrapped_low(vec, L, H, Nc, v - H);
    else vec[L + v] = ring_wrapped_high(vec, L, H, Nc, L + (v - H));   // position L + (v - H) lives at index L + v
}

// The vector position p in [-H, L + H) SHOULD hold, computed from the owned sites alone (the stored halo vectors and
// tail lanes are not read): what a colour pass of the cooperative sweep kernel loads for a neighbour position outside
// [0, ptail), so that no halo refresh (and no second grid barrier) is needed between its passes.
__device__ __forceinline__ uint4 ring_load_wrapped(const uint4* vec, int64_t L, int64_t H, int64_t Nc, int64_t ptail, bool fast, int64_t p)
{
    if (fast) {
        if (p < 0) return ring_wrapped_low(vec, L, H, Nc, p);
        if (p >= L) return ring_wrapped_high(vec, L, H, Nc, p);
        uint4 s = __ldcg(vec + (p + H));   // ptail <= p < L: lane 15 holds no site here
        const uint32_t b15 = ring_site(reinterpret_cast<const uint8_t*>(vec), L, H, pos_mod(15 * L + p, Nc));
        s.w = (s.w & 0x00FFFFFFu) | (b15 << 24);
        return s;
    }
    const uint8_t* base = reinterpret_cast<const uint8_t*>(vec);
    uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll 1
    for (int b = 0; b < 16; ++b) {
        const int64_t kraw = (int64_t)b * L + p;
        const bool real = p >= 0 && p < L && kraw < Nc;
        const uint32_t val = real ? (uint32_t)__ldcg(base + (p + H) * 16 + b) : (uint32_t)ring_site(base, L, H, pos_mod(kraw, Nc));
        w[b >> 2] |= val << (8 * (b & 3));
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

int ring_geom_init(RingGeom* g, int64_t nx, int64_t ny, int64_t nz /*0 for 2D*/);
// split the fold across ranks (needs 16 | Nc and Lloc >= H on every rank)
int ring_geom_set_slab(RingGeom* g, int rank, int nranks);
int ring_alloc(RingStore* s);
void ring_free(RingStore* s);
int ring_fill(RingStore* s, uint8_t value, cudaStream_t st);
int ring_halo(RingStore* s, int colour, cudaStream_t st);
// host int32 arrays in the reference layout spins(1-P : N+P)
// n_states: valid stored values are 0 .. n_states-1 (RING_MAP_PM1: the host values must be -1 / +1); anything else is B200MC_ERR_ARG
int ring_import_i32(RingStore* s, const int32_t* host, RingValueMap map, cudaStream_t st, int rep = 0, int32_t n_states = 2);
int ring_export_i32(RingStore* s, int32_t* host, RingValueMap map, cudaStream_t st, int rep = 0);

// direct transport set-up: export my handles, map the two neighbours' (prev == next when nranks == 2)
int ring_p2p_export(RingStore* s, char out[RING_IPC_BYTES]);
int ring_p2p_connect(RingStore* s, const char prev[RING_IPC_BYTES], const char next[RING_IPC_BYTES]);
int ring_p2p_connect_self(RingStore* s);
// map the flag buffers of all ranks (handles: nranks x RING_IPC_BYTES, as exported by ring_p2p_export; call after
// ring_p2p_connect): enables ring_sum_exchange
int ring_p2p_connect_sums(RingStore* s, const char* handles);
// buf[0..n) (n <= 3 u64 partial sums of this rank, device memory) -> sums over all ranks, in buf and in host_out (pinned)
int ring_sum_exchange(RingStore* s, unsigned long long* buf, int n, unsigned long long* host_out, cudaStream_t st);
void ring_p2p_close(RingStore* s);
// stream-ordered wait until every push issued so far by both neighbours has landed
int ring_p2p_quiesce(RingStore* s, cudaStream_t st);

// NCCL, resolved at run time with dlopen("libnccl.so.2") so that single-GPU use has no NCCL dependency
int dist_unique_id(char out[128]);
int dist_comm_init(void** comm, int rank, int nranks, const char id[128]);
void dist_comm_destroy(void* comm);
int dist_allreduce_u64(void* comm, unsigned long long* buf, int n, cudaStream_t st);
int dist_allreduce_f64(void* comm, double* buf, int n, cudaStream_t st);
int dist_exchange_ring(void* comm, int rank, int nranks, const void* first, const void* last, void* low, void* high, size_t bytes, cudaStream_t st);
