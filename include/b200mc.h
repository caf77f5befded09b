/* b200mc -- C ABI of the B200-native checkerboard spin Monte Carlo library
 * (libb200mc.so, built from cuda_fortran_mc_simulation_spin_b200/csrc).
 *
 * This is the drop-in boundary for the hot path of
 * osada-yum/CUDA_Fortran_MC_simulation_spin.  The reference has no FFI: its
 * callers (the app/ *_relaxation.f90 programs) `use` a Fortran module and call type-bound
 * procedures of a derived type.  Every entry point below replaces one such
 * procedure (cited as file:line of /root/reference) and is what the
 * ISO_C_BINDING shim modules in cuda_fortran_mc_simulation_spin_b200/fortran/
 * bind.
 *
 * Conventions
 *  - plain C types only; `void*` opaque handles own all device memory;
 *  - every function returns 0 on success, non-zero on error (B200MC_ERR_*);
 *    b200mc_last_error() returns the message of the last failure on the
 *    calling thread.  The reference stores cuRAND/CUDA codes in a public
 *    `*_stat` variable and never checks them (src/ising2d_gpu_m.f90:8); the
 *    Fortran shims mirror our return code into that same variable;
 *  - one handle is used by one host thread at a time;
 *  - host arrays use the REFERENCE's layouts (halo cells included);
 *  - `update` is asynchronous on the handle's stream; every function that
 *    returns data to the host synchronises.
 *  - there is no CPU fallback: without a CUDA device `create` fails.
 */
#ifndef B200MC_H
#define B200MC_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
#pragma GCC visibility push(default)

#define B200MC_OK 0
#define B200MC_ERR_ARG 1
#define B200MC_ERR_CUDA 2
#define B200MC_ERR_STATE 3
#define B200MC_ERR_UNSUPPORTED 4

#define B200MC_METROPOLIS 0 /* the reference's only Ising update */
#define B200MC_HEATBATH 1   /* north-star addition, no reference symbol */

const char* b200mc_last_error(void);
/* the same message on stderr (what the Fortran shims call before `error stop`: no C string handling on their side) */
void b200mc_print_last_error(void);
int b200mc_version(void);
/* number of CUDA kernels this library has launched so far in this process (all handles) */
unsigned long long b200mc_launch_count(void);
unsigned long long b200mc_debug_slab_wait_ns(void* h); /* slab mode: ns spent waiting for neighbour flags (block 0) */
int b200mc_debug_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]); /* device Philox, for KAT tests */

/* ------------------------------------------------------------------------
 * Ising 3D -- type(ising3d_gpu), src/ising3d_gpu_m.f90:15-48
 * ------------------------------------------------------------------------ */
/* init, :50-71.  Helical boundary; nx, ny odd and nz even are REQUIRED (the
 * reference silently races otherwise) -> B200MC_ERR_ARG. */
int b200mc_ising3d_create(void** h, int64_t nx, int64_t ny, int64_t nz, double kbt, int32_t iseed);
int b200mc_ising3d_destroy(void* h);
int b200mc_ising3d_set_stream(void* h, void* cuda_stream);
/* skip_curand, :72-77: n_skip counts uniforms; advances the draw counter by ceil(n_skip / nall) */
int b200mc_ising3d_skip_curand(void* h, int64_t n_skip);
int b200mc_ising3d_set_allup_spin(void* h);              /* :79-82 */
int b200mc_ising3d_set_random_spin(void* h);             /* :84-100 */
int b200mc_ising3d_set_kbt(void* h, double kbt);         /* :124-128 */
int b200mc_ising3d_set_beta(void* h, double beta);       /* :130-135 (+ update_ws :138-172) */
int b200mc_ising3d_set_method(void* h, int32_t method);  /* B200MC_METROPOLIS (default) | B200MC_HEATBATH */
int b200mc_ising3d_update(void* h);                      /* one MCS, :174-206 */
int b200mc_ising3d_update_n(void* h, int32_t n_sweeps);  /* n MCS back to back */
/* one MCS reading uniforms from a host array randoms(1:nall) in the reference's
 * index order instead of the built-in generator (parity / cuRAND-stream mode) */
int b200mc_ising3d_update_with_randoms(void* h, const double* randoms);
int b200mc_ising3d_calc_energy_sum(void* h, int64_t* e); /* :239-257 */
int b200mc_ising3d_calc_magne_sum(void* h, int64_t* m);  /* :259-276 */
int b200mc_ising3d_measure(void* h, int64_t* e, int64_t* m); /* both, one pass */
/* the drivers' inner loop (app/ising3d_gpu_relaxation.f90:40-46) on the device: mcs x [update; calc_magne_sum;
 * calc_energy_sum]; e[i], m[i] = the sums after MCS i+1 (either may be NULL).  One host synchronisation in all. */
int b200mc_ising3d_run_relaxation(void* h, int32_t mcs, int64_t* e, int64_t* m);
/* the drivers' whole measurement (app/ising3d_gpu_relaxation.f90:37-55) on the device: tot_sample x [set_allup_spin (or
 * set_random_spin when random_start != 0); mcs x (update; calc_magne_sum; calc_energy_sum; add_data(m / N, e / N))] with the
 * Kahan mean / variance / covariance accumulators of the drivers' `variance_covariance_kahan` kept per MCS on the device
 * (a batch handle runs n_multi samples per pass; tot_sample must be a multiple of n_multi).  One host synchronisation.
 * out[8 * i + k], i = 0 .. mcs-1: k = 0 num_sample, 1 mean1 (<m>), 2 mean2 (<e>), 3 square_mean1, 4 square_mean2,
 * 5 var1, 6 var2, 7 cov (unbiased; the table prints them times nall, :52-54). */
int b200mc_ising3d_run_relaxation_stats(void* h, int32_t mcs, int32_t tot_sample, int32_t random_start, double* out);
/* one row of the drivers' output table, :49-55 (columns blank-separated, reals with 17 significant digits) */
int b200mc_format_relaxation_row(int64_t nall, int32_t i, const double row[8], char* buf, int32_t buflen);
/* spins(), :232-236: int32 0/1, layout spins(1-nxy : nall+nxy) -> nall + 2 nxy elements */
int b200mc_ising3d_get_spins(void* h, int32_t* out);
int b200mc_ising3d_set_spins(void* h, const int32_t* in); /* inverse (halo cells of `in` ignored) */
int64_t b200mc_ising3d_nx(void* h);   /* :208-223 */
int64_t b200mc_ising3d_ny(void* h);
int64_t b200mc_ising3d_nz(void* h);
int64_t b200mc_ising3d_nall(void* h);
double b200mc_ising3d_kbt(void* h);   /* :224-227 */
double b200mc_ising3d_beta(void* h);  /* :228-231 */
/* host copy of ws(0:6, 0:1) as built by update_ws (:153-171): out[S + 7*s] */
int b200mc_ising3d_get_ws(void* h, double out[14]);
/* per-launch device timing of the colour-pass kernel (CUDA events on the handle's stream):
 * set_timing(1) resets and enables; get_timing returns the launches seen and their summed duration */
int b200mc_ising3d_set_timing(void* h, int32_t on);
int b200mc_ising3d_get_timing(void* h, int64_t* launches, double* total_ms);
int b200mc_ising3d_sync(void* h);

/* ------------------------------------------------------------------------
 * Batch of independent samples (single GPU): n_multi lattices with the same parameters, updated by the same
 * launches (one launch per colour pass for the whole batch) -- the drivers' `do sample = 1, tot_sample` loop
 * (app/ising3d_gpu_relaxation.f90:38) n_multi samples at a time, which is what fills a B200 at the reference's
 * default lattice sizes (~10^6 sites).  Not a reference symbol for the Ising types (the reference batches only its
 * clock model, src/clock_gpu_multi_m.f90).  Sample j draws the counters (position, j, ...): sample 0 of a batch
 * equals the plain handle.  update / update_n / set_* act on all samples; calc_*_sum / measure / get_spins /
 * set_spins act on sample 0; run_relaxation fills e[j * mcs + i], m[j * mcs + i] (sample-major).
 * ------------------------------------------------------------------------ */
int b200mc_ising3d_create_multi(void** h, int64_t nx, int64_t ny, int64_t nz, double kbt, int32_t iseed, int32_t n_multi);
int b200mc_ising2d_create_multi(void** h, int64_t nx, int64_t ny, double kbt, int32_t iseed, int32_t n_multi);
int32_t b200mc_ising3d_n_multi(void* h);
int32_t b200mc_ising2d_n_multi(void* h);
int b200mc_ising3d_measure_multi(void* h, int64_t* e, int64_t* m);   /* n_multi values each */
int b200mc_ising2d_measure_multi(void* h, int64_t* e, int64_t* m);
int b200mc_ising3d_get_spins_multi(void* h, int32_t sample, int32_t* out);
int b200mc_ising2d_get_spins_multi(void* h, int32_t sample, int32_t* out);
int b200mc_ising3d_set_spins_multi(void* h, int32_t sample, const int32_t* in);
int b200mc_ising2d_set_spins_multi(void* h, int32_t sample, const int32_t* in);

/* ------------------------------------------------------------------------
 * Slab decomposition over several GPUs, one process (rank) per GPU (the
 * reference is single-GPU, its `norishiro` halo, src/ising3d_gpu_m.f90:102-122, is the
 * cell set a rank boundary exchanges).  The helical lattice is a ring of linear indices;
 * every rank owns an equal share of each of the 16 byte-lanes of the folded ring, so a
 * site's (lane, position) -- and with it its random stream -- does not depend on the
 * number of ranks: an N-rank run is bit-identical to the 1-GPU run of the same lattice.
 * After each colour pass the first / last H owned vectors go to the neighbouring ranks
 * (NCCL send/recv on a second stream, overlapped with the interior launch); observables
 * are all-reduced, every rank gets the global value.  get_spins returns INT32_MIN for
 * sites owned by another rank (merge with an elementwise max); set_spins takes the
 * full array on every rank.  Needs (sites per colour) % 16 == 0.
 * ------------------------------------------------------------------------ */
/* NCCL unique id of the job: call on rank 0, broadcast the 128 bytes to the other ranks */
int b200mc_dist_unique_id(char out[128]);
int b200mc_ising3d_create_slab(void** h, int64_t nx, int64_t ny, int64_t nz, double kbt, int32_t iseed,
                               int32_t rank, int32_t nranks, const char nccl_id[128]);
int b200mc_ising2d_create_slab(void** h, int64_t nx, int64_t ny, double kbt, int32_t iseed,
                               int32_t rank, int32_t nranks, const char nccl_id[128]);
/* optional direct NVLink transport for the per-pass halo exchange: every rank exports three
 * cudaIpcMemHandle_t (192 bytes), the caller hands each rank the bytes of rank-1 and rank+1 (ring);
 * afterwards the boundary launch of a colour pass stores its results straight into the neighbours'
 * halo cells (one fused kernel: update + transfer) instead of NCCL send/recv.  B200MC_ERR_UNSUPPORTED
 * if CUDA IPC is not available between the two processes (keep using the NCCL transport then --
 * all ranks must take the same decision). */
int b200mc_ising3d_p2p_handles(void* h, char out[192]);
int b200mc_ising2d_p2p_handles(void* h, char out[192]);
int b200mc_ising3d_p2p_connect(void* h, const char prev[192], const char next[192]);
int b200mc_ising2d_p2p_connect(void* h, const char prev[192], const char next[192]);
/* after p2p_connect: map the flag buffers of ALL ranks (handles = nranks x 192 bytes, rank order, as written by
 * p2p_handles).  calc_energy_sum / calc_magne_sum then add the per-rank sums up through peer memory (a one-warp kernel
 * that stores its partial sums into every rank's mailbox and writes the totals to pinned host memory) instead of
 * ncclAllReduce + copy.  Optional: without it the NCCL all-reduce is used. */
int b200mc_ising3d_p2p_connect_sums(void* h, const char* handles);
int b200mc_ising2d_p2p_connect_sums(void* h, const char* handles);
int b200mc_ising3d_rank_info(void* h, int32_t* rank, int32_t* nranks);
int b200mc_ising2d_rank_info(void* h, int32_t* rank, int32_t* nranks);
/* host-only: the slab a rank would own. out = {Nc, L, H, p0, Lloc, ptail} (nz = 0 for 2D) */
int b200mc_ring_slab_geometry(int64_t nx, int64_t ny, int64_t nz, int32_t rank, int32_t nranks, int64_t out[6]);

/* ------------------------------------------------------------------------
 * Bit-packed (multi-spin coded) Ising 3D / 2D: one bit per site, Metropolis, one GPU.  The same procedures of
 * type(ising3d_gpu) / type(ising2d_gpu) (file:line as for b200mc_ising3d_* / b200mc_ising2d_* above and below), the same
 * host layouts and conventions; its own random stream (oracle/rng_contract.c, orc_isingbits_uniforms).  Shapes: helical
 * validity as for the int8 handles, and nx ny nz / 2 a multiple of 128 (e.g. the last extent a multiple of 256) with the
 * fold not shorter than the halo -> B200MC_ERR_UNSUPPORTED otherwise.
 * ------------------------------------------------------------------------ */
int b200mc_ising3dp_create(void** h, int64_t nx, int64_t ny, int64_t nz, double kbt, int32_t iseed);
/* slab mode (one process per GPU, as b200mc_ising3d_create_slab): every rank owns an equal share of every bit-lane, so an
 * N-rank run is bit-identical to the 1-GPU run; halos through ncclSend/Recv per colour pass, observables all-reduced */
int b200mc_ising3dp_create_slab(void** h, int64_t nx, int64_t ny, int64_t nz, double kbt, int32_t iseed, int32_t rank, int32_t nranks, const char* nccl_id);
int b200mc_ising2dp_create_slab(void** h, int64_t nx, int64_t ny, double kbt, int32_t iseed, int32_t rank, int32_t nranks, const char* nccl_id);
int b200mc_ising3dp_rank_info(void* h, int32_t* rank, int32_t* nranks);
int b200mc_ising2dp_rank_info(void* h, int32_t* rank, int32_t* nranks);
int b200mc_ising3dp_destroy(void* h);
int b200mc_ising3dp_set_stream(void* h, void* cuda_stream);
int b200mc_ising3dp_skip_curand(void* h, int64_t n_skip);
int b200mc_ising3dp_set_allup_spin(void* h);
int b200mc_ising3dp_set_random_spin(void* h);
int b200mc_ising3dp_set_kbt(void* h, double kbt);
int b200mc_ising3dp_set_beta(void* h, double beta);
int b200mc_ising3dp_update(void* h);
int b200mc_ising3dp_update_n(void* h, int32_t n_sweeps);
int b200mc_ising3dp_calc_energy_sum(void* h, int64_t* e);
int b200mc_ising3dp_calc_magne_sum(void* h, int64_t* m);
int b200mc_ising3dp_measure(void* h, int64_t* e, int64_t* m);
int b200mc_ising3dp_get_spins(void* h, int32_t* out);      /* spins(1-nxy : nall+nxy), int32 0/1 */
int b200mc_ising3dp_set_spins(void* h, const int32_t* in);
int b200mc_ising3dp_get_ws(void* h, double out[14]);
int64_t b200mc_ising3dp_nx(void* h);
int64_t b200mc_ising3dp_ny(void* h);
int64_t b200mc_ising3dp_nz(void* h);
int64_t b200mc_ising3dp_nall(void* h);
double b200mc_ising3dp_kbt(void* h);
double b200mc_ising3dp_beta(void* h);
int b200mc_ising3dp_sync(void* h);
int b200mc_ising3dp_set_timing(void* h, int32_t on);       /* CUDA events around every colour-pass launch */
int b200mc_ising3dp_get_timing(void* h, int64_t* launches, double* total_ms);
int b200mc_ising2dp_create(void** h, int64_t nx, int64_t ny, double kbt, int32_t iseed);
int b200mc_ising2dp_destroy(void* h);
int b200mc_ising2dp_set_stream(void* h, void* cuda_stream);
int b200mc_ising2dp_skip_curand(void* h, int64_t n_skip);
int b200mc_ising2dp_set_allup_spin(void* h);
int b200mc_ising2dp_set_random_spin(void* h);
int b200mc_ising2dp_set_kbt(void* h, double kbt);
int b200mc_ising2dp_set_beta(void* h, double beta);
int b200mc_ising2dp_update(void* h);
int b200mc_ising2dp_update_n(void* h, int32_t n_sweeps);
int b200mc_ising2dp_calc_energy_sum(void* h, int64_t* e);
int b200mc_ising2dp_calc_magne_sum(void* h, int64_t* m);
int b200mc_ising2dp_measure(void* h, int64_t* e, int64_t* m);
int b200mc_ising2dp_get_spins(void* h, int32_t* out);      /* spins(1-nx : nall+nx), int32 -1/+1 */
int b200mc_ising2dp_set_spins(void* h, const int32_t* in);
int64_t b200mc_ising2dp_nx(void* h);
int64_t b200mc_ising2dp_ny(void* h);
int64_t b200mc_ising2dp_nall(void* h);
double b200mc_ising2dp_kbt(void* h);
double b200mc_ising2dp_beta(void* h);
int b200mc_ising2dp_sync(void* h);
int b200mc_ising2dp_set_timing(void* h, int32_t on);
int b200mc_ising2dp_get_timing(void* h, int64_t* launches, double* total_ms);

/* ------------------------------------------------------------------------
 * Ising 2D / 3D with TRUE PERIODIC boundaries (torus), int8, Metropolis / heat-bath; one GPU or (3D) slabs of planes.
 * No reference module: the reference's Ising types are helical and valid for odd nx only
 * (src/ising3d_gpu_m.f90:60-62,196, src/ising2d_gpu_m.f90:56-58), so L = 1024^3 -- the size
 * BASELINE.json's north_star and BASELINE.md C2 / C1 / C5 name ("1024^3 periodic", "1024^2 periodic",
 * "65536^2 periodic") -- needs this boundary condition.  The procedures are those of type(ising3d_gpu) /
 * type(ising2d_gpu) (file:line as for b200mc_ising3d_* / b200mc_ising2d_* above), with the reference's update
 * rule, tables, value conventions (3D 0 / 1, 2D -1 / +1) and observables; colour = (x + y + z) & 1, colour 0
 * first.  ndim = 2 | 3 (nz ignored in 2D); nx % 32 == 0, ny and nz even.  Host arrays: s[x + nx (y + ny z)],
 * no halo cells.  CPU restatement: oracle/oracle.c orc_isingp_*.
 * ------------------------------------------------------------------------ */
int b200mc_ising_torus_create(void** h, int32_t ndim, int64_t nx, int64_t ny, int64_t nz, double kbt, int32_t iseed);
/* slab mode (one process per GPU, as b200mc_ising3d_create_slab): the nz planes are split over the ranks (3D, nx % 1024 == 0,
 * nz % nranks == 0), a rank's ghost planes arrive by ncclSend/Recv after the boundary planes of a colour pass while the
 * interior planes are updated, observables are all-reduced.  A site's random stream does not depend on the number of ranks:
 * an N-rank run is the 1-GPU run of the same lattice.  get_spins / set_spins then move the rank's own planes
 * (nx ny nz_local values, planes z0 .. z0 + nz_local - 1); nall / nz report the whole lattice. */
int b200mc_ising_torus_create_slab(void** h, int32_t ndim, int64_t nx, int64_t ny, int64_t nz, double kbt, int32_t iseed, int32_t rank, int32_t nranks, const char nccl_id[128]);
int b200mc_ising_torus_rank_info(void* h, int32_t* rank, int32_t* nranks, int64_t* z0, int64_t* nz_local);
/* direct transport for the ghost planes (as b200mc_ising3d_p2p_handles / _p2p_connect): export this rank's CUDA IPC handles
 * (192 bytes), exchange them by any means, map the two neighbours'; the ghost planes are then stored straight into the
 * neighbours' arrays over NVLink by a small kernel after every colour pass instead of travelling by ncclSend/Recv.
 * Either every rank connects or none does. */
int b200mc_ising_torus_p2p_handles(void* h, char out[192]);
int b200mc_ising_torus_p2p_connect(void* h, const char prev[192], const char next[192]);
int b200mc_ising_torus_destroy(void* h);
int b200mc_ising_torus_set_stream(void* h, void* cuda_stream);
int b200mc_ising_torus_skip_curand(void* h, int64_t n_skip);     /* src/ising3d_gpu_m.f90:72-77 */
int b200mc_ising_torus_set_allup_spin(void* h);                  /* :79-82 */
int b200mc_ising_torus_set_random_spin(void* h);                 /* :84-100 */
int b200mc_ising_torus_set_kbt(void* h, double kbt);             /* :124-128 */
int b200mc_ising_torus_set_beta(void* h, double beta);           /* :130-135 */
int b200mc_ising_torus_set_method(void* h, int32_t method);      /* B200MC_METROPOLIS (default) | B200MC_HEATBATH */
int b200mc_ising_torus_update(void* h);                          /* one MCS, :174-206 */
int b200mc_ising_torus_update_n(void* h, int32_t n_sweeps);
/* one MCS with the uniforms of a caller array indexed like the spins (the reference's randoms(idx)), real64 compare */
int b200mc_ising_torus_update_with_randoms(void* h, const double* randoms);
int b200mc_ising_torus_calc_energy_sum(void* h, int64_t* e);     /* :239-257 / src/ising2d_gpu_m.f90:198-213 */
int b200mc_ising_torus_calc_magne_sum(void* h, int64_t* m);      /* :259-276 / src/ising2d_gpu_m.f90:215-228 */
int b200mc_ising_torus_measure(void* h, int64_t* e, int64_t* m);
int b200mc_ising_torus_get_spins(void* h, int32_t* out);         /* nx ny nz values */
int b200mc_ising_torus_set_spins(void* h, const int32_t* in);
int64_t b200mc_ising_torus_nx(void* h);                          /* :208-231 */
int64_t b200mc_ising_torus_ny(void* h);
int64_t b200mc_ising_torus_nz(void* h);                          /* 0 for the 2D model */
int64_t b200mc_ising_torus_nall(void* h);
double b200mc_ising_torus_kbt(void* h);
double b200mc_ising_torus_beta(void* h);
/* the acceptance table w[s * 8 + S] (s = 0 / 1 the site's spin, S = number of up neighbours) as the reference builds it */
int b200mc_ising_torus_get_table(void* h, double out[16]);
int b200mc_ising_torus_set_timing(void* h, int32_t on);
int b200mc_ising_torus_get_timing(void* h, int64_t* launches, double* total_ms);
int b200mc_ising_torus_sync(void* h);


/* ------------------------------------------------------------------------
 * Ising 2D -- type(ising2d_gpu), src/ising2d_gpu_m.f90:12-42
 * ------------------------------------------------------------------------ */
/* init, :44-61.  nx odd, ny even REQUIRED. */
int b200mc_ising2d_create(void** h, int64_t nx, int64_t ny, double kbt, int32_t iseed);
int b200mc_ising2d_destroy(void* h);
int b200mc_ising2d_set_stream(void* h, void* cuda_stream);
int b200mc_ising2d_skip_curand(void* h, int64_t n_skip); /* not in the 2D reference type; same meaning as 3D */
int b200mc_ising2d_set_allup_spin(void* h);              /* :63-66 */
int b200mc_ising2d_set_random_spin(void* h);             /* :68-84 */
int b200mc_ising2d_set_kbt(void* h, double kbt);         /* :108-112 */
int b200mc_ising2d_set_beta(void* h, double beta);       /* :114-119 (+ update_exparr :122-131) */
int b200mc_ising2d_set_method(void* h, int32_t method);
int b200mc_ising2d_update(void* h);                      /* :133-162 */
int b200mc_ising2d_update_n(void* h, int32_t n_sweeps);
int b200mc_ising2d_update_with_randoms(void* h, const double* randoms);
int b200mc_ising2d_calc_energy_sum(void* h, int64_t* e); /* :198-212 */
int b200mc_ising2d_calc_magne_sum(void* h, int64_t* m);  /* :214-228 */
int b200mc_ising2d_measure(void* h, int64_t* e, int64_t* m);
int b200mc_ising2d_run_relaxation(void* h, int32_t mcs, int64_t* e, int64_t* m); /* app/ising2d_gpu_relaxation.f90:38-43 */
int b200mc_ising2d_run_relaxation_stats(void* h, int32_t mcs, int32_t tot_sample, int32_t random_start, double* out); /* :33-52, as for 3D */
/* spins(), :184-188: int32 +1/-1, layout spins(1-nx : nall+nx) -> nall + 2 nx elements */
int b200mc_ising2d_get_spins(void* h, int32_t* out);
int b200mc_ising2d_set_spins(void* h, const int32_t* in);
int64_t b200mc_ising2d_nx(void* h);
int64_t b200mc_ising2d_ny(void* h);
int64_t b200mc_ising2d_nall(void* h);
double b200mc_ising2d_kbt(void* h);
double b200mc_ising2d_beta(void* h);
/* host copy of exparr(-8:8) as built by update_exparr (:126-130): out[d + 8] */
int b200mc_ising2d_get_exparr(void* h, double out[17]);
int b200mc_ising2d_set_timing(void* h, int32_t on);
int b200mc_ising2d_get_timing(void* h, int64_t* launches, double* total_ms);
int b200mc_ising2d_sync(void* h);

/* ------------------------------------------------------------------------
 * q-state clock, helical -- type(clock_gpu), src/clock_gpu_m.f90:13-47, and its
 * batched twin src/clock_gpu_multi_m.f90:13-48 (n_multi independent replicas,
 * strict accept test r < w instead of r <= w, :230-235).
 * Observables are REAL64 in the reference (sums of table values, :245-280); here
 * they are computed from exact integer histograms (get_histograms) on the host.
 * Arrays over replicas are replica-major: element (site, j) at j * stride + site.
 * ------------------------------------------------------------------------ */
int b200mc_clock_create(void** h, int64_t nx, int64_t ny, double kbt, int32_t state, int32_t iseed);   /* init, clock_gpu_m :49-79 */
int b200mc_clock_multi_create(void** h, int64_t nx, int64_t ny, double kbt, int32_t state, int32_t n_multi, int32_t iseed); /* clock_gpu_multi_m :50-83 */
int b200mc_clock_destroy(void* h);
int b200mc_clock_set_stream(void* h, void* cuda_stream);
int b200mc_clock_skip_curand(void* h, int64_t n_skip);
int b200mc_clock_set_allup_spin(void* h);            /* :81-84 (all states 0) */
int b200mc_clock_set_random_spin(void* h);           /* :86-104 */
int b200mc_clock_set_kbt(void* h, double kbt);       /* :169-173 */
int b200mc_clock_set_beta(void* h, double beta);     /* :175-181 (+ update_ws :105-146) */
int b200mc_clock_update(void* h);                    /* :183-216 */
int b200mc_clock_update_n(void* h, int32_t n_sweeps);
/* randoms(1:nall[, n_multi]) and next_states(...) in the reference's index order */
int b200mc_clock_update_with_randoms(void* h, const double* randoms, const double* next_states);
int b200mc_clock_calc_energy_sum(void* h, double* res /* n_multi values */); /* :245-262 / multi :265-289 */
int b200mc_clock_calc_magne_sum(void* h, double* res);                       /* :264-280 / multi :291-315 */
/* exact integer observables per replica: hist[j*q + c] = #{s = c}; bond_left[j*q + d] = #{(s(i-1) - s(i)) mod q = d};
 * bond_down: same for (s(i-nx) - s(i)) */
int b200mc_clock_get_histograms(void* h, int64_t* hist, int64_t* bond_left, int64_t* bond_down);
/* spins(), :238-242: int32 0..q-1, layout spins(1-nx : nall+nx[, n_multi]) */
int b200mc_clock_get_spins(void* h, int32_t* out);
int b200mc_clock_set_spins(void* h, const int32_t* in);
/* host copy of ws(0:q-1, ...) (q^6 doubles) exactly as update_ws builds it */
int b200mc_clock_get_ws(void* h, double* out);
int64_t b200mc_clock_nx(void* h);
int64_t b200mc_clock_ny(void* h);
int64_t b200mc_clock_nall(void* h);
int32_t b200mc_clock_state(void* h);
int32_t b200mc_clock_n_multi(void* h);
/* A batch split across GPUs (one handle per rank, no exchange: the replicas of clock_gpu_multi_m :13-48 are
 * independent): this handle's samples are samples first_sample .. first_sample + n_multi - 1 of the job, i.e. they
 * draw the random streams those samples have in a single handle holding the whole batch.  Call before the first
 * set_random_spin / update. */
int b200mc_clock_set_sample_offset(void* h, int32_t first_sample);
double b200mc_clock_kbt(void* h);
double b200mc_clock_beta(void* h);
int b200mc_clock_sync(void* h);

/* ------------------------------------------------------------------------
 * Periodic q-state clock with the full q^6 table -- module clock_tableall_gpu_m,
 * src/clock/clock_tableall_gpu_m.f90:43-45 (module procedures, module-global state), and its
 * compact two-colour twin module clock_dual_lattice_tableall_gpu_m,
 * src/clock/clock_dual_lattice_tableall_m.f90:43-45 (same trajectory: its randoms are indexed
 * by the full-lattice coordinate, :144-152).  One handle serves both: the state is stored as
 * the dual-lattice colour arrays; get/set_sixclock convert to the tableall array.
 * nx, ny, kbt, mstate are compile-time parameters in the reference (:10-15), run-time here;
 * n_multi independent samples ("multi-sample batch") are updated by one launch per colour,
 * arrays over samples are sample-major.  True periodic boundary, colour = (x + y) parity;
 * nx and ny must be even.
 * ------------------------------------------------------------------------ */
int b200mc_sixclock_create(void** h, int64_t nx, int64_t ny, double kbt, int32_t mstate, int32_t n_multi, int32_t iseed); /* init_sixclock :57-88 */
/* variant 0: the delta-E expression of clock_tableall_gpu_m (:72-75) and clock_table_gpu_m (src/clock/clock_table_gpu_m.f90:
 * 119-122, same expression evaluated per site); variant 1: clock_simple_gpu_m's sum over the four neighbours
 * (src/clock/clock_simple_gpu_m.f90:108-113) -- the two differ in the last bits of delta-E, hence in the table */
int b200mc_sixclock_create_variant(void** h, int64_t nx, int64_t ny, double kbt, int32_t mstate, int32_t n_multi, int32_t iseed, int32_t variant);
int b200mc_sixclock_destroy(void* h);
int b200mc_sixclock_set_stream(void* h, void* cuda_stream);
int b200mc_sixclock_skip_curand_clock(void* h, int64_t n_skip);   /* :51-55 */
int b200mc_sixclock_init_sixclock_order(void* h);                 /* :90-92 */
int b200mc_sixclock_set_kbt(void* h, double kbt);                 /* rebuilds states_to_prob (:66-86) */
int b200mc_sixclock_update_metropolis(void* h);                   /* one MCS, :94-152 */
int b200mc_sixclock_update_metropolis_n(void* h, int32_t n_sweeps);
/* one MCS reading rnds(2, nx, ny[, n_multi]) from the host in the reference's order (:95) */
int b200mc_sixclock_update_with_rnds(void* h, const double* rnds);
int b200mc_sixclock_calc_energy(void* h, double* res /* n_multi per-site values */); /* :167-181 */
int b200mc_sixclock_calc_magne(void* h, double* res);                                /* :155-165 */
/* exact integer observables per sample: hist[j*q + c] = #{s = c}; bond_right[j*q + d] =
 * #{(s(x+1, y) - s(x, y)) mod q = d}; bond_up: same for (x, y+1) */
int b200mc_sixclock_get_histograms(void* h, int64_t* hist, int64_t* bond_right, int64_t* bond_up);
/* sixclock(nx, ny[, n_multi]) int32, Fortran order (tableall :22) */
int b200mc_sixclock_get_sixclock(void* h, int32_t* out);
int b200mc_sixclock_set_sixclock(void* h, const int32_t* in);
/* sixclock_even / sixclock_odd (nx/2, ny[, n_multi]) int32 (dual lattice :22) */
int b200mc_sixclock_get_dual(void* h, int32_t* even, int32_t* odd);
int b200mc_sixclock_set_dual(void* h, const int32_t* even, const int32_t* odd);
/* host copies of states_to_prob(c, new_c, r, u, l, d) (q^6) and state_center_right_up_to_energy (q^3) */
int b200mc_sixclock_get_states_to_prob(void* h, double* out);
int b200mc_sixclock_get_energy_table(void* h, double* out);
int64_t b200mc_sixclock_nx(void* h);
int64_t b200mc_sixclock_ny(void* h);
int64_t b200mc_sixclock_nall(void* h);
int32_t b200mc_sixclock_mstate(void* h);
int32_t b200mc_sixclock_n_multi(void* h);
/* as b200mc_clock_set_sample_offset, for the batched periodic clock modules */
int b200mc_sixclock_set_sample_offset(void* h, int32_t first_sample);
double b200mc_sixclock_kbt(void* h);
double b200mc_sixclock_beta(void* h);
int b200mc_sixclock_set_timing(void* h, int32_t on);
int b200mc_sixclock_get_timing(void* h, int64_t* launches, double* total_ms);
int b200mc_sixclock_sync(void* h);

/* ------------------------------------------------------------------------
 * XY 2D, periodic -- type(xy2d_gpu), src/xy2d_periodic_gpu_m.f90:14-59.
 * State: one fp32 angle per site (turns); energy / magnetisation are real64
 * sums, compared with the reference at 1e-5 relative (BASELINE north_star).
 * nx >= 8 even and ny even, like the reference (:377-380); rows are padded to whole float4 groups internally.
 * ------------------------------------------------------------------------ */
int b200mc_xy2d_create(void** h, int64_t nx, int64_t ny, double kbt, int32_t iseed);  /* init, :61-78 */
/* Slab of a lattice of ny_global rows split along y over nranks GPUs (one process per GPU): this rank holds
 * ny_global / nranks rows (an even number).  After every colour pass the first / last row of the colour just written
 * goes to the neighbouring ranks' halo rows (ncclSend/Recv); E, Mx, My are all-reduced; every site draws the random
 * numbers it draws in the one-GPU run.  In slab mode ny() is the LOCAL row count, nall() the whole lattice, get_angles /
 * get_spins return the local rows, and the correlation sums are unsupported.
 * Validated on 2 GPUs against the one-GPU run (angles bit-identical; profiles/r02_slab_parity_2gpu.log).
 * nranks == 1 is b200mc_xy2d_create. */
int b200mc_xy2d_create_slab(void** h, int64_t nx, int64_t ny_global, double kbt, int32_t iseed, int32_t rank, int32_t nranks,
                            const char* nccl_id /* 128 bytes from b200mc_dist_unique_id; NULL when nranks == 1 */);
int b200mc_xy2d_destroy(void* h);
int b200mc_xy2d_set_stream(void* h, void* cuda_stream);
int b200mc_xy2d_skip_curand(void* h, int64_t n_skip);                 /* :79-84 */
int b200mc_xy2d_set_allup_spin(void* h);                              /* :86-101 */
int b200mc_xy2d_set_random_spin(void* h);                             /* :105-122 */
int b200mc_xy2d_set_kbt(void* h, double kbt);                         /* :329-333 */
int b200mc_xy2d_set_beta(void* h, double beta);                       /* :335-339 */
int b200mc_xy2d_update(void* h);                                      /* Metropolis MCS, :353-397 */
int b200mc_xy2d_update_n(void* h, int32_t n_sweeps);
/* one Metropolis MCS reading the accept uniforms randoms(nx, ny) and the candidate uniforms candidates(nx, ny)
 * (the two arrays the reference fills with curandGenerate at :355-356 and reads at :382-384) from host arrays */
int b200mc_xy2d_update_with_randoms(void* h, const double* randoms, const double* candidates);
int b200mc_xy2d_update_over_relaxation(void* h, int32_t n_steps);     /* :400-439 */
int b200mc_xy2d_calc_energy_sum(void* h, double* e);                  /* :469-472,496-508 */
int b200mc_xy2d_calc_magne_sum(void* h, double* mx);                  /* :474-477,510-521 */
int b200mc_xy2d_calc_magne_y_sum(void* h, double* my);                /* :479-482,523-534 */
int b200mc_xy2d_measure(void* h, double* e, double* mx, double* my);  /* all three, one pass */
int b200mc_xy2d_set_initial_magne_autocorrelation_state(void* h);     /* :342-350 */
int b200mc_xy2d_calc_autocorrelation_sum(void* h, double* r);         /* :484-487,536-549 */
int b200mc_xy2d_calc_correlation_sum(void* h, double* r);             /* :489-492,551-567 */
/* rotate_summation_magne_toward_xaxis (:219-232); with_autocorrelation != 0 also rotates the
 * stored snapshot (rotate_summation_magne_and_autocorrelation_toward_xaxis, :235-250) */
int b200mc_xy2d_rotate_summation_magne_toward_xaxis(void* h, int32_t with_autocorrelation);
/* one application of metropolis_by_field_sub (:198-216) to every site: candidate accepted iff
 * r <= 1 - exp(dE), dE = -(h . (cand - s)); draws one (randoms, candidates) pair of arrays */
int b200mc_xy2d_metropolis_by_field(void* h, double hx, double hy);
/* initial-state preparation (random start, field sweeps until |m| meets the criterion, rotate M onto
 * the x axis): set_finite_magne_spin :126-154, set_random_small_spin :158-175, set_random_near_spin
 * :179-196.  B200MC_ERR_STATE if the criterion is not met after 4096 field sweeps: the reference
 * loops for ever then, and its set_finite_magne_spin heuristic (field doubled / halved-and-reversed)
 * does cycle for most targets. */
int b200mc_xy2d_set_finite_magne_spin(void* h, double init_magne);
int b200mc_xy2d_set_random_small_spin(void* h, double near_magne);
int b200mc_xy2d_set_random_near_spin(void* h, double near_magne, double diff_parcent);
/* spins(), :461-465: real64 (cos, sin), layout spins(0:nx+1, 0:ny+1, 1:2) -> 2 (nx+2)(ny+2) doubles;
 * the halo frame is returned refreshed, corners 0 */
int b200mc_xy2d_get_spins(void* h, double* out);
/* native state: angles in turns, fp32, [ny][nx] row-major (exact round trip; used by the parity tests) */
int b200mc_xy2d_get_angles(void* h, float* out);
int b200mc_xy2d_set_angles(void* h, const float* in);
int64_t b200mc_xy2d_nx(void* h);
int64_t b200mc_xy2d_ny(void* h);
int64_t b200mc_xy2d_nall(void* h);
double b200mc_xy2d_kbt(void* h);
double b200mc_xy2d_beta(void* h);
int b200mc_xy2d_sync(void* h);

/* ------------------------------------------------------------------------
 * XY 2D, helical boundary -- type(xy2d_gpu) of module xy2d_gpu_m, src/xy2d_gpu_m.f90:12-43
 * (SURVEY 8 f3).  Linear-index colouring like ising2d_gpu_m: nx odd, ny even REQUIRED.
 * State: one fp32 angle (turns) per site; E and Mx are real64 sums (1e-5 relative vs the reference).
 * ------------------------------------------------------------------------ */
int b200mc_xy2dh_create(void** h, int64_t nx, int64_t ny, double kbt, int32_t iseed);  /* init, :45-62 */
int b200mc_xy2dh_destroy(void* h);
int b200mc_xy2dh_set_stream(void* h, void* cuda_stream);
int b200mc_xy2dh_skip_curand(void* h, int64_t n_skip);                /* :63-68 */
int b200mc_xy2dh_set_allup_spin(void* h);                             /* :70-88 */
int b200mc_xy2dh_set_random_spin(void* h);                            /* :90-105 */
int b200mc_xy2dh_set_kbt(void* h, double kbt);                        /* :127-131 */
int b200mc_xy2dh_set_beta(void* h, double beta);                      /* :133-137 */
int b200mc_xy2dh_update(void* h);                                     /* :138-174 */
int b200mc_xy2dh_update_n(void* h, int32_t n_sweeps);
int b200mc_xy2dh_update_over_relaxation(void* h, int32_t n_steps);    /* :176-213 */
int b200mc_xy2dh_calc_energy_sum(void* h, double* e);                 /* :259-276 */
int b200mc_xy2dh_calc_magne_sum(void* h, double* mx);                 /* :278-291 */
/* spins(), :236-240: real64 (cos, sin), layout spins(1-nx : nall+nx, 1:2) -> 2 (nall + 2 nx) doubles, halo refreshed */
int b200mc_xy2dh_get_spins(void* h, double* out);
/* native state: angles in turns, fp32, linear index order [nall] */
int b200mc_xy2dh_get_angles(void* h, float* out);
int b200mc_xy2dh_set_angles(void* h, const float* in);
int64_t b200mc_xy2dh_nx(void* h);
int64_t b200mc_xy2dh_ny(void* h);
int64_t b200mc_xy2dh_nall(void* h);
double b200mc_xy2dh_kbt(void* h);
double b200mc_xy2dh_beta(void* h);
int b200mc_xy2dh_sync(void* h);

#pragma GCC visibility pop
#ifdef __cplusplus
}
#endif
#endif /* B200MC_H */
