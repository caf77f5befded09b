/* A plain-C consumer of include/b200mc.h: the drivers' inner loop
 *     do i = 1, mcs:  call update();  m = calc_magne_sum();  e = calc_energy_sum()
 * (app/ising3d_gpu_relaxation.f90:38-48, app/ising2d_gpu_relaxation.f90:36-45 of the reference) through the C ABI
 * alone -- no Python, no C++, no CUDA headers.  Built by tests/test_c_consumer.py with
 *     gcc -std=c11 -Wall -Wextra -pedantic -Werror
 * Usage: relaxation_loop 2|3|-2|-3 nx ny nz kbt iseed mcs allup|random   -> one line "i e m" per MCS on stdout
 * (a negative dimension: the same loop on the torus, b200mc_ising_torus_*).
 * Exit codes: 0 ok, 2 create failed (message on stderr: e.g. no CUDA device), 3 any other API error. */
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200mc.h"

#define CHECK(call)                                                              \
    do {                                                                         \
        const int rc_ = (call);                                                  \
        if (rc_ != B200MC_OK) {                                                  \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, b200mc_last_error()); \
            return 3;                                                            \
        }                                                                        \
    } while (0)

int main(int argc, char** argv)
{
    if (argc != 9) {
        fprintf(stderr, "usage: %s 2|3|-2|-3 nx ny nz kbt iseed mcs allup|random\n", argv[0]);
        return 1;
    }
    const int dim = atoi(argv[1]);
    const int64_t nx = atoll(argv[2]), ny = atoll(argv[3]), nz = atoll(argv[4]);
    const double kbt = atof(argv[5]);
    const int32_t iseed = (int32_t)atoi(argv[6]);
    const int mcs = atoi(argv[7]);
    const int random_start = strcmp(argv[8], "random") == 0;
    void* h = NULL;
    if (dim < 0) {   /* true periodic boundaries: L = 1024^3 and other even shapes the reference's helical types cannot run */
        const int rct = b200mc_ising_torus_create(&h, -dim, nx, ny, nz, kbt, iseed);
        if (rct != B200MC_OK) {
            fprintf(stderr, "create -> %d: %s\n", rct, b200mc_last_error());
            return 2;
        }
        if (random_start) CHECK(b200mc_ising_torus_set_random_spin(h));
        else CHECK(b200mc_ising_torus_set_allup_spin(h));
        for (int i = 1; i <= mcs; ++i) {
            int64_t e = 0, m = 0;
            CHECK(b200mc_ising_torus_update(h));
            CHECK(b200mc_ising_torus_calc_magne_sum(h, &m));
            CHECK(b200mc_ising_torus_calc_energy_sum(h, &e));
            printf("%d %" PRId64 " %" PRId64 "\n", i, e, m);
        }
        CHECK(b200mc_ising_torus_destroy(h));
        printf("launches %llu version %d\n", b200mc_launch_count(), b200mc_version());
        return 0;
    }
    const int rc = dim == 3 ? b200mc_ising3d_create(&h, nx, ny, nz, kbt, iseed) : b200mc_ising2d_create(&h, nx, ny, kbt, iseed);
    if (rc != B200MC_OK) {
        fprintf(stderr, "create -> %d: %s\n", rc, b200mc_last_error());
        return 2;
    }
    if (dim == 3) {
        if (random_start) CHECK(b200mc_ising3d_set_random_spin(h));
        else CHECK(b200mc_ising3d_set_allup_spin(h));
        for (int i = 1; i <= mcs; ++i) {
            int64_t e = 0, m = 0;
            CHECK(b200mc_ising3d_update(h));
            CHECK(b200mc_ising3d_calc_magne_sum(h, &m));
            CHECK(b200mc_ising3d_calc_energy_sum(h, &e));
            printf("%d %" PRId64 " %" PRId64 "\n", i, e, m);
        }
        CHECK(b200mc_ising3d_destroy(h));
    } else {
        if (random_start) CHECK(b200mc_ising2d_set_random_spin(h));
        else CHECK(b200mc_ising2d_set_allup_spin(h));
        for (int i = 1; i <= mcs; ++i) {
            int64_t e = 0, m = 0;
            CHECK(b200mc_ising2d_update(h));
            CHECK(b200mc_ising2d_calc_magne_sum(h, &m));
            CHECK(b200mc_ising2d_calc_energy_sum(h, &e));
            printf("%d %" PRId64 " %" PRId64 "\n", i, e, m);
        }
        CHECK(b200mc_ising2d_destroy(h));
    }
    printf("launches %llu version %d\n", b200mc_launch_count(), b200mc_version());
    return 0;
}
