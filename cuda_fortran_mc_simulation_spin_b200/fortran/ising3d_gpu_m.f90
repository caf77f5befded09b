!> Drop-in replacement for the reference's `module ising3d_gpu_m`
!> (src/ising3d_gpu_m.f90): same module name, same `type(ising3d_gpu)`, same
!> type-bound procedure names and argument kinds, so app/ising3d_gpu_relaxation.f90
!> compiles unchanged.  All work is done by libb200mc.so (include/b200mc.h)
!> through ISO_C_BINDING; this file contains no CUDA Fortran and builds with any
!> Fortran 2008 compiler.  NOT COMPILED IN THE BUILD IMAGE (no Fortran compiler there).
module ising3d_gpu_m
  use, intrinsic :: iso_fortran_env
  use, intrinsic :: iso_c_binding
  implicit none
  private
  !> same public status variable as the reference (src/ising3d_gpu_m.f90:8); holds the last C return code
  integer(int32), public, protected :: ising3d_gpu_stat = 0
  public :: ising3d_gpu
  type :: ising3d_gpu
     private
     type(c_ptr) :: h_ = c_null_ptr
   contains
     procedure, pass :: init => init_ising3d_gpu
     procedure, pass :: skip_curand => skip_curand_ising3d_gpu
     procedure, pass :: set_allup_spin => set_allup_spin_ising3d_gpu
     procedure, pass :: set_random_spin => set_random_spin_ising3d_gpu
     procedure, pass :: set_kbt => set_kbt_ising3d_gpu
     procedure, pass :: set_beta => set_beta_ising3d_gpu
     procedure, pass :: update => update_ising3d_gpu
     procedure, pass :: nx => nx_ising3d_gpu
     procedure, pass :: ny => ny_ising3d_gpu
     procedure, pass :: nz => nz_ising3d_gpu
     procedure, pass :: nall => nall_ising3d_gpu
     procedure, pass :: kbt => kbt_ising3d_gpu
     procedure, pass :: beta => beta_ising3d_gpu
     procedure, pass :: spins => spins_ising3d_gpu
     procedure, pass :: calc_energy_sum => calc_energy_sum_ising3d_gpu
     procedure, pass :: calc_magne_sum => calc_magne_sum_ising3d_gpu
     !> additions (not in the reference type)
     procedure, pass :: set_method => set_method_ising3d_gpu   !< 0 Metropolis (default), 1 heat-bath
     procedure, pass :: update_n => update_n_ising3d_gpu       !< n MCS back to back
     !> mcs x [update; calc_magne_sum; calc_energy_sum] on the device
     procedure, pass :: run_relaxation => run_relaxation_ising3d_gpu
     !> the drivers' whole measurement on the device: tot_sample x [initial state; mcs x (update; m; e; add_data)],
     !> Kahan mean / variance / covariance per MCS (app/ising3d_gpu_relaxation.f90), see include/b200mc.h
     procedure, pass :: run_relaxation_stats => run_relaxation_stats_ising3d_gpu
     final :: destroy_ising3d_gpu
  end type ising3d_gpu

  interface
     integer(c_int) function b200mc_ising3d_create(h, nx, ny, nz, kbt, iseed) bind(C, name="b200mc_ising3d_create")
       import; type(c_ptr), intent(out) :: h
       integer(c_int64_t), value :: nx, ny, nz; real(c_double), value :: kbt; integer(c_int32_t), value :: iseed
     end function
     integer(c_int) function b200mc_ising3d_destroy(h) bind(C, name="b200mc_ising3d_destroy")
       import; type(c_ptr), value :: h
     end function
     integer(c_int) function b200mc_ising3d_skip_curand(h, n) bind(C, name="b200mc_ising3d_skip_curand")
       import; type(c_ptr), value :: h; integer(c_int64_t), value :: n
     end function
     integer(c_int) function b200mc_ising3d_set_allup_spin(h) bind(C, name="b200mc_ising3d_set_allup_spin")
       import; type(c_ptr), value :: h
     end function
     integer(c_int) function b200mc_ising3d_set_random_spin(h) bind(C, name="b200mc_ising3d_set_random_spin")
       import; type(c_ptr), value :: h
     end function
     integer(c_int) function b200mc_ising3d_set_kbt(h, kbt) bind(C, name="b200mc_ising3d_set_kbt")
       import; type(c_ptr), value :: h; real(c_double), value :: kbt
     end function
     integer(c_int) function b200mc_ising3d_set_beta(h, beta) bind(C, name="b200mc_ising3d_set_beta")
       import; type(c_ptr), value :: h; real(c_double), value :: beta
     end function
     integer(c_int) function b200mc_ising3d_set_method(h, method) bind(C, name="b200mc_ising3d_set_method")
       import; type(c_ptr), value :: h; integer(c_int32_t), value :: method
     end function
     integer(c_int) function b200mc_ising3d_update(h) bind(C, name="b200mc_ising3d_update")
       import; type(c_ptr), value :: h
     end function
     integer(c_int) function b200mc_ising3d_update_n(h, n) bind(C, name="b200mc_ising3d_update_n")
       import; type(c_ptr), value :: h; integer(c_int32_t), value :: n
     end function
     integer(c_int) function b200mc_ising3d_run_relaxation(h, mcs, e, m) bind(C, name="b200mc_ising3d_run_relaxation")
       import; type(c_ptr), value :: h; integer(c_int32_t), value :: mcs; integer(c_int64_t), intent(out) :: e(*), m(*)
     end function
     integer(c_int) function b200mc_ising3d_run_relaxation_stats(h, mcs, tot_sample, random_start, res) &
         bind(C, name="b200mc_ising3d_run_relaxation_stats")
       import; type(c_ptr), value :: h; integer(c_int32_t), value :: mcs, tot_sample, random_start
       real(c_double), intent(out) :: res(*)
     end function
     integer(c_int) function b200mc_ising3d_calc_energy_sum(h, e) bind(C, name="b200mc_ising3d_calc_energy_sum")
       import; type(c_ptr), value :: h; integer(c_int64_t), intent(out) :: e
     end function
     integer(c_int) function b200mc_ising3d_calc_magne_sum(h, m) bind(C, name="b200mc_ising3d_calc_magne_sum")
       import; type(c_ptr), value :: h; integer(c_int64_t), intent(out) :: m
     end function
     integer(c_int) function b200mc_ising3d_get_spins(h, out) bind(C, name="b200mc_ising3d_get_spins")
       import; type(c_ptr), value :: h; integer(c_int32_t), intent(out) :: out(*)
     end function
     integer(c_int64_t) function b200mc_ising3d_nx(h) bind(C, name="b200mc_ising3d_nx")
       import; type(c_ptr), value :: h
     end function
     integer(c_int64_t) function b200mc_ising3d_ny(h) bind(C, name="b200mc_ising3d_ny")
       import; type(c_ptr), value :: h
     end function
     integer(c_int64_t) function b200mc_ising3d_nz(h) bind(C, name="b200mc_ising3d_nz")
       import; type(c_ptr), value :: h
     end function
     integer(c_int64_t) function b200mc_ising3d_nall(h) bind(C, name="b200mc_ising3d_nall")
       import; type(c_ptr), value :: h
     end function
     real(c_double) function b200mc_ising3d_kbt(h) bind(C, name="b200mc_ising3d_kbt")
       import; type(c_ptr), value :: h
     end function
     real(c_double) function b200mc_ising3d_beta(h) bind(C, name="b200mc_ising3d_beta")
       import; type(c_ptr), value :: h
     end function
     subroutine b200mc_print_last_error() bind(C, name="b200mc_print_last_error")
     end subroutine
  end interface
contains
  !> src/ising3d_gpu_m.f90:50-71
  impure subroutine init_ising3d_gpu(this, nx, ny, nz, kbt, iseed)
    class(ising3d_gpu), intent(inout) :: this
    integer(int64), intent(in) :: nx, ny, nz
    real(real64), intent(in) :: kbt
    integer(int32), intent(in) :: iseed
    if (c_associated(this%h_)) ising3d_gpu_stat = b200mc_ising3d_destroy(this%h_)
    ising3d_gpu_stat = b200mc_ising3d_create(this%h_, nx, ny, nz, kbt, iseed)
    if (ising3d_gpu_stat /= 0) then
       call b200mc_print_last_error()   ! the library's own message (shape, device, ...) on stderr
       error stop "ising3d_gpu%init: b200mc_ising3d_create failed (shape must have nx and ny odd," // &
          " nz even -- the reference races on other shapes, e.g. its own 1001 x 1000 x 1000 default; or no CUDA device)"
    end if
  end subroutine init_ising3d_gpu
  impure subroutine destroy_ising3d_gpu(this)
    type(ising3d_gpu), intent(inout) :: this
    if (c_associated(this%h_)) ising3d_gpu_stat = b200mc_ising3d_destroy(this%h_)
    this%h_ = c_null_ptr
  end subroutine destroy_ising3d_gpu
  !> :72-77
  impure subroutine skip_curand_ising3d_gpu(this, n_skip)
    class(ising3d_gpu), intent(inout) :: this
    integer(int64), intent(in) :: n_skip
    ising3d_gpu_stat = b200mc_ising3d_skip_curand(this%h_, n_skip)
  end subroutine skip_curand_ising3d_gpu
  !> :79-82
  impure subroutine set_allup_spin_ising3d_gpu(this)
    class(ising3d_gpu), intent(inout) :: this
    ising3d_gpu_stat = b200mc_ising3d_set_allup_spin(this%h_)
  end subroutine set_allup_spin_ising3d_gpu
  !> :84-90
  impure subroutine set_random_spin_ising3d_gpu(this)
    class(ising3d_gpu), intent(inout) :: this
    ising3d_gpu_stat = b200mc_ising3d_set_random_spin(this%h_)
  end subroutine set_random_spin_ising3d_gpu
  !> :124-128
  impure subroutine set_kbt_ising3d_gpu(this, kbt)
    class(ising3d_gpu), intent(inout) :: this
    real(real64), intent(in) :: kbt
    ising3d_gpu_stat = b200mc_ising3d_set_kbt(this%h_, kbt)
  end subroutine set_kbt_ising3d_gpu
  !> :130-135
  impure subroutine set_beta_ising3d_gpu(this, beta)
    class(ising3d_gpu), intent(inout) :: this
    real(real64), intent(in) :: beta
    ising3d_gpu_stat = b200mc_ising3d_set_beta(this%h_, beta)
  end subroutine set_beta_ising3d_gpu
  impure subroutine set_method_ising3d_gpu(this, method)
    class(ising3d_gpu), intent(inout) :: this
    integer(int32), intent(in) :: method
    ising3d_gpu_stat = b200mc_ising3d_set_method(this%h_, method)
  end subroutine set_method_ising3d_gpu
  !> :174-188
  impure subroutine update_ising3d_gpu(this)
    class(ising3d_gpu), intent(inout) :: this
    ising3d_gpu_stat = b200mc_ising3d_update(this%h_)
  end subroutine update_ising3d_gpu
  !> the drivers' inner loop (app/ising3d_gpu_relaxation.f90) without a host round trip per MCS:
  !> e(i), m(i) = calc_energy_sum(), calc_magne_sum() after MCS i
  impure subroutine run_relaxation_ising3d_gpu(this, mcs, e, m)
    class(ising3d_gpu), intent(inout) :: this
    integer(int32), intent(in) :: mcs
    integer(int64), intent(out) :: e(mcs), m(mcs)
    ising3d_gpu_stat = b200mc_ising3d_run_relaxation(this%h_, mcs, e, m)
  end subroutine run_relaxation_ising3d_gpu
  !> res(1:8, i) = num_sample, mean1 (m), mean2 (e), square_mean1, square_mean2, var1, var2, cov after MCS i
  impure subroutine run_relaxation_stats_ising3d_gpu(this, mcs, tot_sample, random_start, res)
    class(ising3d_gpu), intent(inout) :: this
    integer(int32), intent(in) :: mcs, tot_sample
    logical, intent(in) :: random_start
    real(real64), intent(out) :: res(8, mcs)
    ising3d_gpu_stat = b200mc_ising3d_run_relaxation_stats(this%h_, mcs, tot_sample, merge(1_int32, 0_int32, random_start), res)
  end subroutine run_relaxation_stats_ising3d_gpu
  impure subroutine update_n_ising3d_gpu(this, n_sweeps)
    class(ising3d_gpu), intent(inout) :: this
    integer(int32), intent(in) :: n_sweeps
    ising3d_gpu_stat = b200mc_ising3d_update_n(this%h_, n_sweeps)
  end subroutine update_n_ising3d_gpu
  !> :208-231
  impure integer(int64) function nx_ising3d_gpu(this) result(res)
    class(ising3d_gpu), intent(in) :: this
    res = b200mc_ising3d_nx(this%h_)
  end function nx_ising3d_gpu
  impure integer(int64) function ny_ising3d_gpu(this) result(res)
    class(ising3d_gpu), intent(in) :: this
    res = b200mc_ising3d_ny(this%h_)
  end function ny_ising3d_gpu
  impure integer(int64) function nz_ising3d_gpu(this) result(res)
    class(ising3d_gpu), intent(in) :: this
    res = b200mc_ising3d_nz(this%h_)
  end function nz_ising3d_gpu
  impure integer(int64) function nall_ising3d_gpu(this) result(res)
    class(ising3d_gpu), intent(in) :: this
    res = b200mc_ising3d_nall(this%h_)
  end function nall_ising3d_gpu
  impure real(real64) function kbt_ising3d_gpu(this) result(res)
    class(ising3d_gpu), intent(in) :: this
    res = b200mc_ising3d_kbt(this%h_)
  end function kbt_ising3d_gpu
  impure real(real64) function beta_ising3d_gpu(this) result(res)
    class(ising3d_gpu), intent(in) :: this
    res = b200mc_ising3d_beta(this%h_)
  end function beta_ising3d_gpu
  !> :232-236 -- raw array, halo cells included, same bounds as the reference
  impure function spins_ising3d_gpu(this) result(res)
    class(ising3d_gpu), intent(in) :: this
    integer(int32), allocatable :: res(:)
    integer(int64) :: nxy, nall
    nxy = b200mc_ising3d_nx(this%h_) * b200mc_ising3d_ny(this%h_)
    nall = b200mc_ising3d_nall(this%h_)
    allocate(res(1 - nxy : nall + nxy))
    ising3d_gpu_stat = b200mc_ising3d_get_spins(this%h_, res)
  end function spins_ising3d_gpu
  !> :239-257
  impure integer(int64) function calc_energy_sum_ising3d_gpu(this) result(res)
    class(ising3d_gpu), intent(in) :: this
    ising3d_gpu_stat = b200mc_ising3d_calc_energy_sum(this%h_, res)
  end function calc_energy_sum_ising3d_gpu
  !> :259-276
  impure integer(int64) function calc_magne_sum_ising3d_gpu(this) result(res)
    class(ising3d_gpu), intent(in) :: this
    ising3d_gpu_stat = b200mc_ising3d_calc_magne_sum(this%h_, res)
  end function calc_magne_sum_ising3d_gpu
end module ising3d_gpu_m
