!> Ising 2D with TRUE PERIODIC boundaries (torus): `module ising2d_periodic_gpu_m`, exporting a `type(ising2d_gpu)` with
!> the type-bound procedure names and argument kinds of the reference's src/ising2d_gpu_m.f90 -- a driver such as
!> app/ising2d_gpu_relaxation.f90 switches by changing its `use ising2d_gpu_m` line and can then run 1024^2 or 65536^2
!> (BASELINE.md C1 / C5; the reference's helical type needs odd nx; its kernels race on other shapes).
!> NOT A REFERENCE MODULE (include/b200mc.h, b200mc_ising_torus_*): nx % 32 == 0, ny even; `spins()` returns
!> nx*ny values s(x + nx y), -1 / +1, WITHOUT halo cells.  All work is done by libb200mc.so through
!> ISO_C_BINDING; no CUDA Fortran.  NOT COMPILED IN THE BUILD IMAGE (no Fortran compiler there).
module ising2d_periodic_gpu_m
  use, intrinsic :: iso_fortran_env
  use, intrinsic :: iso_c_binding
  implicit none
  private
  integer(int32), public, protected :: ising2d_gpu_stat = 0
  public :: ising2d_gpu
  type :: ising2d_gpu
     private
     type(c_ptr) :: h_ = c_null_ptr
   contains
     procedure, pass :: init => init_ising2d_gpu
     procedure, pass :: skip_curand => skip_curand_ising2d_gpu
     procedure, pass :: set_allup_spin => set_allup_spin_ising2d_gpu
     procedure, pass :: set_random_spin => set_random_spin_ising2d_gpu
     procedure, pass :: set_kbt => set_kbt_ising2d_gpu
     procedure, pass :: set_beta => set_beta_ising2d_gpu
     procedure, pass :: update => update_ising2d_gpu
     procedure, pass :: nx => nx_ising2d_gpu
     procedure, pass :: ny => ny_ising2d_gpu
     procedure, pass :: nall => nall_ising2d_gpu
     procedure, pass :: kbt => kbt_ising2d_gpu
     procedure, pass :: beta => beta_ising2d_gpu
     procedure, pass :: spins => spins_ising2d_gpu
     procedure, pass :: calc_energy_sum => calc_energy_sum_ising2d_gpu
     procedure, pass :: calc_magne_sum => calc_magne_sum_ising2d_gpu
     !> additions (not in the reference type)
     procedure, pass :: set_method => set_method_ising2d_gpu   !< 0 Metropolis (default), 1 heat-bath
     procedure, pass :: update_n => update_n_ising2d_gpu       !< n MCS back to back
     final :: destroy_ising2d_gpu
  end type ising2d_gpu

  interface
     integer(c_int) function b200mc_ising_torus_create(h, ndim, nx, ny, nz, kbt, iseed) bind(C, name="b200mc_ising_torus_create")
       import; type(c_ptr), intent(out) :: h; integer(c_int32_t), value :: ndim
       integer(c_int64_t), value :: nx, ny, nz; real(c_double), value :: kbt; integer(c_int32_t), value :: iseed
     end function
     integer(c_int) function b200mc_ising_torus_destroy(h) bind(C, name="b200mc_ising_torus_destroy")
       import; type(c_ptr), value :: h
     end function
     integer(c_int) function b200mc_ising_torus_skip_curand(h, n) bind(C, name="b200mc_ising_torus_skip_curand")
       import; type(c_ptr), value :: h; integer(c_int64_t), value :: n
     end function
     integer(c_int) function b200mc_ising_torus_set_allup_spin(h) bind(C, name="b200mc_ising_torus_set_allup_spin")
       import; type(c_ptr), value :: h
     end function
     integer(c_int) function b200mc_ising_torus_set_random_spin(h) bind(C, name="b200mc_ising_torus_set_random_spin")
       import; type(c_ptr), value :: h
     end function
     integer(c_int) function b200mc_ising_torus_set_kbt(h, kbt) bind(C, name="b200mc_ising_torus_set_kbt")
       import; type(c_ptr), value :: h; real(c_double), value :: kbt
     end function
     integer(c_int) function b200mc_ising_torus_set_beta(h, beta) bind(C, name="b200mc_ising_torus_set_beta")
       import; type(c_ptr), value :: h; real(c_double), value :: beta
     end function
     integer(c_int) function b200mc_ising_torus_set_method(h, method) bind(C, name="b200mc_ising_torus_set_method")
       import; type(c_ptr), value :: h; integer(c_int32_t), value :: method
     end function
     integer(c_int) function b200mc_ising_torus_update(h) bind(C, name="b200mc_ising_torus_update")
       import; type(c_ptr), value :: h
     end function
     integer(c_int) function b200mc_ising_torus_update_n(h, n) bind(C, name="b200mc_ising_torus_update_n")
       import; type(c_ptr), value :: h; integer(c_int32_t), value :: n
     end function
     integer(c_int) function b200mc_ising_torus_calc_energy_sum(h, e) bind(C, name="b200mc_ising_torus_calc_energy_sum")
       import; type(c_ptr), value :: h; integer(c_int64_t), intent(out) :: e
     end function
     integer(c_int) function b200mc_ising_torus_calc_magne_sum(h, m) bind(C, name="b200mc_ising_torus_calc_magne_sum")
       import; type(c_ptr), value :: h; integer(c_int64_t), intent(out) :: m
     end function
     integer(c_int) function b200mc_ising_torus_get_spins(h, out) bind(C, name="b200mc_ising_torus_get_spins")
       import; type(c_ptr), value :: h; integer(c_int32_t), intent(out) :: out(*)
     end function
     integer(c_int64_t) function b200mc_ising_torus_nx(h) bind(C, name="b200mc_ising_torus_nx")
       import; type(c_ptr), value :: h
     end function
     integer(c_int64_t) function b200mc_ising_torus_ny(h) bind(C, name="b200mc_ising_torus_ny")
       import; type(c_ptr), value :: h
     end function
     integer(c_int64_t) function b200mc_ising_torus_nall(h) bind(C, name="b200mc_ising_torus_nall")
       import; type(c_ptr), value :: h
     end function
     real(c_double) function b200mc_ising_torus_kbt(h) bind(C, name="b200mc_ising_torus_kbt")
       import; type(c_ptr), value :: h
     end function
     real(c_double) function b200mc_ising_torus_beta(h) bind(C, name="b200mc_ising_torus_beta")
       import; type(c_ptr), value :: h
     end function
     subroutine b200mc_print_last_error() bind(C, name="b200mc_print_last_error")
     end subroutine
  end interface
contains
  !> src/ising2d_gpu_m.f90:50-71
  impure subroutine init_ising2d_gpu(this, nx, ny, kbt, iseed)
    class(ising2d_gpu), intent(inout) :: this
    integer(int64), intent(in) :: nx, ny
    real(real64), intent(in) :: kbt
    integer(int32), intent(in) :: iseed
    if (c_associated(this%h_)) ising2d_gpu_stat = b200mc_ising_torus_destroy(this%h_)
    ising2d_gpu_stat = b200mc_ising_torus_create(this%h_, 2_int32, nx, ny, 0_int64, kbt, iseed)
    if (ising2d_gpu_stat /= 0) then
       call b200mc_print_last_error()   ! the library's own message (shape, device, ...) on stderr
       error stop "ising2d_gpu%init (periodic): b200mc_ising_torus_create failed" // &
          " (nx must be a multiple of 32, ny even; or no CUDA device)"
    end if
  end subroutine init_ising2d_gpu
  impure subroutine destroy_ising2d_gpu(this)
    type(ising2d_gpu), intent(inout) :: this
    if (c_associated(this%h_)) ising2d_gpu_stat = b200mc_ising_torus_destroy(this%h_)
    this%h_ = c_null_ptr
  end subroutine destroy_ising2d_gpu
  !> :72-77
  impure subroutine skip_curand_ising2d_gpu(this, n_skip)
    class(ising2d_gpu), intent(inout) :: this
    integer(int64), intent(in) :: n_skip
    ising2d_gpu_stat = b200mc_ising_torus_skip_curand(this%h_, n_skip)
  end subroutine skip_curand_ising2d_gpu
  !> :79-82
  impure subroutine set_allup_spin_ising2d_gpu(this)
    class(ising2d_gpu), intent(inout) :: this
    ising2d_gpu_stat = b200mc_ising_torus_set_allup_spin(this%h_)
  end subroutine set_allup_spin_ising2d_gpu
  !> :84-90
  impure subroutine set_random_spin_ising2d_gpu(this)
    class(ising2d_gpu), intent(inout) :: this
    ising2d_gpu_stat = b200mc_ising_torus_set_random_spin(this%h_)
  end subroutine set_random_spin_ising2d_gpu
  !> :124-128
  impure subroutine set_kbt_ising2d_gpu(this, kbt)
    class(ising2d_gpu), intent(inout) :: this
    real(real64), intent(in) :: kbt
    ising2d_gpu_stat = b200mc_ising_torus_set_kbt(this%h_, kbt)
  end subroutine set_kbt_ising2d_gpu
  !> :130-135
  impure subroutine set_beta_ising2d_gpu(this, beta)
    class(ising2d_gpu), intent(inout) :: this
    real(real64), intent(in) :: beta
    ising2d_gpu_stat = b200mc_ising_torus_set_beta(this%h_, beta)
  end subroutine set_beta_ising2d_gpu
  impure subroutine set_method_ising2d_gpu(this, method)
    class(ising2d_gpu), intent(inout) :: this
    integer(int32), intent(in) :: method
    ising2d_gpu_stat = b200mc_ising_torus_set_method(this%h_, method)
  end subroutine set_method_ising2d_gpu
  !> :174-188
  impure subroutine update_ising2d_gpu(this)
    class(ising2d_gpu), intent(inout) :: this
    ising2d_gpu_stat = b200mc_ising_torus_update(this%h_)
  end subroutine update_ising2d_gpu
  impure subroutine update_n_ising2d_gpu(this, n_sweeps)
    class(ising2d_gpu), intent(inout) :: this
    integer(int32), intent(in) :: n_sweeps
    ising2d_gpu_stat = b200mc_ising_torus_update_n(this%h_, n_sweeps)
  end subroutine update_n_ising2d_gpu
  !> :208-231
  impure integer(int64) function nx_ising2d_gpu(this) result(res)
    class(ising2d_gpu), intent(in) :: this
    res = b200mc_ising_torus_nx(this%h_)
  end function nx_ising2d_gpu
  impure integer(int64) function ny_ising2d_gpu(this) result(res)
    class(ising2d_gpu), intent(in) :: this
    res = b200mc_ising_torus_ny(this%h_)
  end function ny_ising2d_gpu
  impure integer(int64) function nall_ising2d_gpu(this) result(res)
    class(ising2d_gpu), intent(in) :: this
    res = b200mc_ising_torus_nall(this%h_)
  end function nall_ising2d_gpu
  impure real(real64) function kbt_ising2d_gpu(this) result(res)
    class(ising2d_gpu), intent(in) :: this
    res = b200mc_ising_torus_kbt(this%h_)
  end function kbt_ising2d_gpu
  impure real(real64) function beta_ising2d_gpu(this) result(res)
    class(ising2d_gpu), intent(in) :: this
    res = b200mc_ising_torus_beta(this%h_)
  end function beta_ising2d_gpu
  !> :232-236 -- the torus has no halo cells: res(1 : nall), site (x, y) at 1 + x + nx y
  impure function spins_ising2d_gpu(this) result(res)
    class(ising2d_gpu), intent(in) :: this
    integer(int32), allocatable :: res(:)
    integer(int64) :: nall
    nall = b200mc_ising_torus_nall(this%h_)
    allocate(res(1 : nall))
    ising2d_gpu_stat = b200mc_ising_torus_get_spins(this%h_, res)
  end function spins_ising2d_gpu
  !> :239-257
  impure integer(int64) function calc_energy_sum_ising2d_gpu(this) result(res)
    class(ising2d_gpu), intent(in) :: this
    ising2d_gpu_stat = b200mc_ising_torus_calc_energy_sum(this%h_, res)
  end function calc_energy_sum_ising2d_gpu
  !> :259-276
  impure integer(int64) function calc_magne_sum_ising2d_gpu(this) result(res)
    class(ising2d_gpu), intent(in) :: this
    ising2d_gpu_stat = b200mc_ising_torus_calc_magne_sum(this%h_, res)
  end function calc_magne_sum_ising2d_gpu
end module ising2d_periodic_gpu_m
