!=============================================================================
! collision_b200.f90 -- drop-in replacement of the reference's collision.f90
! (UDel-CFD/D3Q19-Single-Phase, Channel-Flow/collision.f90).
!
! It defines the SAME argument-less subroutines that main.f90's time loop calls
!     collision_MRT   (collision.f90:24)      main.f90:76,157
!     macrovar        (collision.f90:378)     main.f90:102,136,161
!     rhoupdat        (collision.f90:469)     main.f90:74
!     avedensity      (collision.f90:487)     main.f90:166
!     FORCING         (collision.f90:515)     main.f90:61,132
!     FORCINGP        (collision.f90:529)     (no live call site)
! and forwards each of them through ISO_C_BINDING to libd3q19b200.so
! (include/d3q19_b200.h), whose device-resident populations are authoritative
! between calls.  main.f90, para.f90, var_inc.f90, initial.f90 and saveload.f90
! stay as they are; link this file instead of collision.f90:
!
!     mpif90 -O3 -r8 var_inc.f90 main.f90 para.f90 collision_b200.f90 \
!            initial.f90 saveload.f90 -L<repo>/d3q19-single-phase_b200 -ld3q19b200
!
! Host-array coherence (SURVEY.md section 8(b)):
!   * f is uploaded lazily by the first call that needs it (after initpop /
!     loadcntdflow) and written back to the host f only when the driver is about to
!     read it: in the last pre-relaxation iteration (for saveinitflow, main.f90:101;
!     the library predicts the loop exit of main.f90:85 from the same max|rho-rhop|)
!     and on request (d3q19_b200_sync_f_to_host, to be called before savecntdflow if
!     that call at main.f90:224 is re-enabled).
!   * rho,ux,uy,uz come back on the steps where the driver reads them: the
!     initialisation calls, mod(istep,ndiag)==0, mod(istep,nflowout)==0,
!     mod(istep,ntime)==0, the last step (probe, main.f90:221) and avedensity steps.
!
! Topology: the library decomposes in z only (x is never split; y is periodic inside
! the kernel), so para.f90:219 must read `nprocY = 1` (then nprocZ = nproc, one rank
! per GPU).  The shim stops with a message otherwise.
!
! NOTE: no Fortran compiler exists in the build image of this repository.  What stands in
! for one: oracle/shim2c.py (a) checks every bind(c) interface and the d3q19_config mirror
! below against include/d3q19_b200.h (tests/test_fortran_shim.py) and (b) translates the
! executable part of this file to C, which is linked with the machine-translated reference
! MINUS its collision.f90; the reference's own PROGRAM main then runs through this shim and
! is compared with the all-reference build, bit for bit with cfg%math = STRICT
! (tests/test_reference_driver.py: host-sim build on the build box, the real library on a B200).
!=============================================================================
      module d3q19_b200_shim
      use iso_c_binding
      implicit none

      integer(c_int32_t), parameter :: D3Q19_ABI_VERSION = 1
      integer(c_int32_t), parameter :: D3Q19_SCHEME_AUTO = 2
      integer(c_int32_t), parameter :: D3Q19_MATH_FAST = 0

      ! mirror of d3q19_config (include/d3q19_b200.h)
      type, bind(c) :: d3q19_config
        integer(c_int32_t) :: abi_version
        integer(c_int32_t) :: lx, ly, lz
        integer(c_int32_t) :: nx, ny, nz
        integer(c_int32_t) :: globalz
        integer(c_int32_t) :: rank, nranks
        integer(c_int32_t) :: device
        integer(c_int32_t) :: scheme
        integer(c_int32_t) :: math
        integer(c_int32_t) :: ipart
        integer(c_int32_t) :: overlap
        integer(c_int32_t) :: nccl_max_ctas, pf_blocks, halo_timeout_s, halo_split_min, force_idx64   ! tuning knobs, 0 = default
        real(c_double) :: s1, s2, s4, s9, s10, s13, s16
        real(c_double) :: omegepsl, omegepslj, omegxx
        real(c_double) :: rhopart
        real(c_double) :: reserved_d(5)
        integer(c_signed_char) :: nccl_id(128)
      end type d3q19_config

      type(c_ptr), save :: handle = c_null_ptr
      logical, save :: bound = .false.

      interface
        integer(c_int) function d3q19_create(cfg, out) bind(c, name='d3q19_create')
          import :: c_int, c_ptr, d3q19_config
          type(d3q19_config), intent(in) :: cfg
          type(c_ptr), intent(out) :: out
        end function
        integer(c_int) function d3q19_destroy(h) bind(c, name='d3q19_destroy')
          import :: c_int, c_ptr
          type(c_ptr), value :: h
        end function
        integer(c_int) function d3q19_nccl_unique_id(id) bind(c, name='d3q19_nccl_unique_id')
          import :: c_int, c_signed_char
          integer(c_signed_char), intent(out) :: id(128)
        end function
        integer(c_int) function d3q19_device_count(n) bind(c, name='d3q19_device_count')
          import :: c_int, c_int32_t
          integer(c_int32_t), intent(out) :: n
        end function
        function d3q19_last_error() bind(c, name='d3q19_last_error') result(msg)
          import :: c_ptr
          type(c_ptr) :: msg
        end function
        integer(c_int) function d3q19_shim_bind_arrays(h, f, rho, ux, uy, uz, frx, fry, frz, ibnodes, isnodes,   &
            has_isnodes, ndiag, nflowout, nsteps, istep0, ntime, maxiter, rhoepsl)                              &
            bind(c, name='d3q19_shim_bind_arrays')
          import :: c_int, c_ptr, c_int32_t, c_double
          type(c_ptr), value :: h
          ! the arrays go by reference, as Fortran passes them: the library keeps their addresses
          ! (allocarray allocates them once and never frees them, para.f90:418-503)
          real(c_double) :: f(*), rho(*), ux(*), uy(*), uz(*), frx(*), fry(*), frz(*)
          integer(c_int32_t) :: ibnodes(*), isnodes(*)
          integer(c_int32_t), value :: has_isnodes, ndiag, nflowout, nsteps, istep0, ntime, maxiter
          real(c_double), value :: rhoepsl
        end function
        integer(c_int) function d3q19_shim_set_schedule(h, ndiag, nflowout, nsteps, istep0) &
            bind(c, name='d3q19_shim_set_schedule')
          import :: c_int, c_ptr, c_int32_t
          type(c_ptr), value :: h
          integer(c_int32_t), value :: ndiag, nflowout, nsteps, istep0
        end function
        integer(c_int) function d3q19_shim_forcing(h, force_in_y, force_mag) bind(c, name='d3q19_shim_forcing')
          import :: c_int, c_ptr, c_double
          type(c_ptr), value :: h
          real(c_double), value :: force_in_y, force_mag
        end function
        integer(c_int) function d3q19_set_force_field(h, fx, fy, fz) bind(c, name='d3q19_set_force_field')
          import :: c_int, c_ptr, c_double
          type(c_ptr), value :: h
          real(c_double), intent(in) :: fx(*), fy(*), fz(*)
        end function
        integer(c_int) function d3q19_forcingp(h, istep, force_in_y) bind(c, name='d3q19_forcingp')
          import :: c_int, c_ptr, c_double, c_int32_t
          type(c_ptr), value :: h
          integer(c_int32_t), value :: istep
          real(c_double), value :: force_in_y
        end function
        integer(c_int) function d3q19_shim_rhoupdat(h) bind(c, name='d3q19_shim_rhoupdat')
          import :: c_int, c_ptr
          type(c_ptr), value :: h
        end function
        integer(c_int) function d3q19_shim_collision_mrt(h) bind(c, name='d3q19_shim_collision_mrt')
          import :: c_int, c_ptr
          type(c_ptr), value :: h
        end function
        integer(c_int) function d3q19_shim_macrovar(h, istep) bind(c, name='d3q19_shim_macrovar')
          import :: c_int, c_ptr, c_int32_t
          type(c_ptr), value :: h
          integer(c_int32_t), value :: istep
        end function
        integer(c_int) function d3q19_shim_avedensity(h) bind(c, name='d3q19_shim_avedensity')
          import :: c_int, c_ptr
          type(c_ptr), value :: h
        end function
        integer(c_int) function d3q19_shim_sync_f_to_host(h) bind(c, name='d3q19_shim_sync_f_to_host')
          import :: c_int, c_ptr
          type(c_ptr), value :: h
        end function
        integer(c_int) function d3q19_shim_sync_f_to_device(h) bind(c, name='d3q19_shim_sync_f_to_device')
          import :: c_int, c_ptr
          type(c_ptr), value :: h
        end function
        integer(c_int) function d3q19_ipc_export(h, blob) bind(c, name='d3q19_ipc_export')
          import :: c_int, c_ptr, c_signed_char
          type(c_ptr), value :: h
          integer(c_signed_char), intent(out) :: blob(256)
        end function
        integer(c_int) function d3q19_set_halo_mode(h, mode) bind(c, name='d3q19_set_halo_mode')
          import :: c_int, c_ptr, c_int32_t
          type(c_ptr), value :: h
          integer(c_int32_t), value :: mode
        end function
        integer(c_int) function d3q19_ipc_connect(h, blobs) bind(c, name='d3q19_ipc_connect')
          import :: c_int, c_ptr, c_signed_char
          type(c_ptr), value :: h
          integer(c_signed_char), intent(in) :: blobs(*)
        end function
        integer(c_size_t) function c_strlen(s) bind(c, name='strlen')
          import :: c_size_t, c_ptr
          type(c_ptr), value :: s
        end function
      end interface

      contains

!     ---- error convention: the reference's subroutines cannot fail (SURVEY 8(b));
!     a nonzero status is printed and the run is aborted
      subroutine d3q19_b200_check(rc, what)
      use mpi
      integer(c_int), intent(in) :: rc
      character(len=*), intent(in) :: what
      type(c_ptr) :: cmsg
      character(kind=c_char), pointer :: fmsg(:)
      integer :: n, ierr_
      if (rc == 0) return
      cmsg = d3q19_last_error()
      n = int(c_strlen(cmsg))
      call c_f_pointer(cmsg, fmsg, [n])
      write(*,*) 'd3q19_b200: ', what, ' failed: ', fmsg(1:n)
      call MPI_ABORT(MPI_COMM_WORLD, 1, ierr_)
      end subroutine d3q19_b200_check

!     ---- lazy creation: para and allocarray have run by the first hot-path call
      subroutine d3q19_b200_ensure
      use mpi
      use var_inc
      type(d3q19_config) :: cfg
      integer(c_int32_t) :: ndev
      integer :: ierr_
      integer(c_signed_char) :: ipc_mine(256)
      integer(c_signed_char), allocatable :: ipc_all(:)
      if (bound) return
      if (nprocY /= 1) then
        if (myid == 0) write(*,*) 'd3q19_b200: set nprocY = 1 in para.f90:219 (z-slab decomposition, one rank per GPU)'
        call MPI_ABORT(MPI_COMM_WORLD, 1, ierr_)
      endif
      call d3q19_b200_check(d3q19_device_count(ndev), 'd3q19_device_count')
      cfg%abi_version = D3Q19_ABI_VERSION
      cfg%lx = lx;  cfg%ly = ly;  cfg%lz = lz
      cfg%nx = nx;  cfg%ny = ny;  cfg%nz = nz
      cfg%globalz = globalz
      cfg%rank = indz
      cfg%nranks = nprocZ
      cfg%device = mod(myid, max(ndev, 1))
      cfg%scheme = D3Q19_SCHEME_AUTO
      cfg%math = D3Q19_MATH_FAST
      cfg%ipart = merge(1, 0, ipart)
      cfg%overlap = 1
      cfg%nccl_max_ctas = 0; cfg%pf_blocks = 0; cfg%halo_timeout_s = 0; cfg%halo_split_min = 0; cfg%force_idx64 = 0
      cfg%s1 = s1;  cfg%s2 = s2;  cfg%s4 = s4;  cfg%s9 = s9
      cfg%s10 = s10;  cfg%s13 = s13;  cfg%s16 = s16
      cfg%omegepsl = omegepsl;  cfg%omegepslj = omegepslj;  cfg%omegxx = omegxx
      cfg%rhopart = rhopart
      cfg%reserved_d = 0.0d0
      cfg%nccl_id = 0
      if (nprocZ > 1) then
        ! the bootstrap the reference's MPI job already has: rank 0 makes the id, everybody gets it
        if (myid == 0) call d3q19_b200_check(d3q19_nccl_unique_id(cfg%nccl_id), 'd3q19_nccl_unique_id')
        call MPI_BCAST(cfg%nccl_id, 128, MPI_BYTE, 0, MPI_COMM_WORLD, ierr_)
      endif
      call d3q19_b200_check(d3q19_create(cfg, handle), 'd3q19_create')
      ! The z faces travel through NVLink peer memory, moved by the copy engines (D3Q19_HALO_PUT = 1: the fastest
      ! transport in every configuration measured, profiles/r02e_two_gpus.md, r02g_eight_gpus.md): every rank exports
      ! its cudaIpc handles and MPI_ALLGATHER distributes them -- the bootstrap of what then replaces
      ! MPI_ISEND/IRECV/WAITALL (collision.f90:349-356).  d3q19_ipc_connect returns 2 when all ranks agreed that peer
      ! memory cannot be mapped; the faces then travel by NCCL send/recv.
      if (nprocZ > 1 .and. lz >= 2) then
        allocate(ipc_all(256*nproc))
        call d3q19_b200_check(d3q19_set_halo_mode(handle, 1_c_int32_t), 'd3q19_set_halo_mode')
        call d3q19_b200_check(d3q19_ipc_export(handle, ipc_mine), 'd3q19_ipc_export')
        call MPI_ALLGATHER(ipc_mine, 256, MPI_BYTE, ipc_all, 256, MPI_BYTE, MPI_COMM_WORLD, ierr_)
        ierr_ = d3q19_ipc_connect(handle, ipc_all)
        if (ierr_ /= 0 .and. ierr_ /= 2) call d3q19_b200_check(ierr_, 'd3q19_ipc_connect')
        deallocate(ipc_all)
      endif

      ! main.f90:85: the pre-relaxation loop ends after 15000 iterations; rhoepsl: para.f90:285
      if (ipart) then
        call d3q19_b200_check(d3q19_shim_bind_arrays(handle, f, rho, ux, uy, uz, force_realx, force_realy,      &
             force_realz, ibnodes, isnodes, 1, ndiag, nflowout, nsteps, 0, ntime, 15000, rhoepsl), 'd3q19_shim_bind_arrays')
      else
        call d3q19_b200_check(d3q19_shim_bind_arrays(handle, f, rho, ux, uy, uz, force_realx, force_realy,      &
             force_realz, ibnodes, ibnodes, 0, ndiag, nflowout, nsteps, 0, ntime, 15000, rhoepsl), 'd3q19_shim_bind_arrays')
      endif
      bound = .true.
      end subroutine d3q19_b200_ensure

!     ---- for the checkpoint writers/readers of saveload.f90 (optional, see header)
      subroutine d3q19_b200_sync_f_to_host
      call d3q19_b200_ensure
      call d3q19_b200_check(d3q19_shim_sync_f_to_host(handle), 'd3q19_shim_sync_f_to_host')
      end subroutine

      subroutine d3q19_b200_host_f_changed
      call d3q19_b200_ensure
      call d3q19_b200_check(d3q19_shim_sync_f_to_device(handle), 'd3q19_shim_sync_f_to_device')
      end subroutine

      end module d3q19_b200_shim

!=============================================================================
! The subroutines main.f90 calls -- same names, no arguments, as collision.f90
!=============================================================================
      subroutine collision_MRT
      use var_inc
      use d3q19_b200_shim
      implicit none
      call d3q19_b200_ensure
      call d3q19_b200_check(d3q19_shim_collision_mrt(handle), 'collision_MRT')
      end subroutine collision_MRT

      subroutine macrovar
      use var_inc
      use d3q19_b200_shim
      implicit none
      call d3q19_b200_ensure
      ! istep0/nsteps are final only after loading (main.f90:103,120): refresh the schedule
      call d3q19_b200_check(d3q19_shim_set_schedule(handle, ndiag, nflowout, nsteps, istep0), 'set_schedule')
      call d3q19_b200_check(d3q19_shim_macrovar(handle, istep), 'macrovar')
      end subroutine macrovar

      subroutine rhoupdat
      use var_inc
      use d3q19_b200_shim
      implicit none
      call d3q19_b200_ensure
      call d3q19_b200_check(d3q19_shim_rhoupdat(handle), 'rhoupdat')
      end subroutine rhoupdat

      subroutine avedensity
      use var_inc
      use d3q19_b200_shim
      implicit none
      call d3q19_b200_ensure
      call d3q19_b200_check(d3q19_shim_avedensity(handle), 'avedensity')
      end subroutine avedensity

      SUBROUTINE FORCING
      use var_inc
      use d3q19_b200_shim
      implicit none
      call d3q19_b200_ensure
      ! fills force_real{x,y,z} on the host exactly as collision.f90:522-524 and keeps the
      ! force as three kernel scalars on the device
      call d3q19_b200_check(d3q19_shim_forcing(handle, force_in_y, force_mag), 'FORCING')
      END SUBROUTINE FORCING

      SUBROUTINE FORCINGP
      use var_inc
      use d3q19_b200_shim
      implicit none
      ! The perturbation forcing of collision.f90:529-602 is evaluated on the device for the
      ! current istep and written into the device force field: nothing crosses PCIe.  The host
      ! arrays force_real{x,y,z} are read by nothing but collision_MRT / macrovar
      ! (collision.f90:65-67,415-417), which live on the device now, so they are left alone.
      call d3q19_b200_ensure
      call d3q19_b200_check(d3q19_forcingp(handle, int(istep, c_int32_t), real(force_in_y, c_double)), 'FORCINGP')
      END SUBROUTINE FORCINGP
