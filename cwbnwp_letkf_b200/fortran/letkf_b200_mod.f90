!=======================================================================================
! letkf_b200_mod.f90 -- ISO_C_BINDING interface to libletkf_b200.so (include/letkf_b200.h)
! and the replacement body of the grid-point loop of letkf_driver.
!
! SOURCE ONLY: the build environment of this repository has no Fortran compiler (nor MPI,
! NetCDF, SSL2), so this file has not been compiled here.  It uses only standard
! Fortran 2008 + ISO_C_BINDING and follows the reference's own style.  The same C ABI is
! exercised by the Python/ctypes host mirror (cwbnwp_letkf_b200/host.py) with
! Fortran-layout arrays, which is what the parity tests run.
!
! What it replaces in lopunch/CWBNWP-LETKF (see INTEGRATION.md for the exact edits):
!   module_letkf_core.f90:63-64   build_tree (x2)              -> inside letkf_b200_analyze
!   module_letkf_core.f90:209-240 the i/j/k loop: get_lz, letkf_yoyb, letkf_solve
!                                                              -> letkf_b200_analyze
!   module_letkf_core.f90:252-278 letkf_tune_q                 -> cfg%tune_q = 1
!   module_letkf_core.f90:295     destroy_tree                 -> inside letkf_b200_analyze
!   module_eigen.f90:16,110       eigen workspace set-up/tear-down -> letkf_b200_init/finalize
! Everything else of the reference (namelists, NetCDF I/O, observation ingest, MPI
! scatter/gather) stays as it is.
!=======================================================================================
module letkf_b200

    use, intrinsic :: iso_c_binding
    implicit none

    private
    public :: letkf_b200_type_config, letkf_b200_var_config, letkf_b200_stats
    public :: b200_init, b200_finalize, b200_set_gts, b200_set_radar, b200_make_config, b200_analyze
    public :: LETKF_B200_GTS, LETKF_B200_RADAR

    integer(c_int), parameter :: LETKF_B200_GTS   = 0
    integer(c_int), parameter :: LETKF_B200_RADAR = 1
    integer,        parameter :: MAX_SLOTS = 5, MAX_TYPES = 16

    ! mirrors of the C structs (include/letkf_b200.h)
    type, bind(C) :: letkf_b200_type_config
        integer(c_int32_t) :: family, type, use_it, max_lz_pts
        real(c_float)      :: hclr, vclr
        integer(c_int32_t) :: nvar
        integer(c_int32_t) :: is_assim(MAX_SLOTS)
        real(c_float)      :: err_muti(MAX_SLOTS)
        real(c_float)      :: err_rej(MAX_SLOTS)
    end type letkf_b200_type_config

    type, bind(C) :: letkf_b200_var_config
        integer(c_int32_t) :: ntypes, weight_function
        real(c_float)      :: norain_value, multi_infl
        integer(c_int32_t) :: use_rtpp
        real(c_float)      :: rtpp_alpha
        integer(c_int32_t) :: use_rtps
        real(c_float)      :: rtps_alpha
        integer(c_int32_t) :: tune_q
        type(letkf_b200_type_config) :: types(MAX_TYPES)
    end type letkf_b200_var_config

    type, bind(C) :: letkf_b200_stats
        integer(c_int64_t) :: npts, npts_analysed, rows, units
        integer(c_int32_t) :: ntrees, max_sweeps
        real(c_float)      :: ms_tree, ms_search, ms_gram, ms_eigen, ms_transform, ms_total
        integer(c_int64_t) :: sweeps_sum
    end type letkf_b200_stats

    type(c_ptr), save :: ctx = c_null_ptr

    interface
        function letkf_b200_init(ctx, nmember, real64, device) bind(C, name="letkf_b200_init") result(rc)
            import :: c_ptr, c_int
            type(c_ptr),    intent(out)       :: ctx
            integer(c_int), intent(in), value :: nmember, real64, device
            integer(c_int)                    :: rc
        end function
        function letkf_b200_finalize(ctx) bind(C, name="letkf_b200_finalize") result(rc)
            import :: c_ptr, c_int
            type(c_ptr), value :: ctx
            integer(c_int)     :: rc
        end function
        function letkf_b200_last_error() bind(C, name="letkf_b200_last_error") result(msg)
            import :: c_ptr
            type(c_ptr) :: msg
        end function
        function letkf_b200_set_obs(ctx, family, type, n, nvar, xyz, obs, error, hdxb, qc) &
                 bind(C, name="letkf_b200_set_obs") result(rc)
            import :: c_ptr, c_int
            type(c_ptr),    value :: ctx
            integer(c_int), value :: family, type, n, nvar
            type(c_ptr),    value :: xyz, obs, error, hdxb, qc
            integer(c_int)        :: rc
        end function
        function letkf_b200_analyze(ctx, cfg, npts, xyz_grid, nfields, var, stats) &
                 bind(C, name="letkf_b200_analyze") result(rc)
            import :: c_ptr, c_int, c_int64_t, letkf_b200_var_config, letkf_b200_stats
            type(c_ptr),                 value         :: ctx
            type(letkf_b200_var_config), intent(in)    :: cfg
            integer(c_int64_t),          value         :: npts
            type(c_ptr),                 value         :: xyz_grid, var
            integer(c_int),              value         :: nfields
            type(letkf_b200_stats),      intent(out)   :: stats
            integer(c_int)                             :: rc
        end function
        ! npts = ncol*nz with levels slowest: 2-D localised variables then share weights per column
        function letkf_b200_set_levels(ctx, nz) bind(C, name="letkf_b200_set_levels") result(rc)
            import :: c_ptr, c_int
            type(c_ptr),    value :: ctx
            integer(c_int), value :: nz
            integer(c_int)        :: rc
        end function
        ! CUDA runtime: page-lock the work arrays so that the library's asynchronous chunk copies overlap the
        ! analysis (pageable memory works too, but serialises them); flags = 0 (cudaHostRegisterDefault)
        function cudaHostRegister(ptr, nbytes, flags) bind(C, name="cudaHostRegister") result(rc)
            import :: c_ptr, c_int, c_size_t
            type(c_ptr),       value :: ptr
            integer(c_size_t), value :: nbytes
            integer(c_int),    value :: flags
            integer(c_int)           :: rc
        end function
        function cudaHostUnregister(ptr) bind(C, name="cudaHostUnregister") result(rc)
            import :: c_ptr, c_int
            type(c_ptr), value :: ptr
            integer(c_int)     :: rc
        end function
    end interface

contains

    ! reference convention for failures is `stop "message"` (module_letkf_core.f90:161)
    subroutine check(rc)
        integer(c_int), intent(in) :: rc
        character(kind=c_char), pointer :: s(:)
        integer :: n
        if (rc == 0) return
        call c_f_pointer(letkf_b200_last_error(), s, [1024])
        n = 0
        do while (n < 1024)
            if (s(n+1) == c_null_char) exit
            n = n + 1
        end do
        print *, "letkf_b200: ", s(1:n)
        stop "letkf_b200 failed"
    end subroutine check

    ! replaces set_optimal_workspace_for_eigen(nmember) (cwb_letkf.f90:35); one GPU per rank
    subroutine b200_init(nmember, myid, ngpus_per_node)
        integer, intent(in) :: nmember, myid, ngpus_per_node
        integer(c_int)      :: real64
#ifdef REAL64
        real64 = 1
#else
        real64 = 0
#endif
        call check(letkf_b200_init(ctx, int(nmember, c_int), real64, int(mod(myid, ngpus_per_node), c_int)))
    end subroutine b200_init

    ! replaces destroy_eigen_array (cwb_letkf.f90:63)
    subroutine b200_finalize
        call check(letkf_b200_finalize(ctx))
        ctx = c_null_ptr
    end subroutine b200_finalize

    ! one gts platform (module_gts_omboma.f90:13-22); call after wait_jobs (core:50)
    subroutine b200_set_gts(obs_type, nobs, nvar, xyz, obs, error, hdxb, qc)
        integer,                 intent(in)         :: obs_type, nobs, nvar
        real(c_float),   target, intent(in)         :: xyz(:,:), obs(:,:), error(:,:), hdxb(:,:,:)
        integer(c_int),  target, intent(in)         :: qc(:,:,:)
        call check(letkf_b200_set_obs(ctx, LETKF_B200_GTS, int(obs_type, c_int), int(nobs, c_int), &
                                      int(nvar, c_int), c_loc(xyz), c_loc(obs), c_loc(error),      &
                                      c_loc(hdxb), c_loc(qc)))
    end subroutine b200_set_gts

    ! one radar type (module_radar.f90:13-16)
    subroutine b200_set_radar(obs_type, nobs, xyz, obs, hdxb)
        integer,               intent(in) :: obs_type, nobs
        real(c_float), target, intent(in) :: xyz(:,:), obs(:), hdxb(:,:)
        call check(letkf_b200_set_obs(ctx, LETKF_B200_RADAR, int(obs_type, c_int), int(nobs, c_int), 1_c_int, &
                                      c_loc(xyz), c_loc(obs), c_null_ptr, c_loc(hdxb), c_null_ptr))
    end subroutine b200_set_radar

    ! namelist slice of variable ivar (module_config.f90:7-34) -> plain C struct
    subroutine b200_make_config(ivar, is_q, cfg)
        use config
        use param
        integer,                     intent(in)  :: ivar
        logical,                     intent(in)  :: is_q     ! variable goes through letkf_tune_q (core:252-278)
        type(letkf_b200_var_config), intent(out) :: cfg
        integer :: n

        cfg%weight_function = weight_function
        cfg%norain_value    = norain_value
        cfg%multi_infl      = multi_infl(ivar)
        cfg%use_rtpp        = merge(1, 0, use_RTPP(ivar))
        cfg%rtpp_alpha      = RTPP_Alpha(ivar)
        cfg%use_rtps        = merge(1, 0, use_RTPS(ivar))
        cfg%rtps_alpha      = RTPS_Alpha(ivar)
        cfg%tune_q          = merge(1, 0, is_q)
        n = 0
        call add_gts(sound_nml, sound, 4)
        call add_gts(synop_nml, synop, 5)
        call add_gts(gpspw_nml, gpspw, 1)
        call add_gts(metar_nml, metar, 5)
        call add_gts(ships_nml, ships, 5)
        call add_rad(radar_nml%dbz, dbz)
        call add_rad(radar_nml%vr,  vr)
        call add_rad(radar_nml%zdr, zdr)
        call add_rad(radar_nml%kdp, kdp)
        cfg%ntypes = n
    contains
        subroutine add_gts(nml, obs_type, nvar)
            type(gts_config), intent(in) :: nml
            integer,          intent(in) :: obs_type, nvar
            n = n + 1
            associate(t => cfg%types(n))
                t%family = LETKF_B200_GTS;  t%type = obs_type;  t%use_it = merge(1, 0, nml%use_it)
                t%max_lz_pts = nml%max_lz_pts;  t%hclr = nml%hclr(ivar);  t%vclr = nml%vclr(ivar)
                t%nvar = nvar;  t%is_assim = 0;  t%err_muti = 1.;  t%err_rej = 5.
                select case (nvar)     ! slot order of letkf_yoyb (core:356-400,411-417)
                case (5)
                    t%is_assim(1:5) = merge(1, 0, [nml%u%is_assim(ivar), nml%v%is_assim(ivar), nml%t%is_assim(ivar), &
                                                   nml%p%is_assim(ivar), nml%q%is_assim(ivar)])
                    t%err_muti(1:5) = [nml%u%err_muti, nml%v%err_muti, nml%t%err_muti, nml%p%err_muti, nml%q%err_muti]
                    t%err_rej(1:5)  = [nml%u%err_rej,  nml%v%err_rej,  nml%t%err_rej,  nml%p%err_rej,  nml%q%err_rej]
                case (4)
                    t%is_assim(1:4) = merge(1, 0, [nml%u%is_assim(ivar), nml%v%is_assim(ivar), nml%t%is_assim(ivar), &
                                                   nml%q%is_assim(ivar)])
                    t%err_muti(1:4) = [nml%u%err_muti, nml%v%err_muti, nml%t%err_muti, nml%q%err_muti]
                    t%err_rej(1:4)  = [nml%u%err_rej,  nml%v%err_rej,  nml%t%err_rej,  nml%q%err_rej]
                case (1)
                    t%is_assim(1) = merge(1, 0, nml%tpw%is_assim(ivar))
                    t%err_muti(1) = nml%tpw%err_muti
                    t%err_rej(1)  = nml%tpw%err_rej
                end select
            end associate
        end subroutine add_gts
        subroutine add_rad(nml, obs_type)
            type(radar_variable_config), intent(in) :: nml
            integer,                     intent(in) :: obs_type
            n = n + 1
            associate(t => cfg%types(n))
                t%family = LETKF_B200_RADAR;  t%type = obs_type;  t%use_it = merge(1, 0, nml%use_it)
                t%max_lz_pts = nml%max_lz_pts;  t%hclr = nml%hclr(ivar);  t%vclr = nml%vclr(ivar)
                t%nvar = 1;  t%is_assim = 1;  t%err_muti = nml%error;  t%err_rej = nml%err_rej
            end associate
        end subroutine add_rad
    end subroutine b200_make_config

    !-----------------------------------------------------------------------------------
    ! Replacement of module_letkf_core.f90:63-64 + 209-240 (+252-278) + 295 for one variable.
    !   var(loc_nx_arr, loc_ny_arr, nz, 0:nmember-1) as allocated at core:85 (for U / V the array is
    !   one column / row wider than the analysed range, SURVEY Q6: only 1:loc_nx x 1:loc_ny is
    !   analysed, so the analysed points are packed into a contiguous work array first);
    !   lon/lat/alt as at core:165-206; proj as at core:211.
    !-----------------------------------------------------------------------------------
    subroutine b200_analyze(ivar, is_q, proj, loc_nx, loc_ny, nz, lon, lat, alt, var)
        use projection, only : proj_type
        use config,     only : nmember
        integer,         intent(in)    :: ivar, loc_nx, loc_ny, nz
        logical,         intent(in)    :: is_q
        type(proj_type), intent(in)    :: proj
        real,            intent(in)    :: lon(:,:), lat(:,:), alt(:,:,:)
        real,            intent(inout) :: var(:,:,:,0:)

        type(letkf_b200_var_config)        :: cfg
        type(letkf_b200_stats)             :: stats
        real(c_float), allocatable, target :: xyz(:,:), work(:,:)
        real, dimension(2)                 :: xy
        integer                            :: i, j, k, m, n
        integer(c_int64_t)                 :: npts
        logical                            :: pinned
        integer(c_int)                     :: rc_unused

        call b200_make_config(ivar, is_q, cfg)
        npts = int(loc_nx, c_int64_t) * loc_ny * nz
        allocate(xyz(3, npts), work(npts, 0:nmember-1))

        do k = 1, nz                                   ! point order: i fastest, level slowest
        do j = 1, loc_ny
        do i = 1, loc_nx
            n  = i + loc_nx * ((j-1) + loc_ny * (k-1))
            xy = proj % lonlat_to_xy(lon(i,j), lat(i,j))          ! core:211
            xyz(1:2, n) = xy
            xyz(3,   n) = alt(i,j,k)                               ! core:214
            do m = 0, nmember-1
                work(n, m) = var(i,j,k,m)                          ! core:228
            end do
        end do
        end do
        end do

        ! pinned work array: the fields then stream through the device chunk by chunk under the analysis
        pinned = cudaHostRegister(c_loc(work), int(npts, c_size_t) * nmember * 4_c_size_t, 0_c_int) == 0
        ! all levels of a column share (x, y): lets MU / P / PH (vclr <= 0 everywhere) solve once per column
        call check(letkf_b200_set_levels(ctx, int(nz, c_int)))
        call check(letkf_b200_analyze(ctx, cfg, npts, c_loc(xyz), 1_c_int, c_loc(work), stats))
        call check(letkf_b200_set_levels(ctx, 1_c_int))
        if (pinned) rc_unused = cudaHostUnregister(c_loc(work))

        do m = 0, nmember-1
        do k = 1, nz
        do j = 1, loc_ny
        do i = 1, loc_nx
            var(i,j,k,m) = work(i + loc_nx * ((j-1) + loc_ny * (k-1)), m)   ! core:229
        end do
        end do
        end do
        end do
        deallocate(xyz, work)
    end subroutine b200_analyze

end module letkf_b200
