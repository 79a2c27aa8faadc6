! fortran/digital_filtering.f90 -- module DIGITAL_FILTERING over the B200 library (iso_c_binding).
!
! Drop-in for digital-filtering-fortran/df/df.f90: the same public names
!     public :: digital_filter_type, create_digital_filter, filter, DFConfig          (df.f90:6)
! so that a caller written like test/fortran-main.f90:8-26
!     type(digital_filter_type) :: df;  type(DFConfig) :: config
!     config%d_i = ...;  DF = create_digital_filter(config);  call filter(df, dt)
! compiles unchanged; results are read as DF%u%fluc(:), DF%v%fluc(:), DF%w%fluc(:), DF%T_fluc(:),
! DF%rho_fluc(:) in the reference's order idx = (j-1)*Nz + k (df.f90:608-610).
! Unlike the C++ reference, the Fortran one HONOURS DFConfig (df.f90:80-87): so does this module
! (honor_flow_config = 1).  The arithmetic follows the C++ code, which is normative for parity
! (SURVEY quirk 9).  Link with -ldfb200.
!
! Beyond the reference's surface (all optional):
!   create_digital_filter_batch / filter_batch   BASELINE config 5: P planes of one geometry behind one handle (one launch set per step)
!   config%k_begin / k_end, comm_unique_id, comm_init, gather_begin / gather_end / gathered_field
!                                                 BASELINE config 4: this rank's spanwise slab + the NCCL hand-off inside the library
!   face_map, apply_inflow                        N3: inflow face -> plane cell, ghost cell = mean + fluctuation (us3d_user.f90:85-114)
!
! NOTE: NEVER COMPILED -- no Fortran compiler exists in the build image or on the GPU boxes.  The binding is exercised through
! tests/fortran_abi_mimic.c, a C program that makes exactly these calls with exactly this struct layout and by-reference
! argument passing; tests/test_facades.py checks that the bind(C) type mirrors struct dfb_config field for field.
module DIGITAL_FILTERING
    use, intrinsic :: iso_c_binding
    implicit none
    private
    public :: digital_filter_type, create_digital_filter, filter, DFConfig, FilterField, destroy_digital_filter
    public :: digital_filter_batch_type, create_digital_filter_batch, filter_batch
    public :: comm_unique_id, comm_init, gather_begin, gather_end, gathered_field, face_map, apply_inflow

    integer, parameter :: dp = selected_real_kind(15)

    type :: FilterField                               ! df.f90:45-57 (the members a caller reads)
        real(kind=dp), allocatable :: fluc(:), filt(:)
        integer :: Ny_max = 0, Nz_max = 0
        real(kind=dp) :: Lt = 0.0_dp
    end type FilterField

    type :: DFConfig                                  ! df.f90:59-63
        real(kind=dp) :: d_i = 0.0_dp, rho_e = 0.0_dp, U_e = 0.0_dp, mu_e = 0.0_dp
        integer :: vel_file_offset = 0, vel_file_N_values = 0
        character(len=256) :: grid_file = ' ', vel_fluc_file = ' '
        ! extensions (not in the reference)
        character(len=256) :: line_file = ' '
        integer(c_int64_t) :: seed = 0_c_int64_t
        integer :: device = -1, plane_id = 0
        integer :: k_begin = 0, k_end = 0             ! this rank's spanwise slab [k_begin, k_end) (0-based, C convention); 0,0 = the whole plane
    end type DFConfig

    type :: digital_filter_type
        type(c_ptr) :: handle = c_null_ptr
        integer :: Ny = 0, Nz = 0, n_cells = 0
        type(FilterField) :: u, v, w
        real(kind=dp), allocatable :: rho_fluc(:), T_fluc(:)
        real(kind=dp) :: dt = 0.0_dp
    end type digital_filter_type

    ! BASELINE config 5: nplanes independent planes of one geometry; arrays are (n_cells, nplanes), plane p uses stream group plane_id + p - 1
    type :: digital_filter_batch_type
        type(c_ptr) :: handle = c_null_ptr
        integer :: Ny = 0, Nz = 0, n_cells = 0, nplanes = 0
        real(kind=dp), allocatable :: u(:, :), v(:, :), w(:, :), T(:, :), rho(:, :)
    end type digital_filter_batch_type

    ! image of `struct dfb_config` (include/dfb200.h) -- keep the two in step
    type, bind(C) :: dfb_config_c
        real(c_double) :: d_i, rho_e, U_e, mu_e
        integer(c_int) :: vel_file_offset, vel_file_N_values
        type(c_ptr) :: grid_file
        integer(c_int) :: grid_file_len
        type(c_ptr) :: vel_fluc_file
        integer(c_int) :: vel_fluc_file_len
        integer(c_int) :: struct_bytes, honor_flow_config
        type(c_ptr) :: line_file
        integer(c_int) :: line_file_len
        integer(c_int) :: Ny, Nz, geom_per_row
        type(c_ptr) :: yc, dy, dz, rows, scales, N_y, N_z
        integer(c_int64_t) :: seed
        integer(c_int) :: noise_mode, device, plane_id, k_begin, k_end, skip_first_step, kernel_variant
    end type dfb_config_c

    interface
        integer(c_int) function dfb_config_init(cfg) bind(C, name='dfb_config_init')
            import :: c_int, dfb_config_c
            type(dfb_config_c), intent(inout) :: cfg
        end function
        integer(c_int) function dfb_create_f(cfg, handle) bind(C, name='dfb_create_f')
            import :: c_int, c_ptr, dfb_config_c
            type(dfb_config_c), intent(in) :: cfg
            type(c_ptr), intent(out) :: handle
        end function
        integer(c_int) function dfb_dims_f(handle, Ny, Nz) bind(C, name='dfb_dims_f')
            import :: c_int, c_ptr
            type(c_ptr), intent(in) :: handle
            integer(c_int), intent(out) :: Ny, Nz
        end function
        integer(c_int) function dfb_filter_to_host_f(handle, dt, u, v, w, T, rho) bind(C, name='dfb_filter_to_host_f')
            import :: c_int, c_ptr, c_double
            type(c_ptr), intent(in) :: handle
            real(c_double), intent(in) :: dt
            real(c_double), intent(inout) :: u(*), v(*), w(*), T(*), rho(*)
        end function
        integer(c_int) function dfb_get_field(handle, which, dst, on_device) bind(C, name='dfb_get_field')
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: handle
            integer(c_int), value :: which, on_device
            real(c_double), intent(inout) :: dst(*)
        end function
        integer(c_int) function dfb_create_batch_f(cfg, nplanes, handle) bind(C, name='dfb_create_batch_f')
            import :: c_int, c_ptr, dfb_config_c
            type(dfb_config_c), intent(in) :: cfg
            integer(c_int), intent(in) :: nplanes
            type(c_ptr), intent(out) :: handle
        end function
        integer(c_int) function dfb_host_register(ptr, bytes) bind(C, name='dfb_host_register')
            import :: c_int, c_ptr, c_size_t
            type(c_ptr), value :: ptr
            integer(c_size_t), value :: bytes
        end function
        integer(c_int) function dfb_comm_unique_id(id) bind(C, name='dfb_comm_unique_id')
            import :: c_int, c_char
            character(kind=c_char), intent(out) :: id(128)
        end function
        integer(c_int) function dfb_comm_init_f(handle, id, rank, world) bind(C, name='dfb_comm_init_f')
            import :: c_int, c_ptr, c_char
            type(c_ptr), intent(in) :: handle
            character(kind=c_char), intent(in) :: id(128)
            integer(c_int), intent(in) :: rank, world
        end function
        integer(c_int) function dfb_gather_begin_f(handle, dst_rank) bind(C, name='dfb_gather_begin_f')
            import :: c_int, c_ptr
            type(c_ptr), intent(in) :: handle
            integer(c_int), intent(in) :: dst_rank
        end function
        integer(c_int) function dfb_gather_end_f(handle) bind(C, name='dfb_gather_end_f')
            import :: c_int, c_ptr
            type(c_ptr), intent(in) :: handle
        end function
        integer(c_int) function dfb_gathered_to_host_f(handle, which, dst) bind(C, name='dfb_gathered_to_host_f')
            import :: c_int, c_ptr, c_double
            type(c_ptr), intent(in) :: handle
            integer(c_int), intent(in) :: which
            real(c_double), intent(inout) :: dst(*)
        end function
        integer(c_int) function dfb_face_map_f(handle, n, yf, zf, plane_index) bind(C, name='dfb_face_map_f')
            import :: c_int, c_ptr, c_double
            type(c_ptr), intent(in) :: handle
            integer(c_int), intent(in) :: n
            real(c_double), intent(in) :: yf(*), zf(*)
            integer(c_int), intent(inout) :: plane_index(*)
        end function
        integer(c_int) function dfb_destroy_f(handle) bind(C, name='dfb_destroy_f')
            import :: c_int, c_ptr
            type(c_ptr), intent(inout) :: handle
        end function
        function dfb_last_error() bind(C, name='dfb_last_error') result(msg)
            import :: c_ptr
            type(c_ptr) :: msg
        end function
    end interface

contains

    subroutine check(rc, where)
        integer(c_int), intent(in) :: rc
        character(len=*), intent(in) :: where
        character(kind=c_char), pointer :: s(:)
        integer :: n
        if (rc == 0) return
        call c_f_pointer(dfb_last_error(), s, [512])
        n = 1
        do while (n < 512 .and. s(n) /= c_null_char)
            n = n + 1
        end do
        write(*, '(a,a,a,i0,a)', advance='no') 'DIGITAL_FILTERING: ', where, ' failed (', rc, '): '
        write(*, *) s(1:n-1)
        stop 1                                         ! the reference stops on I/O errors too (df.f90:327-330)
    end subroutine check

    subroutine fill_config(config, c)
        type(DFConfig), intent(in), target :: config
        type(dfb_config_c), intent(inout) :: c
        call check(dfb_config_init(c), 'dfb_config_init')
        c%d_i = config%d_i; c%rho_e = config%rho_e; c%U_e = config%U_e; c%mu_e = config%mu_e
        c%vel_file_offset = config%vel_file_offset; c%vel_file_N_values = config%vel_file_N_values
        c%honor_flow_config = 1
        c%grid_file = c_loc(config%grid_file);          c%grid_file_len = len_trim(config%grid_file)
        c%vel_fluc_file = c_loc(config%vel_fluc_file);  c%vel_fluc_file_len = len_trim(config%vel_fluc_file)
        c%line_file = c_loc(config%line_file);          c%line_file_len = len_trim(config%line_file)
        c%seed = config%seed; c%device = config%device; c%plane_id = config%plane_id
        c%k_begin = config%k_begin; c%k_end = config%k_end
    end subroutine fill_config

    ! page-lock an array the library copies into every step (failure only costs speed)
    ! page-lock an array the library copies into every step (a failure only costs speed)
    subroutine pin(a, n)
        integer, intent(in) :: n
        real(kind=dp), intent(in), target :: a(n)
        integer(c_int) :: rc
        rc = dfb_host_register(c_loc(a), int(8, c_size_t) * int(n, c_size_t))
    end subroutine pin

    ! df.f90:74-138
    function create_digital_filter(config) result(DF)
        type(DFConfig), intent(in), target :: config
        type(digital_filter_type) :: DF
        type(dfb_config_c) :: c
        integer(c_int) :: Ny, Nz

        call fill_config(config, c)
        call check(dfb_create_f(c, DF%handle), 'create_digital_filter')
        call check(dfb_dims_f(DF%handle, Ny, Nz), 'dfb_dims')
        DF%Ny = Ny; DF%Nz = Nz; DF%n_cells = Ny * Nz
        allocate(DF%u%fluc(DF%n_cells), DF%v%fluc(DF%n_cells), DF%w%fluc(DF%n_cells))
        allocate(DF%u%filt(DF%n_cells), DF%v%filt(DF%n_cells), DF%w%filt(DF%n_cells))
        allocate(DF%rho_fluc(DF%n_cells), DF%T_fluc(DF%n_cells))
        DF%rho_fluc = 0.0_dp; DF%T_fluc = 0.0_dp
        call pin(DF%u%fluc, DF%n_cells); call pin(DF%v%fluc, DF%n_cells); call pin(DF%w%fluc, DF%n_cells)     ! page-locked: filter()'s
        call pin(DF%T_fluc, DF%n_cells); call pin(DF%rho_fluc, DF%n_cells)                                     ! copies run at the PCIe rate
        ! fluctuations of the first step (df.f90:124-131)
        call check(dfb_get_field(DF%handle, 0_c_int, DF%u%fluc, 0_c_int), 'dfb_get_field')
        call check(dfb_get_field(DF%handle, 1_c_int, DF%v%fluc, 0_c_int), 'dfb_get_field')
        call check(dfb_get_field(DF%handle, 2_c_int, DF%w%fluc, 0_c_int), 'dfb_get_field')
    end function create_digital_filter

    ! df.f90:621-651
    subroutine filter(DF, dt_input)
        type(digital_filter_type), intent(inout) :: DF
        real(kind=dp), intent(in) :: dt_input
        DF%dt = dt_input
        call check(dfb_filter_to_host_f(DF%handle, dt_input, DF%u%fluc, DF%v%fluc, DF%w%fluc, DF%T_fluc, DF%rho_fluc), 'filter')
    end subroutine filter

    ! ---- BASELINE config 5: nplanes planes of one geometry, one launch set per step for all of them ----
    function create_digital_filter_batch(config, nplanes) result(B)
        type(DFConfig), intent(in), target :: config
        integer, intent(in) :: nplanes
        type(digital_filter_batch_type) :: B
        type(dfb_config_c) :: c
        integer(c_int) :: Ny, Nz, np
        call fill_config(config, c)
        np = nplanes
        call check(dfb_create_batch_f(c, np, B%handle), 'create_digital_filter_batch')
        call check(dfb_dims_f(B%handle, Ny, Nz), 'dfb_dims')
        B%Ny = Ny; B%Nz = Nz; B%n_cells = Ny * Nz; B%nplanes = nplanes
        allocate(B%u(B%n_cells, nplanes), B%v(B%n_cells, nplanes), B%w(B%n_cells, nplanes), B%T(B%n_cells, nplanes), B%rho(B%n_cells, nplanes))
    end function create_digital_filter_batch

    subroutine filter_batch(B, dt_input)          ! every plane advances by dt_input; the fields of all planes land in B%u(:, p) ...
        type(digital_filter_batch_type), intent(inout) :: B
        real(kind=dp), intent(in) :: dt_input
        call check(dfb_filter_to_host_f(B%handle, dt_input, B%u, B%v, B%w, B%T, B%rho), 'filter_batch')
    end subroutine filter_batch

    ! ---- BASELINE config 4: the hand-off of the finished plane to the CFD rank (NCCL inside the library) ----
    subroutine comm_unique_id(id)                  ! rank 0; broadcast the 128 characters with MPI_Bcast
        character(kind=c_char), intent(out) :: id(128)
        call check(dfb_comm_unique_id(id), 'comm_unique_id')
    end subroutine comm_unique_id

    subroutine comm_init(DF, id, rank, world)      ! every rank (0-based rank, as MPI gives it)
        type(digital_filter_type), intent(inout) :: DF
        character(kind=c_char), intent(in) :: id(128)
        integer, intent(in) :: rank, world
        integer(c_int) :: r, w
        r = rank; w = world
        call check(dfb_comm_init_f(DF%handle, id, r, w), 'comm_init')
    end subroutine comm_init

    subroutine gather_begin(DF, dst_rank)          ! after filter(DF, dt); the next filter(DF, dt) may follow at once
        type(digital_filter_type), intent(inout) :: DF
        integer, intent(in) :: dst_rank
        integer(c_int) :: d
        d = dst_rank
        call check(dfb_gather_begin_f(DF%handle, d), 'gather_begin')
    end subroutine gather_begin

    subroutine gather_end(DF)
        type(digital_filter_type), intent(inout) :: DF
        call check(dfb_gather_end_f(DF%handle), 'gather_end')
    end subroutine gather_end

    subroutine gathered_field(DF, which, plane)    ! destination rank: which = 0..4 (u', v', w', T', rho'), plane(Ny * Nz_global)
        type(digital_filter_type), intent(inout) :: DF
        integer, intent(in) :: which
        real(kind=dp), intent(inout) :: plane(*)
        integer(c_int) :: wsel
        wsel = which
        call check(dfb_gathered_to_host_f(DF%handle, wsel, plane), 'gathered_field')
    end subroutine gathered_field

    ! ---- N3: inflow face -> plane cell, and the ghost-cell update of us3d_user.f90:85-114 ----
    subroutine face_map(DF, n, yf, zf, cell)       ! cell(i) = 1-based index into DF%u%fluc(:) of the cell holding face i, 0 = not in this rank's slab
        type(digital_filter_type), intent(in) :: DF
        integer, intent(in) :: n
        real(kind=dp), intent(in) :: yf(n), zf(n)
        integer, intent(out) :: cell(n)
        integer(c_int) :: nn
        nn = n
        call check(dfb_face_map_f(DF%handle, nn, yf, zf, cell), 'face_map')
        cell = cell + 1
    end subroutine face_map

    subroutine apply_inflow(n, cell, ghost, mean, fluc, q)    ! q(ghost(i)) = mean(i) + fluc(cell(i)): r(ii), u(ii), ... of us3d_user.f90:106-114
        integer, intent(in) :: n, cell(n), ghost(n)
        real(kind=dp), intent(in) :: mean(n), fluc(:)
        real(kind=dp), intent(inout) :: q(:)
        integer :: i
        do i = 1, n
            if (cell(i) > 0) q(ghost(i)) = mean(i) + fluc(cell(i))
        end do
    end subroutine apply_inflow

    subroutine destroy_digital_filter(DF)
        type(digital_filter_type), intent(inout) :: DF
        integer(c_int) :: rc
        rc = dfb_destroy_f(DF%handle)
    end subroutine destroy_digital_filter

end module DIGITAL_FILTERING
