! fortran/us3d_inflow_example.f90 -- where a US3D-style plugin calls the generator (model: the
! reference's us3d_user.f90:21-48 my_user_init, 51-130 my_user_main_pre, 184-201 user_initialize).
! Not compiled here (needs the proprietary US3D modules); it documents the call shape only.
module inflow_df_plugin
    use DIGITAL_FILTERING
    implicit none
    type(digital_filter_type), save :: df
    integer, allocatable, save :: face_j(:), face_k(:)     ! inflow face -> (j,k) of the filter plane
contains
    subroutine my_user_init()                              ! once, like us3d_user.f90:21-48
        type(DFConfig) :: config
        config%d_i = 0.0013d0; config%rho_e = 0.044d0; config%U_e = 869.1d0; config%mu_e = 7.1212d-6
        config%vel_fluc_file = '../files/RST.dat'
        config%line_file = '../line.dat'
        df = create_digital_filter(config)
    end subroutine my_user_init

    subroutine my_user_main_pre(dt, nface, ife, u, v, w, t, r, umean, vmean, wmean, tmean, rmean)
        ! per timestep, like us3d_user.f90:51-130: ghost cell ii = ife(j,2) gets mean + fluctuation
        real(8), intent(in) :: dt
        integer, intent(in) :: nface, ife(:, :)
        real(8), intent(inout) :: u(:), v(:), w(:), t(:), r(:)
        real(8), intent(in) :: umean(:), vmean(:), wmean(:), tmean(:), rmean(:)
        integer :: f, ii, idx
        call filter(df, dt)
        do f = 1, nface
            ii = ife(f, 2)                                 ! us3d_user.f90:92
            idx = (face_j(f) - 1) * df%Nz + face_k(f)      ! df.f90:608-610
            u(ii) = umean(f) + df%u%fluc(idx)              ! us3d_user.f90:106-114
            v(ii) = vmean(f) + df%v%fluc(idx)
            w(ii) = wmean(f) + df%w%fluc(idx)
            t(ii) = tmean(f) + df%T_fluc(idx)
            r(ii) = rmean(f) + df%rho_fluc(idx)
        end do
    end subroutine my_user_main_pre
end module inflow_df_plugin
