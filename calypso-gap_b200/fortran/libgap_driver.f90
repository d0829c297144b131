! libgap_driver.f90 -- the Fortran side of the drop-in: the reference's entry
! points with their original dummy-argument lists, each body a single call into
! the C ABI of include/gapcu.h through ISO_C_BINDING.  Compile this file in place
! of gappy/libgap/{gap_calc,wacsf,get_bond}.f90 (gappy/setup.py:38-41) and link
! libgapcu.so; f2py derives the same Python signatures as from the reference.
!
! Errors: the message is printed the way the reference prints before it stops.  Where the reference
! itself STOPs (gap_parameters missing / malformed / beyond nsf_max, nsparseX_max: gap_calc.f90:323-341;
! more than 1000 neighbours: gap_calc.f90:107-111) this is a plain STOP (exit status 0); every other
! failure (CUDA error, no device, bad arguments, unknown species) is ERROR STOP 1 so that the shell or
! the scheduler sees a failed run (gapcu_stop below; same rule as csrc/fortran_shim.c).
!
! This image has no Fortran compiler, so the file is exercised only by reading;
! csrc/fortran_shim.c defines the same five external symbols (gfortran mangling,
! all arguments by reference) and is what the tests link.  Keep the two in step.
module gapcu_c_api
    use iso_c_binding
    implicit none
    interface
        integer(c_int) function gapcu_calc(na, species, lat, pos, nsparsex, des_len, theta, mm, qmm, coeff, &
                                           rcut, lgrad, ene, force, stress, variance) bind(C, name='gapcu_calc')
            import :: c_int, c_double
            integer(c_int), value :: na, nsparsex, des_len, lgrad
            real(c_double), value :: rcut
            integer(c_int), intent(in) :: species(*)
            real(c_double), intent(in) :: lat(3,3), pos(*), theta(*), mm(*), qmm(*), coeff(*)
            real(c_double), intent(out) :: ene, force(*), stress(6), variance
        end function
        integer(c_int) function gapcu_read(path, nsparsex, des_len, theta, theta_cap, mm, mm_ld, mm_cols, &
                                           invcmm, invcmm_ld, coeff, coeff_cap) bind(C, name='gapcu_read')
            import :: c_int, c_double, c_char
            character(kind=c_char), intent(in) :: path(*)
            integer(c_int), intent(out) :: nsparsex, des_len
            integer(c_int), value :: theta_cap, mm_ld, mm_cols, invcmm_ld, coeff_cap
            real(c_double), intent(out) :: theta(*), mm(*), invcmm(*), coeff(*)
        end function
        integer(c_int) function gapcu_bond(na, lat, elements, pos, rcut, min_bond) bind(C, name='gapcu_bond')
            import :: c_int, c_double
            integer(c_int), value :: na
            real(c_double), value :: rcut
            real(c_double), intent(in) :: lat(3,3), pos(*)
            integer(c_int), intent(in) :: elements(*)
            real(c_double), intent(out) :: min_bond
        end function
        integer(c_int) function gapcu_car2acsf_table(na, max_neighbor, nf, pos, neighbor, neighbor_count, &
                                                     lgrad, xx, dxdy, strs) bind(C, name='gapcu_car2acsf_table')
            import :: c_int, c_double
            integer(c_int), value :: na, max_neighbor, nf, lgrad
            real(c_double), intent(in) :: pos(*), neighbor(*)
            integer(c_int), intent(in) :: neighbor_count(*)
            real(c_double), intent(out) :: xx(*), dxdy(*), strs(*)
        end function
        ! additive: independent structures over the devices of gapcu_set_devices (C-order arrays)
        integer(c_int) function gapcu_calc_batch(nstruct, natoms, species, lat, pos, rcut, lgrad, &
                                                 ene, force, stress) bind(C, name='gapcu_calc_batch')
            import :: c_int, c_double
            integer(c_int), value :: nstruct, lgrad
            real(c_double), value :: rcut
            integer(c_int), intent(in) :: natoms(*), species(*)
            real(c_double), intent(in) :: lat(*), pos(*)
            real(c_double), intent(out) :: ene(*), force(*), stress(*)
        end function
        integer(c_int) function gapcu_set_devices(n, devices) bind(C, name='gapcu_set_devices')
            import :: c_int
            integer(c_int), value :: n
            integer(c_int), intent(in) :: devices(*)
        end function
        subroutine gapcu_print_last_error() bind(C, name='gapcu_print_last_error')
        end subroutine
    end interface
end module gapcu_c_api

! code: the GAPCU_E* value returned by the C ABI (include/gapcu.h); -1 = EFILE, -2 = ENEIGH, -4 = ELIMIT
SUBROUTINE gapcu_stop(code, limit_is_reference_stop)
    use gapcu_c_api
    implicit none
    integer, intent(in) :: code
    logical, intent(in) :: limit_is_reference_stop
    call gapcu_print_last_error()
    if (code == -1 .or. code == -2 .or. (limit_is_reference_stop .and. code == -4)) stop
    error stop 1
END SUBROUTINE gapcu_stop

! gappy/libgap/gap_calc.f90:1-5
SUBROUTINE FGAP_CALC(NA, SPECIES, LAT, POS, ENE, FORCE, STRESS, VARIANCE, nsparseX, des_len, &
                     theta, MM, qmm, coeff, Rcut, lgrad)
    use gapcu_c_api
    implicit none
    integer :: gapcu_rc
    integer, intent(in) :: NA
    integer, intent(in), dimension(NA) :: SPECIES
    double precision, intent(in), dimension(3,3) :: LAT
    double precision, intent(in), dimension(NA,3) :: POS
    integer, intent(in) :: des_len, nsparseX
    double precision, intent(in), dimension(des_len) :: THETA
    double precision, intent(in), dimension(nsparseX,des_len) :: MM
    double precision, intent(in), dimension(nsparseX,nsparseX) :: QMM
    double precision, intent(in), dimension(nsparseX) :: COEFF
    double precision, intent(in) :: Rcut
    logical, intent(in) :: lgrad
    double precision, intent(out) :: ENE, VARIANCE
    double precision, intent(out), dimension(NA,3) :: FORCE
    double precision, intent(out), dimension(6) :: STRESS
    integer :: ig
    ig = 0
    if (lgrad) ig = 1
    gapcu_rc = gapcu_calc(NA, SPECIES, LAT, POS, nsparseX, des_len, THETA, MM, QMM, COEFF, Rcut, ig, &
                   ENE, FORCE, STRESS, VARIANCE)
    if (gapcu_rc /= 0) call gapcu_stop(gapcu_rc, .false.)
END SUBROUTINE FGAP_CALC

! gappy/libgap/gap_calc.f90:303-314
SUBROUTINE FGAP_READ(nsparseX, des_len, theta, MM, invcmm, coeff)
    use gapcu_c_api
    implicit none
    integer :: gapcu_rc
    integer, parameter :: nsf_max = 100
    integer, parameter :: nsparseX_max = 4000
    integer, intent(out) :: nsparseX, des_len
    double precision, intent(out) :: theta(nsf_max), MM(nsparseX_max, nsf_max)
    double precision, intent(out) :: invcmm(nsparseX_max, nsparseX_max), coeff(nsparseX_max)
    gapcu_rc = gapcu_read('gap_parameters'//c_null_char, nsparseX, des_len, theta, nsf_max, MM, nsparseX_max, nsf_max, &
                   invcmm, nsparseX_max, coeff, nsparseX_max)
    if (gapcu_rc /= 0) call gapcu_stop(gapcu_rc, .true.)
END SUBROUTINE FGAP_READ

! gappy/libgap/get_bond.f90:4-12
SUBROUTINE FGET_BOND(na, lat, elements, pos, rcut, min_bond)
    use gapcu_c_api
    implicit none
    integer :: gapcu_rc
    INTEGER, intent(in) :: na
    REAL(8), intent(in), dimension(3,3) :: lat
    INTEGER, intent(in), dimension(na) :: elements
    REAL(8), intent(in), dimension(na,3) :: pos
    REAL(8), intent(in) :: rcut
    REAL(8), intent(out) :: min_bond
    gapcu_rc = gapcu_bond(na, lat, elements, pos, rcut, min_bond)
    if (gapcu_rc /= 0) call gapcu_stop(gapcu_rc, .false.)
END SUBROUTINE FGET_BOND

! gappy/libgap/wacsf.f90:2-12
SUBROUTINE CAR2ACSF(NA, max_neighbor, nf, pos, neighbor, neighbor_count, xx, dxdy, strs, lgrad)
    use gapcu_c_api
    implicit none
    integer :: gapcu_rc
    INTEGER, intent(in) :: NA, max_neighbor, NF
    REAL(8), intent(in), dimension(NA,3) :: pos
    REAL(8), intent(in), dimension(NA,max_neighbor,6) :: neighbor
    INTEGER, intent(in), dimension(NA) :: neighbor_count
    REAL(8), intent(out), dimension(NF,NA) :: xx
    REAL(8), intent(out), dimension(NF,NA,NA,3) :: dxdy
    REAL(8), intent(out), dimension(3,3,NF,NA) :: strs
    LOGICAL, intent(in) :: lgrad
    integer :: ig
    ig = 0
    if (lgrad) ig = 1
    gapcu_rc = gapcu_car2acsf_table(NA, max_neighbor, nf, pos, neighbor, neighbor_count, ig, xx, dxdy, strs)
    if (gapcu_rc /= 0) call gapcu_stop(gapcu_rc, .false.)
END SUBROUTINE CAR2ACSF

! gappy/libgap/wacsf.f90:798-810 (debug dump; every call site in the reference is commented out)
SUBROUTINE write_array_2dim(n, m, a, name)
    REAL(8), intent(in), dimension(n,m) :: a
    character(*), intent(in) :: name
    integer :: i, j
    open(2244, file=trim(adjustl(name)))
    do i = 1, n
        do j = 1, m
            write(2244, '(F20.10,$)') a(i,j)
        enddo
        write(2244, *)
    enddo
    close(2244)
END SUBROUTINE

! Additive (not in the reference): a CALYPSO-style batch of independent structures in one call,
! what gappy/tools/cgg2.py:111-115 does with one process per structure.  Fortran layouts:
! NATOMS(NS); SPECIES, and the columns of POS3/FORCE3, concatenated over the structures (NTOT
! atoms); LAT3(3,3,NS) with LAT3(:,c,s) = lattice vector c of structure s, POS3(3,NTOT),
! FORCE3(3,NTOT), STRESS6(6,NS) -- i.e. the component index runs fastest, which is the C order
! gapcu_calc_batch expects, so no copy is made.  NDEV > 0 selects devices DEVICES(1:NDEV) first.
SUBROUTINE FGAP_CALC_BATCH(NS, NTOT, NATOMS, SPECIES, LAT3, POS3, Rcut, lgrad, ENE, FORCE3, STRESS6, NDEV, DEVICES)
    use gapcu_c_api
    implicit none
    integer :: gapcu_rc
    integer, intent(in) :: NS, NTOT, NDEV
    integer, intent(in) :: NATOMS(NS), SPECIES(NTOT), DEVICES(*)
    double precision, intent(in) :: LAT3(3,3,NS), POS3(3,NTOT), Rcut
    logical, intent(in) :: lgrad
    double precision, intent(out) :: ENE(NS), FORCE3(3,NTOT), STRESS6(6,NS)
    integer(c_int) :: ig
    ig = 0
    if (lgrad) ig = 1
    if (NDEV > 0) then
        gapcu_rc = gapcu_set_devices(NDEV, DEVICES)
        if (gapcu_rc /= 0) call gapcu_stop(gapcu_rc, .false.)
    endif
    gapcu_rc = gapcu_calc_batch(NS, NATOMS, SPECIES, LAT3, POS3, Rcut, ig, ENE, FORCE3, STRESS6)
    if (gapcu_rc /= 0) call gapcu_stop(gapcu_rc, .false.)
END SUBROUTINE FGAP_CALC_BATCH
