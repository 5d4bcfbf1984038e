! test_crs_ising with the sweep on the GPU: the reference program (test_crs_ising.f90) with ONE changed `use` and the two
! changed calls marked below; the integrand dfunc_ising_discr (test_crs_ising.f90:176-218) is no longer needed on the host.
! Everything between the banner and the call is the reference's own setup and must stay byte-identical, because the
! parameter blob `par` (nodes | weights | integral id) is what the device integrand reads.
! NOT COMPILED IN THE BUILD CONTAINER (no Fortran compiler); ttcross_b200/programs/test_crs_ising.cpp is its C++ twin.
program main
    use tt_lib
    use dmrgg_cuda_lib                                    ! was: use dmrgg_lib
    use quad_lib
    use default_lib
    use time_lib
    implicit none
    include 'mpif.h'
    type(dtt) :: tt, qq
    integer :: i, m, n, r, piv, info, nproc, me, adj
    integer(kind=8) :: neval
    double precision :: acc, val, tru, t1, t2, tcrs
    double precision, allocatable :: par(:)
    character(len=1) :: a
    logical :: rescale
    double precision, parameter :: tpi = 6.28318530717958647692528676655900577d0

    call readarg(1, a, 'c'); call readarg(2, m, 6); call readarg(3, n, 65); call readarg(4, r, 20); call readarg(5, piv, 1)
    call mpi_init(info); call mpi_comm_size(MPI_COMM_WORLD, nproc, info); call mpi_comm_rank(MPI_COMM_WORLD, me, info)
    if (nproc > 1) call dmrgg_cuda_comm_init(nproc, me)   ! one MPI rank per GPU; NCCL id broadcast over MPI
    adj = 0; if (mod(n, 2) == 0) then; n = n + 1; adj = 1; end if        ! test_crs_ising.f90:40
    acc = 500*epsilon(1.d0)
    allocate (par(2*n + 1))
    select case (a)
    case ('c', 'C'); par(2*n + 1) = 1
    case ('d', 'D'); par(2*n + 1) = 2
    case ('e', 'E'); par(2*n + 1) = 3
    case default; write (*, *) 'unknown integral type:', a; stop
    end select
    tru = 0.d0                                            ! analytic values: test_crs_ising.f90:71-100 (copy as needed)
    call lgwt(n, par(1), par(n + 1))                      ! test_crs_ising.f90:102-104
    par(n + 1:2*n) = par(n + 1:2*n)*0.5d0; par(1:n) = (par(1:n) + 1.d0)/2
    rescale = (a == 'd' .or. a == 'e' .or. a == 'D' .or. a == 'E') .and. m >= 10
    qq%l = 1; qq%m = m - 1; qq%n = n; qq%r = 1; call alloc(qq)          ! test_crs_ising.f90:130-144
    val = dble(n/2); if (rescale) val = 5.d0*val
    par(n + 1:2*n) = par(n + 1:2*n)*val
    do i = 1, m - 1; qq%u(i)%p = 1.d0/dble(n/2); end do

    t1 = timef()
    tt%l = 1; tt%m = m - 1; tt%n = n; tt%r = 1; call alloc(tt)
    ! was: call dtt_dmrgg(tt, dfunc_ising_discr, par, maxrank=r, accuracy=acc, pivoting=piv, neval=neval, quad=qq)
    call dtt_dmrgg_cuda(tt, TTC_ISING, par, 2*n + 1, maxrank=r, accuracy=acc, pivoting=piv, neval=neval, quad=qq, device=me)
    t2 = timef(); tcrs = t2 - t1
    if (me == 0) write (*, '(a,i12,a,e12.4,a)') '...with', neval, ' evaluations completed in ', tcrs, ' sec.'
    val = dtt_quad_cuda(tt)                               ! was: val = dtt_quad(tt, qq)
    if (me == 0) write (*, '(a,e50.40)') 'computed value:', val
    call dmrgg_cuda_finalize(); call dealloc(tt); call dealloc(qq)
    call mpi_finalize(info)
    if (me == 0) write (*, '(a)') 'Good bye.'
end program
