!===============================================================================
! dmrgg_cuda_lib — ISO_C_BINDING shim that lets the reference's Fortran drivers call the
! B200 TT-cross sweep (libttcross_b200.so, include/ttcross_b200.h) with the calling
! convention of dmrgg_lib (reference lib/dmrgg.f90:11-26 and :1261-1267).
!
!   use dmrgg_lib        ->  use dmrgg_cuda_lib
!   call dtt_dmrgg(tt, dfunc_ising_discr, par, maxrank=r, accuracy=acc, pivoting=piv, neval=neval, quad=qq, tru=tru)
!                        ->  call dtt_dmrgg_cuda(tt, TTC_ISING, par, npar, maxrank=r, accuracy=acc, pivoting=piv, &
!                                                neval=neval, quad=qq, tru=tru)
!   val = dtt_quad(tt, qq)  ->  val = dtt_quad_cuda(tt)
!
! A host `external fun` cannot run on the GPU, so the integrand is named by its family
! (TTC_ISING / TTC_STDNORM / TTC_MVN / TTC_COSCOEF) and `par` is the same opaque blob the reference
! hands to `fun` (test_crs_ising.f90:63-69).  Everything else keeps its meaning; the
! callee wipes `arg` to the result exactly like the reference (dmrgg.f90:96-100) and the
! caller frees it with dealloc(tt).
!
! NOT COMPILED IN THE BUILD CONTAINER (no Fortran compiler there, SURVEY F1): build with
!   mpif90 -c dmrgg_cuda_lib.f90 -I<reference module dir> ;  link with -lttcross_b200
! The C++ twins in ttcross_b200/programs/ exercise the same C-ABI calls in CI.
!===============================================================================
module dmrgg_cuda_lib
    use iso_c_binding
    use tt_lib            ! type(dtt), alloc, dealloc (reference lib/tt.f90:18-26, 879-916)
    implicit none
    private
    public :: dtt_dmrgg_cuda, dtt_quad_cuda, dmrgg_cuda_comm_init, dmrgg_cuda_finalize
    public :: ztt_quad_cuda, dtt_ort_cuda, dtt_svd_cuda, dtt_ijk_cuda, dtt_accchk_cuda, dtt_write_cuda
    public :: TTC_ISING, TTC_STDNORM, TTC_MVN, TTC_COSCOEF

    integer(c_int), parameter :: TTC_ISING = 1, TTC_STDNORM = 4, TTC_MVN = 5, TTC_COSCOEF = 6
    type(c_ptr), save :: handle = c_null_ptr          ! one live problem, like the module state of the reference
    logical, save :: comm_wanted = .false.
    integer, save :: comm_size = 1, comm_rank = 0
    character(kind=c_char), save :: comm_id(128)

    interface
        integer(c_int) function ttc_create(out, kind, d, n, par, npar, aux, naux) bind(C, name='ttc_create')
            import; type(c_ptr), intent(out) :: out
            integer(c_int), value :: kind, d; integer(c_int), intent(in) :: n(*)
            real(c_double), intent(in) :: par(*); integer(c_long), value :: npar
            type(c_ptr), value :: aux; integer(c_long), value :: naux
        end function
        subroutine ttc_destroy(h) bind(C, name='ttc_destroy'); import; type(c_ptr), value :: h; end subroutine
        type(c_ptr) function ttc_last_error(h) bind(C, name='ttc_last_error'); import; type(c_ptr), value :: h; end function
        integer(c_int) function ttc_set_device(h, dev) bind(C, name='ttc_set_device'); import; type(c_ptr), value :: h; integer(c_int), value :: dev; end function
        integer(c_int) function ttc_set_partition(h, nparts, own) bind(C, name='ttc_set_partition')
            import; type(c_ptr), value :: h; integer(c_int), value :: nparts; type(c_ptr), value :: own
        end function
        integer(c_int) function ttc_set_quad(h, quad) bind(C, name='ttc_set_quad'); import; type(c_ptr), value :: h; real(c_double), intent(in) :: quad(*); end function
        integer(c_int) function ttc_set_tru(h, present, tru) bind(C, name='ttc_set_tru'); import; type(c_ptr), value :: h; integer(c_int), value :: present; real(c_double), value :: tru; end function
        integer(c_int) function ttc_set_seed(h, seed) bind(C, name='ttc_set_seed'); import; type(c_ptr), value :: h; integer(c_long_long), value :: seed; end function
        integer(c_int) function ttc_set_verbose(h, v) bind(C, name='ttc_set_verbose'); import; type(c_ptr), value :: h; integer(c_int), value :: v; end function
        integer(c_int) function ttc_dmrgg(h, maxrank, accuracy, pivoting) bind(C, name='ttc_dmrgg')
            import; type(c_ptr), value :: h; integer(c_int), value :: maxrank, pivoting; real(c_double), value :: accuracy
        end function
        integer(c_int) function ttc_ranks(h, r) bind(C, name='ttc_ranks'); import; type(c_ptr), value :: h; integer(c_int), intent(out) :: r(*); end function
        integer(c_int) function ttc_core(h, k, out) bind(C, name='ttc_core'); import; type(c_ptr), value :: h; integer(c_int), value :: k; real(c_double), intent(out) :: out(*); end function
        integer(c_int) function ttc_core_range(h, first, last) bind(C, name='ttc_core_range'); import; type(c_ptr), value :: h; integer(c_int), intent(out) :: first, last; end function
        integer(c_long_long) function ttc_neval(h) bind(C, name='ttc_neval'); import; type(c_ptr), value :: h; end function
        integer(c_int) function ttc_quad(h, val) bind(C, name='ttc_quad'); import; type(c_ptr), value :: h; real(c_double), intent(out) :: val; end function
        integer(c_int) function ttc_comm_unique_id(id) bind(C, name='ttc_comm_unique_id'); import; character(kind=c_char), intent(out) :: id(128); end function
        integer(c_int) function ttc_comm_init(h, nranks, rank, id) bind(C, name='ttc_comm_init')
            import; type(c_ptr), value :: h; integer(c_int), value :: nranks, rank; character(kind=c_char), intent(in) :: id(128)
        end function
        integer(c_int) function ttc_quad_complex(h, nsets, wre, wim, ore, oim) bind(C, name='ttc_quad_complex')
            import; type(c_ptr), value :: h; integer(c_int), value :: nsets
            real(c_double), intent(in) :: wre(*), wim(*); real(c_double), intent(out) :: ore(*), oim(*)
        end function
        integer(c_int) function ttc_ort(h) bind(C, name='ttc_ort'); import; type(c_ptr), value :: h; end function
        integer(c_size_t) function c_strlen(s) bind(C, name='strlen'); import; type(c_ptr), value :: s; end function
        integer(c_int) function ttc_set_exp_mode(h, mode) bind(C, name='ttc_set_exp_mode'); import; type(c_ptr), value :: h; integer(c_int), value :: mode; end function
        integer(c_int) function ttc_converged(h) bind(C, name='ttc_converged'); import; type(c_ptr), value :: h; end function
        integer(c_int) function ttc_svd(h, tol, rmax) bind(C, name='ttc_svd'); import; type(c_ptr), value :: h; real(c_double), value :: tol; integer(c_int), value :: rmax; end function
        integer(c_int) function ttc_values(h, count, ind, values) bind(C, name='ttc_values')
            import; type(c_ptr), value :: h; integer(c_long_long), value :: count; integer(c_int), intent(in) :: ind(*); real(c_double), intent(out) :: values(*)
        end function
        integer(c_int) function ttc_accchk(h, nlot, seed, out4, pivot) bind(C, name='ttc_accchk')
            import; type(c_ptr), value :: h; integer(c_long_long), value :: nlot, seed; real(c_double), intent(out) :: out4(4); type(c_ptr), value :: pivot
        end function
        integer(c_int) function ttc_write(h, path) bind(C, name='ttc_write'); import; type(c_ptr), value :: h; character(kind=c_char), intent(in) :: path(*); end function
    end interface

contains

    ! reference error convention: write(*,*) msg; stop   (e.g. dmrgg.f90:114-117)
    subroutine check(st, what)
        integer(c_int), intent(in) :: st
        character(len=*), intent(in) :: what
        character(kind=c_char), pointer :: msg(:)
        type(c_ptr) :: cmsg
        integer :: i, n
        if (st == 0) return
        cmsg = ttc_last_error(handle)
        write (*, '(3a)', advance='no') what, ': '
        if (c_associated(cmsg)) then
            n = int(c_strlen(cmsg))                       ! exactly the bytes of the C string, never past its terminator
            if (n > 0) then
                call c_f_pointer(cmsg, msg, [n])
                do i = 1, n
                    write (*, '(a)', advance='no') msg(i)
                end do
            end if
        end if
        write (*, *)
        stop
    end subroutine

    ! Replaces mpi_init + MPI_COMM_WORLD of the reference drivers (test_crs_ising.f90:31-36): every MPI rank drives one GPU.
    ! Only records the geometry; the NCCL id is created per cross (fresh_comm_id below).
    subroutine dmrgg_cuda_comm_init(nproc, me)
        integer, intent(in) :: nproc, me
        comm_wanted = .true.; comm_size = nproc; comm_rank = me
    end subroutine

    ! An NCCL unique id is single use: every dtt_dmrgg_cuda call creates a new handle and therefore a new communicator, so rank 0
    ! makes a FRESH id and MPI broadcasts it right before ttc_comm_init (the call is collective like dtt_dmrgg itself).
    subroutine fresh_comm_id()
        include 'mpif.h'
        integer :: info
        integer(c_int) :: st
        if (comm_rank == 0) then
            st = ttc_comm_unique_id(comm_id)
            if (st /= 0) then; write (*, *) 'dtt_dmrgg_cuda: cannot create the NCCL id'; stop; end if
        end if
        call mpi_bcast(comm_id, 128, MPI_CHARACTER, 0, MPI_COMM_WORLD, info)
        if (info /= 0) then; write (*, *) 'dtt_dmrgg_cuda: mpi_bcast fail: ', info; stop; end if
    end subroutine

    subroutine dtt_dmrgg_cuda(arg, kind, par, npar, accuracy, maxrank, mybonds, pivoting, neval, quad, tru, aux, seed, device, verbose)
        type(dtt), intent(inout), target :: arg
        integer, intent(in) :: kind                                   ! TTC_ISING / TTC_STDNORM / TTC_MVN  (was: external fun)
        double precision, intent(in) :: par(*)
        integer, intent(in) :: npar
        double precision, intent(in), optional :: accuracy
        integer, intent(in), optional :: maxrank
        integer, intent(in), optional, target :: mybonds(0:)          ! mybonds(0:nparts): bonds owned by partition p (dmrgg.f90:126-130)
        integer, intent(in), optional :: pivoting
        integer(kind=8), intent(out), optional :: neval
        type(dtt), intent(in), optional :: quad
        double precision, intent(in), optional :: tru
        double precision, intent(in), optional, target :: aux(:)      ! MVN: mu(d) | inv_cov(d,d) | denominator (mvn_pdf.f90:4-11)
        integer(kind=8), intent(in), optional :: seed                 ! lottery stream (rnd.f90:120 is unseeded in the reference)
        integer, intent(in), optional :: device, verbose

        integer :: l, m, d, k, p, piv, mr, nparts, first, last
        double precision :: acc
        integer(c_int), allocatable :: nn(:), rr(:)
        double precision, allocatable :: qw(:)
        integer(c_int) :: st
        type(c_ptr) :: auxp
        integer(c_long) :: naux

        l = arg%l; m = arg%m; d = m - l + 1
        if (d < 2) then; write (*, *) 'dtt_dmrgg_cuda: need at least two cores'; stop; end if
        if (c_associated(handle)) call ttc_destroy(handle)
        allocate (nn(d), rr(0:d))
        nn = arg%n(l:m)
        auxp = c_null_ptr; naux = 0
        if (present(aux)) then; auxp = c_loc(aux); naux = size(aux); end if
        st = ttc_create(handle, int(kind, c_int), int(d, c_int), nn, par, int(npar, c_long), auxp, naux)
        if (st /= 0) then; handle = c_null_ptr; call check(st, 'ttc_create'); end if
        if (present(device)) call check(ttc_set_device(handle, int(device, c_int)), 'ttc_set_device')
        if (present(mybonds)) then
            nparts = ubound(mybonds, 1)
            call check(ttc_set_partition(handle, int(nparts, c_int), c_loc(mybonds)), 'ttc_set_partition')
        else if (comm_wanted) then
            call check(ttc_set_partition(handle, int(comm_size, c_int), c_null_ptr), 'ttc_set_partition')   ! share(), default.f90:80-97
        end if
        if (comm_wanted) then
            call fresh_comm_id()
            call check(ttc_comm_init(handle, int(comm_size, c_int), int(comm_rank, c_int), comm_id), 'ttc_comm_init')
        end if
        if (present(quad)) then                                       ! rank-1 weights as one dense vector n(1)+...+n(d)
            allocate (qw(sum(nn)))
            k = 0
            do p = l, m
                qw(k + 1:k + arg%n(p)) = quad%u(p)%p(1, 1:arg%n(p), 1)
                k = k + arg%n(p)
            end do
            call check(ttc_set_quad(handle, qw), 'ttc_set_quad')
        end if
        if (present(tru)) call check(ttc_set_tru(handle, 1_c_int, tru), 'ttc_set_tru')
        if (present(seed)) call check(ttc_set_seed(handle, int(seed, c_long_long)), 'ttc_set_seed')
        st = 1; if (present(verbose)) st = int(verbose, c_int)
        if (comm_wanted .and. comm_rank /= 0) st = 0                  ! only rank 0 prints the sweep lines (dmrgg.f90:293-300)
        call check(ttc_set_verbose(handle, st), 'ttc_set_verbose')
        acc = -1.d0; if (present(accuracy)) acc = accuracy
        mr = -1; if (present(maxrank)) mr = maxrank
        piv = 3; if (present(pivoting)) piv = pivoting                ! default(3, pivoting), dmrgg.f90:59
        call check(ttc_dmrgg(handle, int(mr, c_int), acc, int(piv, c_int)), 'dtt_dmrgg')

        ! results back into the caller's type(dtt): ranks, then the cores this rank holds (all of them on one GPU)
        call check(ttc_ranks(handle, rr), 'ttc_ranks')
        arg%r(l - 1:m) = rr(0:d)
        call alloc(arg)                                               ! tt.f90:879-892: u(k)%p(r(k-1), n(k), r(k))
        call check(ttc_core_range(handle, first, last), 'ttc_core_range')
        do k = first, last
            call check(ttc_core(handle, int(k, c_int), arg%u(l + k - 1)%p), 'ttc_core')
        end do
        if (present(neval)) neval = ttc_neval(handle)
    end subroutine

    ! dtt_quad(arg, quad) of the train computed by the last dtt_dmrgg_cuda (weights: those passed as quad=, else plain sums);
    ! collective when a communicator is attached; like the reference the value is valid on rank 0 (dmrgg.f90:1413).
    double precision function dtt_quad_cuda(arg) result(val)
        type(dtt), intent(in) :: arg
        real(c_double) :: v
        if (.not. c_associated(handle)) then; write (*, *) 'dtt_quad_cuda: no cross has been computed'; stop; end if
        call check(ttc_quad(handle, v), 'dtt_quad')
        val = v
    end function

    ! reads ranks and cores of the handle's train back into the caller's type(dtt) (after ort / svd changed them)
    subroutine fetch_train(arg)
        type(dtt), intent(inout) :: arg
        integer :: l, m, d, k, first, last
        integer(c_int), allocatable :: rr(:)
        l = arg%l; m = arg%m; d = m - l + 1
        allocate (rr(0:d))
        call check(ttc_ranks(handle, rr), 'ttc_ranks')
        arg%r(l - 1:m) = rr(0:d)
        call alloc(arg)
        call check(ttc_core_range(handle, first, last), 'ttc_core_range')
        do k = first, last
            call check(ttc_core(handle, int(k, c_int), arg%u(l + k - 1)%p), 'ttc_core')
        end do
    end subroutine

    ! ztt_quad(tt_z, qq) (dmrgg.f90:1418-1523) for nsets rank-1 complex weight sets at once; w(:, s) holds the
    ! n(1)+...+n(d) weights of set s.  test_crs_chf.f90:153-168 becomes one call with nsets = 32.
    subroutine ztt_quad_cuda(nsets, w, ans)
        integer, intent(in) :: nsets
        double complex, intent(in) :: w(:, :)
        double complex, intent(out) :: ans(nsets)
        real(c_double), allocatable :: wre(:), wim(:), ore(:), oim(:)
        if (.not. c_associated(handle)) then; write (*, *) 'ztt_quad_cuda: no cross has been computed'; stop; end if
        allocate (wre(size(w)), wim(size(w)), ore(nsets), oim(nsets))
        wre = reshape(dble(w), [size(w)]); wim = reshape(aimag(w), [size(w)])
        call check(ttc_quad_complex(handle, int(nsets, c_int), wre, wim, ore, oim), 'ztt_quad')
        ans = dcmplx(ore, oim)
    end subroutine

    ! call ort(tt) = dtt_ort (tt.f90:130-198) on the train of the last dtt_dmrgg_cuda; arg receives the orthogonalised cores
    subroutine dtt_ort_cuda(arg)
        type(dtt), intent(inout) :: arg
        call check(ttc_ort(handle), 'dtt_ort')
        call fetch_train(arg)
    end subroutine

    ! call svd(tt, tol, rmax) = dtt_svd (tt.f90:307-368); arg receives the rounded train (ranks shrink)
    subroutine dtt_svd_cuda(arg, tol, rmax)
        type(dtt), intent(inout) :: arg
        double precision, intent(in) :: tol
        integer, intent(in), optional :: rmax
        integer :: mr
        mr = 0; if (present(rmax)) mr = rmax
        call check(ttc_svd(handle, tol, int(mr, c_int)), 'dtt_svd')
        call dealloc(arg)
        call fetch_train(arg)
    end subroutine

    ! dtt_ijk(tt, ind) (tt.f90:630-660) for count multi-indices at once; ind(d, count), 1-based like the reference
    subroutine dtt_ijk_cuda(count, ind, values)
        integer, intent(in) :: count
        integer, intent(in) :: ind(:, :)
        double precision, intent(out) :: values(count)
        integer(c_int), allocatable :: flat(:)
        allocate (flat(size(ind)))
        flat = reshape(ind, [size(ind)])
        call check(ttc_values(handle, int(count, c_long_long), flat, values), 'dtt_ijk')
    end subroutine

    ! dtt_accchk(nlot, tt, einf, efro, ainf, afro, fun, par, pivot) (dmrgg.f90:1081-1160) with a seeded index stream
    subroutine dtt_accchk_cuda(nlot, einf, efro, ainf, afro, seed)
        integer, intent(in) :: nlot
        double precision, intent(out) :: einf, efro, ainf, afro
        integer(kind=8), intent(in), optional :: seed
        real(c_double) :: out4(4)
        integer(c_long_long) :: sd
        sd = 1; if (present(seed)) sd = seed
        call check(ttc_accchk(handle, int(nlot, c_long_long), sd, out4, c_null_ptr), 'dtt_accchk')
        einf = out4(1); efro = out4(2); ainf = out4(3); afro = out4(4)
    end subroutine

    ! call write(tt, fnam) = dtt_write (ttio.f90:29-108): same file, byte for byte
    subroutine dtt_write_cuda(fnam)
        character(len=*), intent(in) :: fnam
        call check(ttc_write(handle, trim(fnam)//c_null_char), 'dtt_write')
    end subroutine

    subroutine dmrgg_cuda_finalize()
        if (c_associated(handle)) call ttc_destroy(handle)
        handle = c_null_ptr
    end subroutine
end module
