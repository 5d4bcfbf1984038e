/* =============================================================================
 * ttcross_b200.h — C-ABI of the B200-native TT-cross sweep.
 *
 * Drop-in boundary for ONE hot path of aukeschaap/ttcross: the parallel greedy
 * TT-cross sweep `dtt_dmrgg` (reference lib/dmrgg.f90:11-1050) and the
 * quadrature `dtt_quad` (lib/dmrgg.f90:1261-1415).  Plain pointers and sizes
 * only; callable from Fortran `bind(C)` (fortran/dmrgg_cuda_lib.f90), from C/C++
 * (ttcross_b200/drivers/) and from ctypes (ttcross_b200/api.py).
 *
 * Every function returns 0 on success or a non-zero ttc_status; the message of
 * the last failure is available from ttc_last_error().  Where the reference
 * does `write(*,*) msg; stop` (e.g. lib/dmrgg.f90:114-117), this ABI returns the
 * status and the same message; the Fortran shim reproduces the `stop`.
 *
 * All reals are IEEE double, all indices are 32-bit and 1-based exactly as in
 * the reference, all arrays are column-major.
 * ========================================================================== */
#ifndef TTCROSS_B200_H
#define TTCROSS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ttc_handle ttc_handle;

/* integrand families (replace the `external fun` argument, lib/dmrgg.f90:18) */
enum {
    TTC_ISING = 1,   /* test_crs_ising.f90:176-218; C/D/E selected by par(2n+1) = 1/2/3 like the reference */
    TTC_STDNORM = 4, /* test_crs_stdnorm.f90:154-170 */
    TTC_MVN = 5,     /* lib/mvn_pdf.f90:63-83 via test_crs_mvn.f90:156-172; aux = mu(d) | inv_cov(d,d) | denom */
    TTC_COSCOEF = 6  /* lib/coefficients.f90:33-65 (COS-method coefficients of a Gaussian density, test_crs_coscoeff.f90:186);
                        aux = mu(d) | sigma(d,d) column-major | lower | upper; par is ignored (may be NULL); d <= 24 */
};

enum ttc_status {
    TTC_OK = 0,
    TTC_ERR_ARG = 1,        /* bad argument */
    TTC_ERR_NPROC = 2,      /* "nproc exceeds or equal dimension, cannot proceed" (lib/dmrgg.f90:114-117) */
    TTC_ERR_PIVOTING = 3,   /* "unknown pivoting" (lib/dmrgg.f90:590-592) */
    TTC_ERR_CUDA = 4,       /* CUDA runtime failure, or no CUDA device: there is NO CPU fallback */
    TTC_ERR_STATE = 5,      /* call order (e.g. results queried before ttc_dmrgg) */
    TTC_ERR_RANK = 6,       /* rank capacity exceeded */
    TTC_ERR_COMM = 7        /* multi-GPU communicator failure */
};

/* ---- problem description --------------------------------------------------
 * Replaces the (fun, par) pair of dtt_dmrgg (lib/dmrgg.f90:18-19) and the
 * arg%l, arg%m, arg%n fields of type(dtt) (lib/tt.f90:18-26).
 *   d     number of cores (arg%m - arg%l + 1, with l = 1)
 *   n     mode sizes n(1:d)
 *   par   the opaque parameter blob the driver hands to `fun` (copied)
 *   aux   integrand-private state that the reference keeps in module variables
 *         (mvn_pdf.f90:4-11): MVN -> mu(d) | inv_cov(d,d) column-major | denominator
 */
int ttc_create(ttc_handle** out, int kind, int d, const int* n, const double* par, long npar,
               const double* aux, long naux);
void ttc_destroy(ttc_handle* h);
const char* ttc_last_error(const ttc_handle* h);   /* h may be NULL: message of the last failed ttc_create */

/* ---- optional arguments of dtt_dmrgg (lib/dmrgg.f90:19-26) ------------------ */
int ttc_set_device(ttc_handle* h, int cuda_device);                 /* default 0 */
int ttc_set_partition(ttc_handle* h, int nparts, const int* own);   /* mybonds(0:nparts); own == NULL -> share() of lib/default.f90:80-97 */
int ttc_set_par(ttc_handle* h, const double* par, long npar);       /* new parameter blob of the same length (next ttc_dmrgg uploads it) */
int ttc_set_quad(ttc_handle* h, const double* quad);                /* quad=: rank-1 weights, n(1)+...+n(d) doubles; NULL removes */
int ttc_set_tru(ttc_handle* h, int present, double tru);            /* tru=: only switches ' cnv ' to ' err ' in the sweep log */
int ttc_set_seed(ttc_handle* h, unsigned long long seed);           /* uniform stream of the lottery (rnd.f90:120 is unseeded) */
/* Alternative uniform source: cb(ctx, vrank, count, out) must write `count` uniforms in [0,1) for virtual rank `vrank`. */
typedef void (*ttc_uniform_cb)(void* ctx, int vrank, int count, double* out);
int ttc_set_uniform_callback(ttc_handle* h, ttc_uniform_cb cb, void* ctx);
/* 0 (default): lottery on the device, sweeps enqueued asynchronously; 1: lottery on the host exactly as rnd.f90:105-126
 * writes it (one stream synchronisation per bond visit; implied by a uniform callback); 2: device lottery, synchronous.
 * 3: device lottery, asynchronous, but the plain one-thread-per-chain support kernels instead of the warp-wavefront /
 * shared-memory ones.  4: device lottery, asynchronous, one kernel per step of a bond visit instead of the cluster kernel
 * that runs a virtual rank's whole visit list.  All modes produce identical results; modes 1-4 exist to prove that. */
int ttc_set_lottery_mode(ttc_handle* h, int mode);
int ttc_set_verbose(ttc_handle* h, int verbose);                    /* 1: print the reference's per-sweep lines on stdout */
/* exp of the exp-based integrands (stdnorm: test_crs_stdnorm.f90:168, MVN: lib/mvn_pdf.f90:81).  0 (default): the platform's
 * exp, like the reference's libm call.  1: the deterministic + - * routine of include/ttc_detexp.h ("parity mode"): the CPU
 * oracle has the same switch, and with it GPU and CPU runs of these integrands agree bit for bit (CUDA's and glibc's exp
 * differ in the last ulp, which decides near-tied pivots of symmetric integrands). */
int ttc_set_exp_mode(ttc_handle* h, int mode);
/* 1 if the last ttc_dmrgg ended on its accuracy criterion (three consecutive sweeps with pivotmax <= accuracy * amax,
 * dmrgg.f90:1013-1017), 0 if it ended on the rank bound: maxrank when given, else the capacity of 64 that stands in for
 * "no maxrank" (the reference would keep growing; here the run stops and this flag says it did not converge). */
int ttc_converged(const ttc_handle* h);

/* ---- the sweep: dtt_dmrgg (lib/dmrgg.f90:11) -------------------------------
 *   maxrank   <= 0 : absent          accuracy  < 0 : absent
 *   (absent maxrank: the rank capacity is 64 -- the run then ends at rank 64 at the latest; ttc_converged() tells whether the
 *    accuracy criterion of dmrgg.f90:1013-1017 or that capacity ended it)
 *   pivoting  -1 full superblock, 0 one cross, >= 1 rook depth (reference default 3)
 */
int ttc_dmrgg(ttc_handle* h, int maxrank, double accuracy, int pivoting);

/* ---- results (valid after ttc_dmrgg) --------------------------------------- */
int ttc_ranks(const ttc_handle* h, int* r);            /* r(0:d) -> arg%r */
int ttc_core(ttc_handle* h, int k, double* out);       /* arg%u(k)%p : r(k-1)*n(k)*r(k) doubles, column-major, k = 1..d */
int ttc_cores(ttc_handle* h, double* out, long long cap); /* all cores, concatenated in core order (cap = doubles available) */
/* The reference returns the cores INSIDE the call: `arg` is an inout argument of dtt_dmrgg and holds the train on return
 * (lib/dmrgg.f90:11-26, cores reallocated at :676-685).  ttc_bind_cores gives ttc_dmrgg the same shape: `out` (caller-owned,
 * `cap` doubles, must stay valid until rebound or the handle is destroyed) is filled by every later ttc_dmrgg with this
 * process's cores, concatenated as ttc_cores does, before ttc_dmrgg returns.  The library page-locks the buffer once
 * (cudaHostRegister) so the device writes into it directly, and the transfer runs beside the host's log processing; a later
 * ttc_cores(h, out, ...) with the same pointer returns at once.  A run whose cores exceed `cap` returns TTC_ERR_ARG (the
 * results stay available through ttc_core / ttc_cores); cap = sum of maxrank * n(k) * maxrank always suffices.
 * out == NULL removes the binding. */
int ttc_bind_cores(ttc_handle* h, double* out, long long cap);
long long ttc_neval(const ttc_handle* h);              /* neval= */
int ttc_nsweeps(const ttc_handle* h);
double ttc_seconds(const ttc_handle* h);               /* wall time of the last ttc_dmrgg call (timef difference) */
/* per-sweep series, nsweeps+1 entries (entry 0 = the '0::' line): 0 val, 1 n_evals, 2 amax, 3 pivotmax, 4 erank, 5 time */
int ttc_sweep_series(const ttc_handle* h, int which, double* out);
/* pivot tape: one record per bond visit {it, vrank, bond, ii, jj, kk, qq, upd} + the pivot value */
long ttc_pivlog_count(const ttc_handle* h);
int ttc_pivlog(const ttc_handle* h, int* ints8, double* pivots);
long ttc_text(const ttc_handle* h, char* buf, long cap);   /* the sweep log (reference format, dmrgg.f90:293-300,971-1008) */

/* dtt_quad(arg, quad) on the finalised train (lib/dmrgg.f90:1261); uses the weights of ttc_set_quad, or plain sums */
int ttc_quad(ttc_handle* h, double* val);

/* ---- host-side helpers the reference drivers use --------------------------- */
void ttc_lgwt(int n, double* x, double* w);                         /* lib/quad.f90:97-131 */
void ttc_share(int first, int last, int nproc, int* own);           /* lib/default.f90:80-97 */
double ttc_stream_uniform(unsigned long long seed, int vrank, unsigned long long k);  /* the built-in uniform stream */

/* ---- kernel-level entry points (measurement and kernel parity tests) --------
 * Superblock kernel of the pivoting = -1 branch (lib/dmrgg.f90:341-396) on the handle's CURRENT state
 * (after ttc_dmrgg): evaluates a(i,j,k,q) over r(p-1) x n(p) x n(p+1) x r(p+1) for bond p, forms the residual
 * against col*row (K = r(p)) and returns both first-index argmaxes.  store != 0 also writes `a` to HBM.
 *   out_idx[0] = argmax|a| (0-based linear), out_idx[1] = argmax|b|; out_val[0] = a at argmax, out_val[1] = b at argmax
 *   ms      average device time per launch over `reps` launches (CUDA events)
 */
int ttc_superblock_probe(ttc_handle* h, int bond, int store, int reps, long long* out_idx, double* out_val,
                         double* ms, long long* count);
/* variant 0: the tiled kernel with the reference arithmetic (what ttc_dmrgg runs; same as ttc_superblock_probe);
 * 1: the plain one-thread-per-element kernel (cross-check); 2: the tiled kernel with the K = r(p) residual update
 * contracted into DFMA — NOT bit-exact, never used by ttc_dmrgg, measures the FP64 ceiling of the shape. */
int ttc_superblock_probe_ex(ttc_handle* h, int bond, int store, int reps, int variant, long long* out_idx, double* out_val,
                            double* ms, long long* count);
/* Fiber kernel probe: column fiber (isrow=0) or row fiber (isrow=1) of bond p through pivot (ii,jj,kk,qq); returns
 * the fiber values (r(p-1)*n(p) or n(p+1)*r(p+1) doubles) and its residual. */
int ttc_fiber_probe(ttc_handle* h, int bond, int isrow, int ii, int jj, int kk, int qq, double* fiber, double* resid,
                    int reps, double* ms);
/* The device lottery's closed-form cumulative weights, executed on the host (test hook): cells[x] = the 1-based cell that
 * lottery2 (rnd.f90:105-126) picks for uniform u[x] among m cells of weight 1 except the listed zero-weight cells. */
int ttc_lottery_closed_form(int m, const int* zeros_sorted_distinct, int nz, const double* u, int count, int* cells);
/* the same draws by the table-free rule of the cluster kernel (one multiplication decides unless u lies within the
 * rounding window of a boundary, where the boundary is formed by literal sequential addition) */
int ttc_lottery_fast(int m, const int* zeros_sorted_distinct, int nz, const double* u, int count, int* cells);
/* Counters: kernels launched by this handle since creation, device time of the last ttc_dmrgg (CUDA events, ms) */
long long ttc_launch_count(const ttc_handle* h);
/* write `bytes` (> L2 size) of scratch HBM so the next timed run starts with a cold L2 (measurement hygiene) */
int ttc_l2_flush(ttc_handle* h, long long bytes);
double ttc_device_ms(const ttc_handle* h);
/* the persistent sweep kernel (ttc_sweep.cuh): device time of its single launch in the last ttc_dmrgg (0 when the per-sweep
 * schedule ran), and its geometry (CTAs per cluster, threads per CTA); returns 1 when the last run used it. */
double ttc_sweep_kernel_ms(const ttc_handle* h);
int ttc_sweep_geometry(const ttc_handle* h, int* cluster, int* threads);
/* measured FP64 ceiling of the device in TFLOP/s (roofline denominator of the evaluation / residual kernels):
 * fma = 1 DFMA chains, fma = 0 separate DMUL + DADD chains (the reference arithmetic has no FMA contraction) */
int ttc_fp64_peak(int device, int fma, double* tflops);
/* per-kernel-class accounting of the last ttc_dmrgg: names[i] (static strings), launches, total ms (events, only when
 * profiling was enabled with ttc_set_profile(h, 1)) */
int ttc_set_profile(ttc_handle* h, int on);
/* diagnostic device timeline of the next ttc_dmrgg: every kernel stamps %globaltimer when its first CTA starts (id = index
 * into names[]; id 100 = a fused argmax fold finished).  Returns the number of stamps. */
int ttc_set_timeline(ttc_handle* h, int on);
long ttc_timeline(const ttc_handle* h, long cap, int* ids, unsigned long long* t_ns, const char** names, int names_cap);
int ttc_profile(const ttc_handle* h, int cap, const char** names, long long* launches, double* ms);

/* ---- tall-skinny Householder QR: ort0_d (lib/ort.f90:17-81 = LAPACK dgeqrf + dorgqr), SURVEY 8 row a21 ----------
 * a: m x n column-major (leading dimension m) on the HOST; q: m x n orthonormal factor; r: n x n upper triangular with
 * LAPACK's sign convention (zeros below the diagonal).  m < n follows the reference's early return (ort.f90:32-46).
 * Not called by the sweep; it is the kernel of TT orthogonalisation (dtt_ort, lib/tt.f90:130-198).
 * ms (may be NULL): device time of one factorisation, averaged over `reps` runs.  Failure message: ttc_last_error(NULL). */
int ttc_qr_thin(int device, int m, int n, const double* a, double* q, double* r, int reps, double* ms);
/* dtt_ort (lib/tt.f90:130-198): orthogonalise the train of the last ttc_dmrgg from the left, in place on the device (QR of
 * every unfolding, R normalised and pushed into the next core, norms equalised over the cores).  Afterwards ttc_core /
 * ttc_cores / ttc_quad see the orthogonalised train.  First row of SURVEY 8(f); single process only. */
int ttc_ort(ttc_handle* h);
/* dtt_svd (lib/tt.f90:307-368): TT rounding on the device — ttc_ort, then right to left an SVD of every unfolding truncated by
 * chop(tol, rmax) of lib/mat.f90:433-455 (tol < 0 / rmax <= 0: absent).  Ranks shrink; ttc_ranks / ttc_core / ttc_quad see
 * the rounded train.  Single process only; ranks up to 64. */
int ttc_svd(ttc_handle* h, double tol, int rmax);

/* ztt_quad (lib/dmrgg.f90:1418-1523): quadrature of the train against `nsets` COMPLEX rank-1 weight tensors in one launch
 * (test_crs_chf.f90:153-168 and test_crs_pdf.f90:128-190 loop over 32 frequencies).  wre / wim: [nsets][n(1)+...+n(d)],
 * out_re / out_im: [nsets]; single-rank chain order of the reference.  SURVEY 8(f) rank 2; single process only. */
int ttc_quad_complex(ttc_handle* h, int nsets, const double* wre, const double* wim, double* out_re, double* out_im);

/* dtt_ijk (lib/tt.f90:630-652) for `count` multi-indices at once: ind [count][d] (1-based), values [count]. */
int ttc_values(ttc_handle* h, long long count, const int* ind, double* values);
/* dtt_accchk (lib/dmrgg.f90:1081-1166): the train against the integrand at nlot random entries (indices int(u*n)+1 from
 * ttc_stream_uniform(seed, 0x7fffffff, sample*d + position)).  out[0..3] = einf, efro, ainf, afro; pivot[d] (may be
 * NULL) = multi-index of the largest error.  SURVEY 8(f) rank 4 (the checker half); single process only. */
int ttc_accchk(ttc_handle* h, long long nlot, unsigned long long seed, double* out, int* pivot);

/* ---- TT files in the reference's stream format: dtt_write / dtt_read (lib/ttio.f90:10-17, 29-108, 196-296) ----------
 * 128-byte header 'TT      ' | ver | inf | comment | i(8), then l, m, n(l:m), r(l-1:m) as int32 and the cores as float64,
 * little-endian without record markers — files are interchangeable with the reference's.  Host-only helpers
 * (n: d mode sizes, r: d+1 ranks, cores concatenated column-major) and ttc_write for the train a handle holds.
 * Failure message: ttc_last_error(NULL) (ttc_write: ttc_last_error(h)). */
int ttc_tt_write(const char* path, int l, int m, const int* n, const int* r, const double* cores);
int ttc_tt_read_header(const char* path, int* l, int* m, int* n, int* r, int cap, long long* ncore);
int ttc_tt_read_cores(const char* path, double* cores, long long cap);
int ttc_write(ttc_handle* h, const char* path);

/* ---- multi-GPU: one process per GPU, core blocks partitioned over ranks ------
 * Replaces MPI_COMM_WORLD of the reference (lib/dmrgg.f90:86-95, 763-959, 1209-1246, 1355-1405).  The communicator id
 * is an NCCL unique id (128 bytes) created on rank 0 by ttc_comm_unique_id and broadcast by the caller (MPI_Bcast in a
 * Fortran/MPI driver, torch.distributed in bench.py).  The nparts partitions of ttc_set_partition (nparts >= nranks)
 * are mapped block-wise onto the ranks: rank g runs partitions floor(nparts*g/nranks) .. floor(nparts*(g+1)/nranks)-1.
 * Results (pivots, ranks, values) are a function of the partition, not of nranks.  After ttc_comm_init, ttc_dmrgg and
 * ttc_quad are COLLECTIVE: every rank must call them with the same arguments.  Sweep log, ranks, pivot tape, neval
 * and values are complete on every rank; cores are held by their owners only (like the reference, lib/dmrgg.f90:1248-1257):
 * ttc_core_range gives the cores [first, last] of this rank. */
int ttc_comm_unique_id(void* id128);
int ttc_comm_init(ttc_handle* h, int nranks, int rank, const void* id128);
int ttc_comm_rank(const ttc_handle* h, int* nranks, int* rank);
int ttc_core_range(const ttc_handle* h, int* first, int* last);

int ttc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* TTCROSS_B200_H */
