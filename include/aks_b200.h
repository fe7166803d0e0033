/*
 * aks_b200.h -- C ABI of the B200-native SCF inner loop for radial Kohn-Sham atoms.
 *
 * Drop-in boundary for ONE hot path of Theozeud/AtomicKohnSham (Julia): the body of
 * `solve!`'s while loop, i.e. `scf_step!(::ODA, ::KSESolver)`
 * (reference src/solver/groundstate.jl:53-62, src/algorithms/oda/loop.jl:1-116), for a
 * BATCH of independent problems (atoms / ions / functionals) that share one
 * discretization.  The reference has no FFI: these entry points are what a Julia
 * `ccall` shim binds behind `groundstate(...; backend = :b200)`.
 *
 * Conventions
 *   - plain C types only; every function returns an int status (AKS_OK == 0).
 *   - all host arrays are caller-owned and copied at *_create; outputs are written into
 *     caller-allocated buffers in the reference's (Julia, column-major) layouts.
 *   - integer index tables are 1-based, as Julia holds them.
 *   - a ctx is bound to one CUDA device and one stream; calls return after the work
 *     they name has completed; a ctx is not thread-safe, different ctxs are independent.
 */
#ifndef AKS_B200_H
#define AKS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------------------ */
#define AKS_OK 0
#define AKS_EINVAL (-1)   /* bad argument                                              */
#define AKS_ECUDA (-2)    /* CUDA runtime error (see aks_last_error)                   */
#define AKS_ENOTCONV (-3) /* aks_scf_run: at least one problem hit maxiter             */
#define AKS_EDEGEN3 (-4)  /* >2-fold degeneracy (reference error(), optimized.jl:139)   */
#define AKS_ENAN (-5)     /* non-finite value in a problem's energies                   */
#define AKS_ENOSYS (-6)   /* feature not built (e.g. a functional on a Double64 batch)     */
#define AKS_EEIG (-7)     /* an eigenpair the aufbau needs could not be converged and certified */

/* per-problem status word in aks_iter_record.status */
#define AKS_RUNNING 0
#define AKS_SUCCESS 1  /* reference "SUCCESS"  (src/solution/types.jl:123) */
#define AKS_MAXITERS 2 /* reference "MAXITERS"                              */

/* exchange / correlation ids (Libxc numbering for the functionals the reference uses) */
#define AKS_XC_NONE 0
#define AKS_LDA_X 1     /* Slater exchange, src/physics/exchange correlation/SlaterXa.jl */
#define AKS_LDA_C_PW 12 /* PW92 correlation, .../PerdewWang.jl                           */

typedef struct aks_ctx aks_ctx;
typedef struct aks_disc aks_disc;
typedef struct aks_batch aks_batch;

/* ---- context ----------------------------------------------------------------------- */
int aks_ctx_create(int device, aks_ctx** out);
void aks_ctx_destroy(aks_ctx* ctx);
const char* aks_last_error(const aks_ctx* ctx);
/* the CUDA stream (cudaStream_t) all work of this ctx is launched on */
void* aks_ctx_stream(const aks_ctx* ctx);
/* execution options (no reference counterpart; results do not depend on them):
 *   use_graphs   1: the kernel sequence of a step is captured once per batch and replayed as a CUDA graph
 *                (default; environment AKS_NO_GRAPH=1 starts with 0), 0: plain stream launches, < 0: unchanged
 *   check_every  aks_scf_run reads the problems' status words every this many steps (default 4, environment
 *                AKS_CHECK_EVERY); < 1: unchanged                                                         */
int aks_ctx_configure(aks_ctx* ctx, int32_t use_graphs, int32_t check_every);

/* ---- discretization: replaces KSEDiscretization + init_cache! tables --------------------
 * (src/discretization/discretization.jl:270-341, src/fem/integration methods.jl:30-70,
 *  src/fem/generators.jl:79-138).  Sizes: n = p+1 generators, C cells, Nh = p*C-1.       */
typedef struct aks_disc_desc {
  int32_t dtype;  /* 0 = Float64; 1 = Double64 (the reference's T = Double64, discretization.jl:289): every
                     `const double*` table below then holds interleaved (hi, lo) pairs (Julia's
                     Vector{Double64} layout, twice the length); x, w, Pgenx, y, wy, ycell, Pgeny, invshifts
                     may be NULL with Q = G = 0 (batches without a functional only); nspin = 1             */
  int32_t p;      /* ordermax                                                             */
  int32_t C;      /* number of cells = length(mesh)-1                                     */
  int32_t Nh;     /* basis size                                                           */
  int32_t Q;      /* Gauss points per cell  (GaussLegendre.npoints)                       */
  int32_t G;      /* points of the global energy grid (== Q in the reference)             */
  int32_t lh;     /* max angular momentum; L = lh+1 channels                              */
  int32_t nh;     /* eigenpairs per channel                                               */
  int32_t nspin;  /* 1 or 2                                                               */
  const double* mesh;      /* [C+1]                                                       */
  const double* invshifts; /* [2*C]  (c1,c0) per cell: r = c1*x + c0  (basis.invshifts)   */
  const double* x;         /* [Q]    GaussLegendre.x                                      */
  const double* w;         /* [Q]    GaussLegendre.w                                      */
  const double* Pgenx;     /* [(p+1)*Q] column-major (p+1,Q): generator g at node q       */
  const double* y;         /* [G]    GaussLegendre.y                                      */
  const double* wy;        /* [G]    GaussLegendre.wy                                     */
  const int64_t* ycell;    /* [G]    findindex(mesh, y[q]) (1-based cell)                 */
  const double* Pgeny;     /* [(p+1)*G] column-major (p+1,G): generators at phi(y[q])     */
  const double* M0;        /* [Nh*Nh] dense, symmetric (femops.M0)                        */
  const double* A;         /* [Nh*Nh] dense (Matrix(femops.A))                            */
  const double* Mm1;       /* [Nh*Nh] dense (Matrix(femops.M₋₁))                          */
  const double* Mm2;       /* [Nh*Nh] dense or NULL when lh == 0                          */
  int64_t nF;              /* entries of the mass tensor Dict femops.F (0 if no Hartree)  */
  const int64_t* Fi;       /* [nF] 1-based (i,j,m) keys                                   */
  const int64_t* Fj;
  const int64_t* Fm;
  const double* Fv;        /* [nF]                                                        */
  const int64_t* cell_len;        /* [C]   length(cells_to_indices[c])                    */
  const int64_t* cell_indices;    /* [C*(p+1)] row c: cells_to_indices[c], 1-based, 0-padded   */
  const int64_t* cell_generators; /* [C*(p+1)] row c: cells_to_generators[c], 1-based, 0-padded */
} aks_disc_desc;

int aks_disc_create(aks_ctx* ctx, const aks_disc_desc* desc, aks_disc** out);
void aks_disc_destroy(aks_disc* disc);

/* ---- batch of models: replaces KSEModel + KSESolver/ODACache state ----------------------
 * (src/physics/models.jl:19-58, src/algorithms/oda/types.jl:58-107).
 * lh_p / nh_p may be NULL (every problem uses the disc's lh, nh) or give per-problem
 * values <= the disc's (a sweep where basis_size depends on Z).                          */
int aks_batch_create(aks_disc* disc, int32_t nprob, const double* Z, const double* N,
                     const double* hartree, const int32_t* ex_id, const int32_t* ec_id,
                     const int32_t* lh_p, const int32_t* nh_p, aks_batch** out);
void aks_batch_destroy(aks_batch* batch);

/* ODA(...) + OptimizedAufbau(...) options (oda/types.jl:29-45, aufbau/optimized.jl:18-27) */
int aks_batch_set_oda(aks_batch* batch, double scftol, double tinit, int32_t frozen_t,
                      int32_t maxiter_ls, double abstol_ls, double reltol_ls,
                      int32_t max_degen, double degen_tol, int32_t handle_degen_spin_1iter);

/* The reference chooses the algorithm and the iteration cap per problem
 * (examples/atomic density comparison/run.jl:40-47,74-76: iron gets ODA(tinit = 0.15, frozen_t = true,
 * scftol = 1e-12) and maxiter = 800, every other atom the defaults).  Overrides scftol, tinit, frozen_t of ONE
 * problem of the batch (aks_batch_set_oda resets every problem to its arguments) and sets the cap `maxiter` on
 * that problem's iteration count, enforced on the device: the problem stops with status AKS_MAXITERS when its
 * niter reaches it (0 = no per-problem cap: only the maxiter of aks_scf_run applies).  Takes effect from the
 * problem's next step; tinit also replaces its current t.  Not built for Double64 batches (AKS_ENOSYS).       */
int aks_batch_set_oda_problem(aks_batch* batch, int32_t prob, double scftol, double tinit, int32_t frozen_t,
                              int32_t maxiter);

/* occupation scheme (src/algorithms/aufbau/): kind 0 = OptimizedAufbau (default, options of
 * aks_batch_set_oda), 1 = SmearedAufbau(Temp = temp) (aufbau/smeared.jl:68-101), 2 = FrozenAufbau with the
 * cache's nfix (aufbau/frozen.jl:36-88): nfix[nprob][(L, nh[, 2]) column-major], L = lh+1, nh of the disc  */
int aks_batch_set_aufbau(aks_batch* batch, int32_t kind, double temp, const double* nfix);

/* Eigenpair window (no reference counterpart; the results are the reference's).  The reference solves
 * all nh eigenpairs of every channel in every iteration (oda/loop.jl:65-77) although only the levels
 * the aufbau can reach matter for the iteration.  With margin >= 0 a step solves, per channel, the
 * levels occupied in the previous steps + margin, certifies on the device (one inertia count per
 * channel) that no unsolved level could have been occupied (else solves the rest and repeats the aufbau
 * within the same step), and the remaining eigenpairs of the last Hamiltonian are solved when
 * aks_get_state asks for eps / U.  margin < 0: every step solves all nh eigenpairs.  Default 0.
 * Double64 batches: margin >= 0 switches their (certified) window on, margin < 0 off; its size is chosen on the device. */
int aks_batch_set_eig_window(aks_batch* batch, int32_t margin);

/* back to niter = 0, D = 0 (KSESolver constructor, src/solver/solver.jl:61-88) */
int aks_batch_reset(aks_batch* batch);

/* what LogBook / LogFileCallback read after every iteration (src/solver/logging.jl:26-37) */
typedef struct aks_iter_record {
  int32_t niter;  /* iterations completed                                  */
  int32_t status; /* AKS_RUNNING / AKS_SUCCESS / AKS_MAXITERS / <0 error   */
  double stop;    /* stopping criterion of the last step                   */
  double Etot, Ekin, Ecou, Ehar, Eexc;
  double t;       /* ODA mixing parameter of the last step                 */
} aks_iter_record;

/* one scf_step! for every problem still running.  With out == NULL the step is only enqueued
 * on the ctx stream (no host synchronisation); otherwise out[nprob] receives the records.  */
int aks_scf_step(aks_batch* batch, aks_iter_record* out);

/* solve!: iterate until every problem converged or reached maxiter.  history may be NULL
 * or [maxiter*nprob] (record of problem i after iteration k at history[k*nprob+i]);
 * final[nprob] receives the last record of each problem.  Convergence is decided on the device
 * (per-problem status word); the host reads the status words every few steps (AKS_CHECK_EVERY, default 4)
 * instead of after every step, steps enqueued for a finished problem do nothing, and the history is written
 * on the device and copied once at the end.                                                */
int aks_scf_run(aks_batch* batch, int32_t maxiter, aks_iter_record* history,
                aks_iter_record* final_records);

/* copy one problem's state out, Julia layouts; any pointer may be NULL:
 *   D (Nh,Nh[,2])  U (Nh,nh,L[,2])  eps,n (L,nh[,2])  W (Nh) = A\B of the final D
 *   (postcomputations!, oda/loop.jl:51-54; multiplied by the model's hartree coefficient) */
int aks_get_state(aks_batch* batch, int32_t prob, double* D, double* U, double* eps, double* n,
                  double* W);
/* Batches on a dtype = 1 (Double64) discretization: D, U, eps, n, W of aks_get_state / aks_get_state_all are
 * (hi, lo) pairs in the same layouts (buffers twice as long); aks_iter_record carries the leading doubles and
 * aks_get_records_dd the full pairs of the last step: out[nprob][7][2] = stop, Etot, Ekin, Ecou, Ehar, Eexc, t.
 * aks_set_density and the step-level hooks return AKS_ENOSYS for such batches.                              */
int aks_get_records_dd(aks_batch* batch, double* out);
/* test hook of the Double64 functionals (evaluate_vrho! / evaluate_zk! with T = Double64, models.jl:97-116):
 * rho, vrho, zk are npts (hi, lo) pairs; densities are clamped at 0 as BuiltinFunctional.jl:112 does       */
int aks_dd_xc_eval(aks_ctx* ctx, int32_t npts, const double* rho, int32_t ex_id, int32_t ec_id, double* vrho,
                   double* zk);
/* Double64 tables built by the library, on the device (no reference counterpart as an interface: replaces what
 * GaussLegendre{Double64}(basis, n) -- QuadGK.gauss(T, n), src/fem/integration methods.jl:50 -- and init_cache!
 * (src/discretization/discretization.jl:320-341 with the local matrices of src/fem/local matrix.jl:5-123 and the mass
 * tensor of src/fem/matrices.jl:297-343) compute on the host for T = Double64).  Outputs are (hi, lo) pairs in
 * caller-owned buffers, in the layouts aks_disc_create (dtype = 1) and the host mirror use:
 *   aks_dd_gauss_legendre : x, w [n][2], nodes ascending on [-1, 1]
 *   aks_dd_build_operators: cell-local blocks in generator order (0: 1+x, 1: 1-x, k+1: integrated Legendre k),
 *       M0, A, Mm1, Mm2 [C][p+1][p+1][2] and the mass tensor F [C][p+1][p+1][p+1][2]; present [C][p+1] marks the
 *       generators a cell has (Dirichlet ends); nq = Gauss nodes per cell of the build (3p + 80 reach 1e-34)       */
int aks_dd_gauss_legendre(aks_ctx* ctx, int32_t n, double* x, double* w);
int aks_dd_build_operators(aks_ctx* ctx, int32_t p, int32_t C, const double* mesh, const int32_t* present, int32_t nq,
                           double* M0_loc, double* A_loc, double* Mm1_loc, double* Mm2_loc, double* F_loc);
/* eps, n, W of every problem at once: eps, n [nprob][(L, nh[, 2])], W [nprob][Nh]; any pointer may be NULL */
int aks_get_state_all(aks_batch* batch, double* eps, double* n, double* W);
/* overwrite one problem's density state from a dense D (Nh,Nh[,2]) -- restart / step-level
 * parity tests ("identical input D").  Energies of the previous step are set from
 * `energies_prev` = {Etot,Ekin,Ecou,Ehar,Eexc} and niter from `niter`.                    */
int aks_set_density(aks_batch* batch, int32_t prob, const double* D, const double* energies_prev,
                    int32_t niter);

/* Post-processing on the device (src/solution/postprocess.jl:34-387): the solution of problem `prob`
 * evaluated at npts radial points x (host array) straight from the device state, written to out (host array).
 *   kind 0  u_{k,l,sigma}(x)/x   eval_orbital (l, k = radial index 0.., sigma = 0 / 1)       postprocess.jl:34-62
 *   kind 5  u_{k,l,sigma}(x)     the H^1_0 function itself
 *   kind 1  rho_sigma(x)         eval_density (sigma = -1: spin-summed)                      postprocess.jl:81-182
 *   kind 2  V_H(x) = W(x)/x + N/Rmax                eval_hartree                             postprocess.jl:207-222
 *   kind 3  v_xc,sigma(x)        eval_vxc (builtin LDA functionals of the problem)           postprocess.jl:328-352
 *   kind 4  -Z/x + l(l+1)/(2x^2) + hartree V_H + v_xc   eval_effective_potential             postprocess.jl:378-387
 * Points outside [0, Rmax] give the value of a vanishing solution.  Float64 batches only (AKS_ENOSYS for Double64). */
int aks_eval_grid(aks_batch* batch, int32_t prob, int32_t kind, int32_t l, int32_t k, int32_t sigma, int64_t npts,
                  const double* x, double* out);

/* ---- step-level hooks (same device code as aks_scf_step; used by the parity tests) ------ */
/* assemble_hartree!  (assemble.jl:63-83):   Har (Nh,Nh) of the problem's current D        */
int aks_assemble_hartree(aks_batch* batch, int32_t prob, double* Har);
/* assemble_exc!      (assemble.jl:277-319): Vxc (Nh,Nh[,2]) of the current D              */
int aks_assemble_vxc(aks_batch* batch, int32_t prob, double* Vxc);
/* find_orbital!      (oda/loop.jl:65-77):   lowest nh eigenpairs of every (l,sigma) channel
 * of H[D]; eps (L,nh[,2]), U (Nh,nh,L[,2]); residuals (L,nh[,2]) = ||(H-eps M0)u||_2 or NULL */
int aks_eig_lowest(aks_batch* batch, int32_t prob, double* eps, double* U, double* residuals);
/* density!           (density.jl:15-41):    D (Nh,Nh[,2]) = sum n u u^T of the current orbitals and
 * occupations (the trial density of the last step, before mixing)                         */
int aks_density(aks_batch* batch, int32_t prob, double* D);
/* compute_exc_energy (energies.jl:217-233) of the current D                               */
int aks_exc_energy(aks_batch* batch, int32_t prob, double* Exc);
/* compute_hartree_energy (energies.jl:96-114) of the current D                            */
int aks_hartree_energy(aks_batch* batch, int32_t prob, double* Ehar);

/* eigensolver diagnostics of the last step: info[nprob*nspin*L*nh], device order [prob][sigma][l][k];
 * each word = (inertia counts << 8) | (0x80: not converged + certified -> the problem ends with AKS_EEIG) |
 * (0x40: certified as a member of a cluster of levels closer than the certificate resolves) | RQI iterations  */
int aks_debug_eig_info(aks_batch* batch, int32_t* info);
/* counts[nprob]: steps in which the eigenpair window of a problem had to be refilled (see
 * aks_batch_set_eig_window) since the batch was created                                     */
int aks_debug_eig_refills(aks_batch* batch, int32_t* counts);
/* per-eigenpair cycle counters [nprob*nspin*L*nh][8] = {setup, factor, solve, apply+dots, final, total,
 * 0, 0}; zeros unless the library was built with -DAKS_EIG_TIMING                              */
int aks_debug_eig_cycles(aks_batch* batch, int64_t* cycles);

/* ---- measurement aids --------------------------------------------------------------------*/
/* number of kernels launched by this ctx so far (bench.py "gpu_launches")                 */
int64_t aks_launch_count(const aks_ctx* ctx);
/* sustained FP64 FMA throughput of the device in TFLOP/s (register-resident DFMA chains);
 * the FP64 roofline denominator, which MEASURED_PEAKS.json does not carry                 */
int aks_measure_fp64_peak(aks_ctx* ctx, double* tflops);
/* the same for the FP64 tensor-core path (mma.sync m8n8k4 f64)                                */
int aks_measure_dmma_peak(aks_ctx* ctx, double* tflops);
/* per-phase device time of the LAST aks_scf_step in ms: {potential, eigensolve, aufbau,
 * density, poisson, linesearch, mix, k_mix_dense alone, k_potential alone, k_xc alone}; valid only if
 * AKS_PHASE_TIMERS=1 in the environment (the batch then runs on one stream instead of its lanes) */
int aks_last_phase_ms(aks_batch* batch, double out[10]);
/* device time of EVERY kernel of the last aks_scf_step, in launch order (same AKS_PHASE_TIMERS mode):
 * ms[i] and the NUL-terminated name of kernel i at names + i * name_stride; *count = kernels in the step     */
int aks_last_kernel_ms(aks_batch* batch, int32_t cap, double* ms, char* names, int32_t name_stride, int32_t* count);

#ifdef __cplusplus
}
#endif
#endif /* AKS_B200_H */
