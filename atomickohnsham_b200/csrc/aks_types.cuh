// Device-side views of a discretization and of a batch of problems.
//
// Data layout in HBM (DESIGN.md "layout"): every operator is kept UNASSEMBLED, one
// (p+1)x(p+1) block per cell in generator order (g = 0: right hat 1+x, g = 1: left hat
// 1-x, g >= 2: bubbles), because every step of the hot path is cell-local.  Global
// vectors (orbitals U, Poisson rhs/solution) use the reference numbering
// (src/fem/generators.jl:79-138) and are reached through `l2g`.
#pragma once
#include <stdint.h>

#include "../../include/aks_b200.h"

struct DiscDev {
  int p, n, nb, npk, C, Nh, Q, Qs, G, L, nh, nspin;
  double Rmax;
  const double* Pg;      // [n][Qs]   generator g at Gauss node q (row stride Qs)
  const double* Pgc;     // [ceil(Q/64)][25][68]  the same table as shared-memory images of 64-node chunks (n <= 24 only): rows
                         //   n..23, the 4 pad columns and the nodes >= Q are zero, row 24 holds the nodes x_q themselves -- one
                         //   bulk copy per chunk (k_potential_tc: 24 rows, k_dens_cells_tc: 25)
  const double* w;       // [Q]
  const double* x;       // [Q]   Gauss nodes on [-1,1]
  const double* inv1;    // [C]  h/2
  const double* inv2;    // [C]  midpoint
  const double* A_loc;   // [C][n][n]
  const double* M0_loc;  // [C][n][n]
  const double* Mm1_loc; // [C][n][n]
  const double* Mm2_loc; // [C][n][n]
  const double* M0pk;    // [C][npk]  packed (hh | hb | bb lower), see pk_index()
  const double* F_loc;   // [C][n (m)][n*n (ij)], symmetric in ij
  const int* l2g;        // [C][n]   global index or -1
  const int* lo_tab;     // [n*n]    idx -> index of its lower-triangle twin
  const int* pk_tab;     // [n*n]    idx -> packed index (pk_index) or -1 above the diagonal
  const double* y;       // [G]
  const double* wy;      // [G]
  const int* ycell;      // [G]   0-based cell
  const double* Pgy;     // [n][G]
  // prefactored radial Poisson operator (stiffness A): static condensation
  const double* Abb_inv; // [C][nb][nb]  inverse of the bubble block
  const double* XA;      // [C][2][nb]   A_bb^{-1} A_bh
  const double* triA_d;  // [C-1] LDL^T of the hat Schur complement
  const double* triA_l;  // [C-1]
  const double* Linv;    // [C][nb][nb] inverse Cholesky factor of the bubble mass block (lower, row-major)
};

#define AKS_KHIST 4             // steps over which the eigenpair window remembers occupied levels
#define AKS_EPS_MISSING 1e300   // eps_new of a level outside the eigenpair window (k_eig)

struct BatchDev {
  int B, s, Lmax, nhmax;
  // model parameters
  const double *Z, *N, *hartree;
  const int *ex, *ec, *Lp, *nhp;
  // ODA / aufbau options
  double scftol, tinit, abstol_ls, reltol_ls, degen_tol;
  int frozen_t, maxiter_ls, max_degen, handle_spin;
  // per-problem ODA options (aks_batch_set_oda fills them for every problem, aks_batch_set_oda_problem for one:
  // the reference chooses algorithm and maxiter per atom, examples/atomic density comparison/run.jl:74-76)
  double* scftol_p;  // [B]
  int* frozen_p;     // [B]
  int* maxiter_p;    // [B]  cap on the problem's niter enforced on the device (status MAXITERS); 0 = none
  int* eig_fail;     // [B]  an eigenpair of the current step could not be certified (k_eig) -> status AKS_EEIG
  // device-side history of aks_scf_run: record of the step that took problem p from niter0[p] + k to
  // niter0[p] + k + 1 at hist[k * hist_stride + p]
  aks_iter_record* hist;
  int* niter0;       // [B]
  int hist_cap, hist_stride;
  int aufbau_kind;   // 0 OptimizedAufbau, 1 SmearedAufbau (temp), 2 FrozenAufbau (nfix)
  double temp;       // smearing width in Hartree (aufbau/smeared.jl)
  double* nfix;      // [B][s][Lmax][nhmax] prescribed occupations (aufbau/frozen.jl)
  // persistent state (mixed density, in every representation the path needs)
  double* Dd;     // [B][s][Nh][Nh]  dense density matrix (stopping norm, output)
  double* rho_q;  // [B][s][C][Q]    rho at the Gauss nodes of every cell
  double* rho_y;  // [B][s][G]       rho on the global energy grid
  double* Bv;     // [B][Nh]         Hartree rhs  F:D
  double* Wv;     // [B][Nh]         A^{-1} B
  double* E;      // [B][5]          Etot,Ekin,Ecou,Ehar,Eexc
  double* t;      // [B]
  double* stop;   // [B]
  int* niter;     // [B]
  int* status;    // [B]
  // per-iteration work
  double* fxc;       // [B][s][C][Q]   w_q * vxc_sigma(rho(q))
  double* Kxc;       // [B][s][C][n*n] XC block of every cell (k_potential_tc -> k_hassemble)
  double* Hpk;       // [B][s][Lmax][C][npk]  packed lower blocks (k_reduce)
  double* Hfull;     // [B][s][Lmax][C][n*n]  the same blocks, full symmetric (mat-vecs in k_eig)
  double* eps;       // [B][s][Lmax][nhmax]
  double* eps_new;   // [B][s][Lmax][nhmax]  written by k_eig, committed by k_aufbau
  double* U;         // [B][s][Lmax][nhmax][Nh]
  int* eig_info;     // [B][s][Lmax][nhmax]  (counts<<8 | rqi its) diagnostics
  long long* eig_dbg; // [B][s][Lmax][nhmax][8] cycle counters (filled only with -DAKS_EIG_TIMING)
  double* red;       // [B][s][Lmax][C][red_stride]  per-cell tridiagonal reduction of the pencil
  double* Pm;        // [B][s][Lmax][C][nb*nb]  P = Q^T L^-1
  double* Pt;        // [B][s][Lmax][C][nb*nb]  P^T
  double* eig_part;  // [B][s][Lmax][nhmax][nchunk][4]  per-chunk x^T M x, x^T H x, sign sum
  double* Xt;        // [B][s][Lmax][nhmax][nb*C + C]  eigenvectors in reduced coordinates (t | h)
  int* warm;         // [B]  1 once eps/U hold a previous solution
  int* kact;         // [B][s][Lmax]  eigenpairs per channel solved in the first pass of the next step
  int* kdone;        // [B][s][Lmax]  eigenpairs per channel that are current (solved for the last H)
  int* redo;         // [B]  aufbau could not be certified on the window: solve the missing levels, repeat
  int* kreq;         // [B][s][Lmax]  levels per channel the refill pass has to reach (k_eig_cert)
  int* full_refill;  // [B]  the aufbau ran out of solved levels: refill every channel completely
  int* cert_cnt;     // [B][s][Lmax]  #{eigenvalues < e_cert} of the channel (k_eig_cert)
  int* redo_cnt;     // [B]  how often that happened (diagnostic)
  double* e_cert;    // [B]  shift of the window certificate: last level the aufbau examined + degen_tol
  int* degen3;       // [B]  > 2-fold degeneracy met by the (not yet certified) aufbau
  int* khist;        // [B][s][Lmax][AKS_KHIST]  occupied levels per channel of the last steps
  int eig_margin;    // window = occupied levels + margin per channel; < 0: always all nh
  double* ekin_orb;  // [B][s][Lmax][nhmax]
  double* ecou_orb;  // [B][s][Lmax][nhmax]
  double* occ;       // [B][2][s][Lmax][nhmax]  slot 0: final / first extreme, slot 1: second extreme
  int* degen;        // [B]  1 when the two-fold branch is active this iteration
  int* occ_cnt;      // [B][2]            number of occupied orbitals per slot
  int* occ_elk;      // [B][2][MAXORB]    packed (e<<16 | l<<8 | k), density! order (sigma, k, l)
  double* occ_n;     // [B][2][MAXORB]    their occupations
  int* degen_idx;    // [B][2]  orbital offsets (into [s][Lmax][nhmax]) of the block
  double* degen_n;   // [B][4]  n1_0, n2_0, n1_1, n2_1
  double* rho_q_new; // [B][2][s][C][Q]
  double* rho_y_new; // [B][2][s][G]
  double* corr_part; // [B][2][C]
  double* Bloc_new;  // [B][2][C][n]
  double* B_new;     // [B][2][Nh]
  double* W_new;     // [B][2][Nh]
  double* scal;      // [B][2][4]   B.W self, Bprev.W, -, -
  double* td;        // [B]  weight of slot 0 in the trial density (1 if not degenerate)
  double* tmix;      // [B]  ODA t of this step
  double* norm_part; // [B][s][ntri]
  int ntile;         // tiles per side of D (ceil(Nh/MIX_TS))
  int ntri;          // ntile (ntile + 1) / 2: tiles on and below the diagonal
  aks_iter_record* rec; // [B]
  double* Enew;      // [B][5] energies of the trial density (before mixing)
};

// packed index of local pair (i >= j), generator order
__host__ __device__ inline int pk_index(int i, int j, int nb) {
  if (i < 2) return (i == 0) ? 0 : (j == 0 ? 1 : 2);
  const int a = i - 2;
  if (j < 2) return 3 + j * nb + a;
  const int bj = j - 2;
  return 3 + 2 * nb + a * (a + 1) / 2 + bj;
}
