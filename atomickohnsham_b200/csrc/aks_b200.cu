// Host side of the C ABI (include/aks_b200.h): table preprocessing, device memory, the
// per-iteration launch sequence.  No CPU fallback: every numerical step of the SCF loop runs
// in the kernels of k_*.cuh; the host only splits the caller's (reference-built) global
// tables into per-cell blocks once, at aks_disc_create.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/aks_b200.h"
#include "aks_device.cuh"
#include "aks_types.cuh"
#include "k_aux.cuh"
#include "k_eig.cuh"
#include "k_potential.cuh"
#include "k_scf.cuh"
#include "k_dd.cuh"
#include "k_dd_tables.cuh"

struct DDDiscHost;
struct DDBatchHost;

struct aks_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  int64_t launches = 0;
  // pinned record buffers and lane streams / events of destroyed batches, reused by the next batch:
  // cudaFreeHost and stream destruction synchronise the whole device (tens of ms), batches come and go
  std::vector<std::pair<void*, size_t>> pinned_free;
  std::vector<cudaStream_t> stream_free;
  std::vector<cudaEvent_t> event_free;
  int check_every = 4;    // aks_scf_run reads the status words every this many steps (AKS_CHECK_EVERY)
  bool mix_pipe = false;  // AKS_MIX_PIPE=1: dense mixing by persistent CTAs with a cp.async double buffer (measured slower)
  int sm_count = 148;
  bool dens_tc = true;    // cell densities by k_dens_cells_tc + k_fd (FP64 tensor cores; orders p <= 23) -- AKS_DENS_TC=0: k_dens_cells
  bool pot_tc = true;     // potential blocks by k_potential_tc (FP64 tensor cores; orders p <= 23) -- AKS_POTENTIAL_TC=0: k_potential
  bool use_graphs = true; // the step sequence of a lane is captured once and replayed as a CUDA graph (AKS_NO_GRAPH=1: off)
};

// Every entry point runs with the ctx's device current and restores the caller's device on return: several
// contexts (one per GPU, each driven by its own host thread) and other CUDA users (torch) share the process.
struct DevGuard {
  int prev = -1;
  bool switched = false;
  explicit DevGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = (cudaSetDevice(dev) == cudaSuccess);
  }
  ~DevGuard() { if (switched) cudaSetDevice(prev); }
  DevGuard(const DevGuard&) = delete;
  DevGuard& operator=(const DevGuard&) = delete;
};
#define GUARD(ctxp) DevGuard dev_guard_((ctxp)->device)
#define AKS_GRAPH_MAX_BATCH 32

struct aks_disc {
  aks_ctx* ctx = nullptr;
  DiscDev dev{};
  std::vector<void*> allocs;
  std::vector<int> l2g;  // host copy [C][n]
  int nmax_tpl = 32;     // template NMAX for the density kernels
  DDDiscHost* dd = nullptr;   // dtype = 1 (Double64): the double-double tables (aks_dd.cuh)
};

// A lane = a contiguous range of problems with its own stream.  Problems are independent, so the lanes of a
// batch run the step sequence concurrently: small-grid and latency-bound kernels of one lane (line search,
// aufbau, refill pass, kernel tails) overlap with the wide kernels of the other.
struct Lane {
  BatchDev dev{};   // view of the batch: every per-problem pointer advanced to the lane's first problem
  int B = 0, p0 = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  cudaGraphExec_t gexec = nullptr;   // the lane's step sequence, captured once per set of options
  int gkernels = 0;                  // kernels per replay
};
struct FieldInfo { size_t offset, bytes_per_problem; };

struct aks_batch {
  aks_disc* disc = nullptr;
  BatchDev dev{};
  std::vector<void*> allocs;
  std::vector<FieldInfo> fields;   // per-problem device arrays of BatchDev (for the lane views)
  std::vector<Lane> lanes;         // >= 2 entries when the step is split over streams
  Lane whole;                      // all problems on the ctx stream
  Lane* cur = nullptr;             // what the launch helpers act on
  cudaEvent_t fork_ev = nullptr;
  cudaEvent_t poll_ev[2] = {nullptr, nullptr};   // aks_scf_run: status read-back of a chunk of steps
  bool forked = false;             // lanes are ahead of the ctx stream only by their own steps
  int B = 0;
  bool fac_in_smem = true;
  double* fac_global = nullptr;
  bool pot_tc = false;
  size_t smem_pot = 0, smem_eig = 0, smem_eig_main = 0, smem_red = 0, smem_chan = 0, smem_dens = 0, smem_poisson = 0;
  std::vector<double> hZ, hN, hHar, hScftol, hTinit;
  std::vector<int> hLp, hnhp, hFrozen, hMaxiter;
  bool graphs_valid = false;         // lanes' graphs match bt->dev (options, pointers)
  size_t hist_cap_alloc = 0;         // records per problem the device history can hold
  aks_iter_record* h_rec = nullptr;  // pinned
  size_t h_rec_bytes = 0;
  bool timers = false;
  bool eig_dirty = false;   // levels outside the eigenpair window are stale (k_eig)
  cudaEvent_t ev[14] = {};   // 8 phase boundaries + [9],[10] around k_mix_dense + [11],[12],[13] around k_xc, k_potential
  double phase_ms[10] = {0};
  // per-kernel device times of the last step (AKS_PHASE_TIMERS): an event after every launch
  std::vector<cudaEvent_t> kev;
  std::vector<const char*> knames;
  std::vector<double> kms;
  size_t kcount = 0;
  DDBatchHost* dd = nullptr;   // batch on a Double64 discretization (aks_dd.cuh)
};

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      char buf_[512];                                                                         \
      snprintf(buf_, sizeof buf_, "%s:%d %s -> %s", __FILE__, __LINE__, #call,                \
               cudaGetErrorString(e_));                                                       \
      ctx->err = buf_;                                                                        \
      return AKS_ECUDA;                                                                       \
    }                                                                                         \
  } while (0)

template <class T>
static int dev_upload(aks_ctx* ctx, std::vector<void*>& allocs, const std::vector<T>& h, const T** out) {
  void* p = nullptr;
  const size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
  CK(cudaMallocAsync(&p, bytes, ctx->stream));   // stream-ordered pool: see aks_ctx_create
  allocs.push_back(p);
  if (!h.empty()) CK(cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *out = (const T*)p;
  return AKS_OK;
}
template <class T>
static int dev_zeros(aks_ctx* ctx, std::vector<void*>& allocs, size_t count, T** out) {
  void* p = nullptr;
  const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
  CK(cudaMallocAsync(&p, bytes, ctx->stream));
  allocs.push_back(p);
  CK(cudaMemsetAsync(p, 0, bytes, ctx->stream));
  *out = (T*)p;
  return AKS_OK;
}

// Double64 path (aks_dd.cuh, included at the end of this file)
static int dd_disc_create(aks_ctx* ctx, const aks_disc_desc* D, aks_disc** out);
static void dd_disc_free(aks_disc* disc);
static int dd_batch_create(aks_disc* disc, int32_t nprob, const double* Z, const double* N, const double* hartree,
                           const int32_t* ex_id, const int32_t* ec_id, const int32_t* lh_p, const int32_t* nh_p,
                           aks_batch** out);
static void dd_batch_free(aks_batch* bt);
static int dd_batch_reset(aks_batch* bt);
static int dd_set_oda(aks_batch* bt, double scftol, double tinit, int frozen_t, int max_degen, double degen_tol,
                      int maxiter_ls, double abstol_ls, double reltol_ls);
static int dd_step_launch(aks_batch* bt);
static int dd_eig_cold(aks_batch* bt, int32_t* counts);
static int dd_set_window(aks_batch* bt, int32_t margin);
static int dd_set_aufbau(aks_batch* bt, int32_t kind, double temp, const double* nfix);
static int dd_get_state(aks_batch* bt, int prob, double* D, double* U, double* eps, double* nocc, double* W);
#define DD_UNSUPPORTED(bt, what)                                                               \
  do {                                                                                         \
    if ((bt)->dd) { (bt)->disc->ctx->err = what " is not built for Double64 batches"; return AKS_ENOSYS; } \
  } while (0)

// ------------------------------------------------------------------------------------ context
// Dynamic shared memory limits are per function and device, i.e. process-global state: they are raised ONCE per
// ctx to the device's opt-in maximum, never to the sizes of one batch (a later, smaller batch would otherwise
// lower the limit under a live larger one).  Launches still request only what they need.
static int set_kernel_attributes(aks_ctx* ctx) {
  int optin = 0;
  CK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
  // dynamic + static shared memory of a kernel must stay within the opt-in maximum
#define SMEM_MAX(fn)                                                                                          \
  do {                                                                                                        \
    cudaFuncAttributes fa_;                                                                                   \
    CK(cudaFuncGetAttributes(&fa_, fn));                                                                      \
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa_.sharedSizeBytes)); \
  } while (0)
  SMEM_MAX((k_potential<1, 4>)); SMEM_MAX((k_potential<2, 2>)); SMEM_MAX(k_potential_tc); SMEM_MAX(k_hassemble); SMEM_MAX(k_mix_dense_pipe); SMEM_MAX(k_dens_cells_tc<1>); SMEM_MAX(k_dens_cells_tc<2>); SMEM_MAX(k_fd);
  SMEM_MAX(k_eig); SMEM_MAX(k_eig_cert);
  SMEM_MAX(k_reduce<8>); SMEM_MAX(k_reduce<12>); SMEM_MAX(k_reduce<20>); SMEM_MAX(k_reduce<28>);
  SMEM_MAX(k_eig_init); SMEM_MAX(k_eig_final); SMEM_MAX(k_poisson);
  SMEM_MAX(k_dens_cells<12>); SMEM_MAX(k_dens_cells<24>); SMEM_MAX(k_dens_cells<32>);
  SMEM_MAX(k_cells_from_dense<12>); SMEM_MAX(k_cells_from_dense<24>); SMEM_MAX(k_cells_from_dense<32>);
  SMEM_MAX(k_eig_residual);
  SMEM_MAX(ddk_eig<DD_EIG_LN>);
#undef SMEM_MAX
  return AKS_OK;
}

extern "C" int aks_ctx_create(int device, aks_ctx** out) {
  if (!out) return AKS_EINVAL;
  aks_ctx* ctx = new aks_ctx();
  ctx->device = device;
  *out = ctx;
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) { ctx->err = "no such CUDA device"; return AKS_ECUDA; }
  GUARD(ctx);
  CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  { int rc = set_kernel_attributes(ctx); if (rc != AKS_OK) return rc; }
  if (const char* e = getenv("AKS_CHECK_EVERY")) ctx->check_every = std::max(1, atoi(e));
  if (const char* e = getenv("AKS_NO_GRAPH")) ctx->use_graphs = !(e[0] == '1');
  if (const char* e = getenv("AKS_POTENTIAL_TC")) ctx->pot_tc = (e[0] != '0');
  if (const char* e = getenv("AKS_DENS_TC")) ctx->dens_tc = (e[0] != '0');
  if (const char* e = getenv("AKS_MIX_PIPE")) ctx->mix_pipe = (e[0] != '0');
  { int v = 0; if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && v > 0) ctx->sm_count = v; }
  {
    // batches come and go (one per sweep / restart): their device arrays are taken from the device's
    // stream-ordered pool, which keeps freed memory instead of returning it to the OS
    cudaMemPool_t pool;
    CK(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t keep = UINT64_MAX;
    CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  }
  XcConsts xc;
  xc.cx = std::pow(3.0 / M_PI, 1.0 / 3.0);
  xc.cx34 = 0.75 * xc.cx;
  xc.cx6_34 = 0.75 * std::pow(6.0 / M_PI, 1.0 / 3.0);
  xc.fourpi = 4.0 * M_PI;
  xc.inv4pi = 1.0 / (4.0 * M_PI);
  xc.f2p0 = 4.0 / (9.0 * (std::pow(2.0, 1.0 / 3.0) - 1.0));
  xc.two43m2 = std::pow(2.0, 4.0 / 3.0) - 2.0;
  xc.third = 1.0 / 3.0;
  xc.crs = std::cbrt(3.0 / (4.0 * M_PI));
  xc.icrs = 1.0 / xc.crs;
  xc.four3 = 4.0 / 3.0;
  CK(cudaMemcpyToSymbol(c_xc, &xc, sizeof xc));
  return AKS_OK;
}
extern "C" void aks_ctx_destroy(aks_ctx* ctx) {
  if (!ctx) return;
  GUARD(ctx);
  for (auto& pf : ctx->pinned_free) cudaFreeHost(pf.first);
  for (cudaStream_t st : ctx->stream_free) cudaStreamDestroy(st);
  for (cudaEvent_t ev : ctx->event_free) cudaEventDestroy(ev);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}
extern "C" const char* aks_last_error(const aks_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }
extern "C" void* aks_ctx_stream(const aks_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" int64_t aks_launch_count(const aks_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" int aks_ctx_configure(aks_ctx* ctx, int32_t use_graphs, int32_t check_every) {
  if (!ctx) return AKS_EINVAL;
  if (use_graphs >= 0) ctx->use_graphs = use_graphs != 0;   // batches capture lazily: takes effect at their next step
  if (check_every >= 1) ctx->check_every = check_every;
  return AKS_OK;
}

// ------------------------------------------------------------------------------------ disc
// small dense helpers (setup only)
static bool chol_inverse(std::vector<double>& A, int m) {
  // in-place inverse of an SPD m x m matrix (row-major) via Gauss-Jordan with no pivoting
  std::vector<double> I(m * m, 0.0);
  for (int i = 0; i < m; ++i) I[i * m + i] = 1.0;
  for (int j = 0; j < m; ++j) {
    const double piv = A[j * m + j];
    if (!(piv > 0.0)) return false;
    const double inv = 1.0 / piv;
    for (int k = 0; k < m; ++k) { A[j * m + k] *= inv; I[j * m + k] *= inv; }
    for (int i = 0; i < m; ++i) {
      if (i == j) continue;
      const double f = A[i * m + j];
      if (f == 0.0) continue;
      for (int k = 0; k < m; ++k) { A[i * m + k] -= f * A[j * m + k]; I[i * m + k] -= f * I[j * m + k]; }
    }
  }
  A = I;
  return true;
}

extern "C" int aks_disc_create(aks_ctx* ctx, const aks_disc_desc* D, aks_disc** out) {
  if (!ctx || !D || !out) return AKS_EINVAL;
  *out = nullptr;
  GUARD(ctx);
  if (D->dtype == 1) return dd_disc_create(ctx, D, out);
  if (D->dtype != 0) { ctx->err = "bad dtype"; return AKS_EINVAL; }
  const int p = D->p, C = D->C, Nh = D->Nh, Q = D->Q, G = D->G, n = p + 1, nb = p - 1;
  if (p < 2 || p > 29 || C < 2 || Nh != p * C - 1 || Q < 1 || G < 1 || D->lh < 0 || D->nh < 1 ||
      (D->nspin != 1 && D->nspin != 2)) {
    ctx->err = "aks_disc_create: need 2 <= p <= 29, C >= 2, Nh == p*C-1, Q,G >= 1, nspin in {1,2}";
    return AKS_EINVAL;
  }
  if ((D->lh + 1) * D->nh * D->nspin > AKS_MAXORB) { ctx->err = "(lh+1)*nh*nspin exceeds AKS_MAXORB"; return AKS_EINVAL; }
  if (!D->mesh || !D->invshifts || !D->x || !D->w || !D->Pgenx || !D->y || !D->wy || !D->ycell ||
      !D->Pgeny || !D->M0 || !D->A || !D->Mm1 || !D->cell_len || !D->cell_indices || !D->cell_generators) {
    ctx->err = "aks_disc_create: NULL table";
    return AKS_EINVAL;
  }
  if (D->lh > 0 && !D->Mm2) { ctx->err = "Mm2 required when lh > 0"; return AKS_EINVAL; }
  aks_disc* disc = new aks_disc();
  disc->ctx = ctx;
  DiscDev& dv = disc->dev;
  dv.p = p; dv.n = n; dv.nb = nb; dv.npk = n * (n + 1) / 2; dv.C = C; dv.Nh = Nh; dv.Q = Q;
  dv.Qs = ((Q + 3) / 4) * 4; dv.G = G; dv.L = D->lh + 1; dv.nh = D->nh; dv.nspin = D->nspin;
  dv.Rmax = D->mesh[C];
  disc->nmax_tpl = (n <= 12) ? 12 : (n <= 24 ? 24 : 32);

  // local-to-global map
  std::vector<int>& l2g = disc->l2g;
  l2g.assign((size_t)C * n, -1);
  for (int c = 0; c < C; ++c) {
    const int len = (int)D->cell_len[c];
    if (len < 1 || len > n) { ctx->err = "bad cell_len"; delete disc; return AKS_EINVAL; }
    for (int j = 0; j < len; ++j) {
      const int64_t I = D->cell_indices[(size_t)c * n + j], g = D->cell_generators[(size_t)c * n + j];
      if (I < 1 || I > Nh || g < 1 || g > n) { ctx->err = "bad cell index/generator"; delete disc; return AKS_EINVAL; }
      l2g[(size_t)c * n + (g - 1)] = (int)(I - 1);
    }
  }
  // structural check: right hat of cell c == left hat of cell c+1, ends are Dirichlet
  for (int c = 0; c + 1 < C; ++c)
    if (l2g[(size_t)c * n] < 0 || l2g[(size_t)c * n] != l2g[(size_t)(c + 1) * n + 1]) {
      ctx->err = "cell maps are not a P1+bubble chain"; delete disc; return AKS_EINVAL;
    }
  if (l2g[1] != -1 || l2g[(size_t)(C - 1) * n] != -1) { ctx->err = "expected Dirichlet ends"; delete disc; return AKS_EINVAL; }
  for (int c = 0; c < C; ++c)
    for (int i = 0; i < nb; ++i)
      if (l2g[(size_t)c * n + 2 + i] != l2g[(size_t)c * n + 2] + i) {
        ctx->err = "bubble indices of a cell must be contiguous"; delete disc; return AKS_EINVAL;
      }
  if (C > 8 * AKS_NW) { ctx->err = "at most 64 cells in this build"; delete disc; return AKS_EINVAL; }

  // split global symmetric matrices into cell-owned blocks (first cell containing both indices)
  auto split = [&](const double* Gm, std::vector<double>& loc) {
    loc.assign((size_t)C * n * n, 0.0);
    if (!Gm) return;
    std::vector<char> owned((size_t)Nh * Nh, 0);
    for (int c = 0; c < C; ++c)
      for (int a = 0; a < n; ++a) {
        const int I = l2g[(size_t)c * n + a];
        if (I < 0) continue;
        for (int bq = 0; bq < n; ++bq) {
          const int J = l2g[(size_t)c * n + bq];
          if (J < 0) continue;
          if (owned[(size_t)I * Nh + J]) continue;
          owned[(size_t)I * Nh + J] = 1;
          loc[((size_t)c * n + a) * n + bq] = 0.5 * (Gm[(size_t)I + (size_t)Nh * J] + Gm[(size_t)J + (size_t)Nh * I]);
        }
      }
  };
  std::vector<double> M0l, Al, M1l, M2l;
  split(D->M0, M0l);
  split(D->A, Al);
  split(D->Mm1, M1l);
  split(D->Mm2, M2l);
  // packed M0
  std::vector<double> M0pk((size_t)C * dv.npk, 0.0);
  for (int c = 0; c < C; ++c)
    for (int i = 0; i < n; ++i)
      for (int j = 0; j <= i; ++j) M0pk[(size_t)c * dv.npk + pk_index(i, j, nb)] = M0l[((size_t)c * n + i) * n + j];
  // mass tensor: COO -> per-cell [m][ij], owner = first cell that holds all three indices
  std::vector<double> Fl((size_t)C * n * n * n, 0.0);
  if (D->nF > 0) {
    if (!D->Fi || !D->Fj || !D->Fm || !D->Fv) { ctx->err = "NULL F arrays"; delete disc; return AKS_EINVAL; }
    std::vector<int> cell_of((size_t)Nh * 2, -1), loc_of((size_t)Nh * 2, -1);
    for (int c = 0; c < C; ++c)
      for (int a = 0; a < n; ++a) {
        const int I = l2g[(size_t)c * n + a];
        if (I < 0) continue;
        const int slot = (cell_of[2 * I] < 0) ? 0 : 1;
        cell_of[2 * I + slot] = c;
        loc_of[2 * I + slot] = a;
      }
    for (int64_t t = 0; t < D->nF; ++t) {
      const int I = (int)D->Fi[t] - 1, J = (int)D->Fj[t] - 1, M = (int)D->Fm[t] - 1;
      if (I < 0 || J < 0 || M < 0 || I >= Nh || J >= Nh || M >= Nh) { ctx->err = "bad F index"; delete disc; return AKS_EINVAL; }
      int owner = -1, la = 0, lb = 0, lm = 0;
      for (int si = 0; si < 2 && owner < 0; ++si) {
        const int c = cell_of[2 * I + si];
        if (c < 0) continue;
        int fj = -1, fm = -1;
        for (int sj = 0; sj < 2; ++sj) if (cell_of[2 * J + sj] == c) fj = loc_of[2 * J + sj];
        for (int sm2 = 0; sm2 < 2; ++sm2) if (cell_of[2 * M + sm2] == c) fm = loc_of[2 * M + sm2];
        if (fj >= 0 && fm >= 0) { owner = c; la = loc_of[2 * I + si]; lb = fj; lm = fm; }
      }
      if (owner < 0) continue;  // not a cell-local triple: cannot carry weight in this basis
      // symmetrise in (i,j): half to (a,b), half to (b,a)
      Fl[(((size_t)owner * n + lm) * n + la) * n + lb] += 0.5 * D->Fv[t];
      Fl[(((size_t)owner * n + lm) * n + lb) * n + la] += 0.5 * D->Fv[t];
    }
  }
  // Poisson prefactor: static condensation of the stiffness operator
  std::vector<double> Abbinv((size_t)C * nb * nb), XA((size_t)C * 2 * nb), triD(C, 1.0), triL(C, 0.0);
  {
    std::vector<double> S00(C), S10(C), S11(C);
    for (int c = 0; c < C; ++c) {
      std::vector<double> bbm((size_t)nb * nb);
      for (int i = 0; i < nb; ++i)
        for (int j = 0; j < nb; ++j) bbm[(size_t)i * nb + j] = Al[((size_t)c * n + 2 + i) * n + 2 + j];
      if (!chol_inverse(bbm, nb)) { ctx->err = "stiffness bubble block not positive definite"; delete disc; return AKS_EINVAL; }
      std::copy(bbm.begin(), bbm.end(), Abbinv.begin() + (size_t)c * nb * nb);
      for (int g = 0; g < 2; ++g)
        for (int i = 0; i < nb; ++i) {
          double acc = 0.0;
          for (int j = 0; j < nb; ++j) acc += bbm[(size_t)i * nb + j] * Al[((size_t)c * n + 2 + j) * n + g];
          XA[((size_t)c * 2 + g) * nb + i] = acc;
        }
      auto hb = [&](int g, int i) { return Al[((size_t)c * n + 2 + i) * n + g]; };
      double t00 = 0, t10 = 0, t11 = 0;
      for (int i = 0; i < nb; ++i) {
        t00 += hb(0, i) * XA[((size_t)c * 2 + 0) * nb + i];
        t10 += hb(1, i) * XA[((size_t)c * 2 + 0) * nb + i];
        t11 += hb(1, i) * XA[((size_t)c * 2 + 1) * nb + i];
      }
      S00[c] = Al[((size_t)c * n + 0) * n + 0] - t00;
      S10[c] = Al[((size_t)c * n + 1) * n + 0] - t10;
      S11[c] = Al[((size_t)c * n + 1) * n + 1] - t11;
    }
    double dprev = 0, offprev = 0;
    for (int j = 0; j < C - 1; ++j) {
      const double diag = S00[j] + S11[j + 1];
      double l = 0, dj = diag;
      if (j > 0) { l = offprev / dprev; dj = diag - l * offprev; }
      if (!(dj > 0.0)) { ctx->err = "stiffness Schur complement not positive definite"; delete disc; return AKS_EINVAL; }
      triD[j] = dj;
      if (j > 0) triL[j - 1] = l;
      dprev = dj;
      offprev = (j + 1 < C - 1) ? S10[j + 1] : 0.0;
    }
  }
  // inverse Cholesky factor of the bubble mass block of every cell (for k_reduce)
  std::vector<double> Linv((size_t)C * nb * nb, 0.0);
  for (int c = 0; c < C; ++c) {
    std::vector<double> Lm((size_t)nb * nb, 0.0);
    for (int i = 0; i < nb; ++i)
      for (int j = 0; j <= i; ++j) {
        double acc = M0l[((size_t)c * n + 2 + i) * n + 2 + j];
        for (int k = 0; k < j; ++k) acc -= Lm[(size_t)i * nb + k] * Lm[(size_t)j * nb + k];
        if (i == j) {
          if (!(acc > 0.0)) { ctx->err = "mass bubble block not positive definite"; delete disc; return AKS_EINVAL; }
          Lm[(size_t)i * nb + i] = std::sqrt(acc);
        } else Lm[(size_t)i * nb + j] = acc / Lm[(size_t)j * nb + j];
      }
    double* Lc = Linv.data() + (size_t)c * nb * nb;
    for (int j = 0; j < nb; ++j) {  // column j of the inverse by forward substitution
      for (int i = j; i < nb; ++i) {
        double acc = (i == j) ? 1.0 : 0.0;
        for (int k = j; k < i; ++k) acc -= Lm[(size_t)i * nb + k] * Lc[(size_t)k * nb + j];
        Lc[(size_t)i * nb + j] = acc / Lm[(size_t)i * nb + i];
      }
    }
  }
  // Gauss tables
  std::vector<double> Pg((size_t)n * dv.Qs, 0.0), Pgy((size_t)n * G, 0.0), inv1(C), inv2(C);
  for (int g = 0; g < n; ++g) {
    for (int q = 0; q < Q; ++q) Pg[(size_t)g * dv.Qs + q] = D->Pgenx[(size_t)g + (size_t)n * q];
    for (int q = 0; q < G; ++q) Pgy[(size_t)g * G + q] = D->Pgeny[(size_t)g + (size_t)n * q];
  }
  for (int c = 0; c < C; ++c) { inv1[c] = D->invshifts[2 * c]; inv2[c] = D->invshifts[2 * c + 1]; }
  std::vector<int> ycell(G);
  for (int q = 0; q < G; ++q) {
    const int64_t c = D->ycell[q];
    if (c < 1 || c > C) { ctx->err = "ycell out of range"; delete disc; return AKS_EINVAL; }
    ycell[q] = (int)c - 1;
  }
  std::vector<int> lo_tab((size_t)n * n), pk_tab((size_t)n * n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      lo_tab[(size_t)i * n + j] = (j > i) ? j * n + i : i * n + j;
      pk_tab[(size_t)i * n + j] = (j <= i) ? pk_index(i, j, nb) : -1;
    }
  std::vector<double> xv(D->x, D->x + Q), wv(D->w, D->w + Q), yv(D->y, D->y + G), wyv(D->wy, D->wy + G);
  int rc;
#define UP(vec, field)                                                                          \
  if ((rc = dev_upload(ctx, disc->allocs, vec, &dv.field)) != AKS_OK) { aks_disc_destroy(disc); return rc; }
  std::vector<double> Pgc;
  if (n <= POT_TC_ROWS) {
    const int nchk = (Q + POT_QC - 1) / POT_QC;
    Pgc.assign((size_t)nchk * POT_IMG_ROWS * POT_TC_QCS, 0.0);
    for (int q = 0; q < Q; ++q) {
      for (int g = 0; g < n; ++g)
        Pgc[((size_t)(q / POT_QC) * POT_IMG_ROWS + g) * POT_TC_QCS + q % POT_QC] = Pg[(size_t)g * dv.Qs + q];
      Pgc[((size_t)(q / POT_QC) * POT_IMG_ROWS + POT_TC_ROWS) * POT_TC_QCS + q % POT_QC] = D->x[q];
    }
  } else {
    Pgc.assign(8, 0.0);
  }
  UP(Pgc, Pgc)
  UP(Pg, Pg) UP(wv, w) UP(xv, x) UP(inv1, inv1) UP(inv2, inv2) UP(Al, A_loc) UP(M0l, M0_loc) UP(M1l, Mm1_loc)
  UP(M2l, Mm2_loc) UP(M0pk, M0pk) UP(Fl, F_loc) UP(l2g, l2g) UP(yv, y) UP(wyv, wy) UP(ycell, ycell) UP(Pgy, Pgy)
  UP(lo_tab, lo_tab) UP(pk_tab, pk_tab) UP(Abbinv, Abb_inv) UP(XA, XA) UP(triD, triA_d) UP(triL, triA_l) UP(Linv, Linv)
#undef UP
  *out = disc;
  return AKS_OK;
}

extern "C" void aks_disc_destroy(aks_disc* disc) {
  if (!disc) return;
  GUARD(disc->ctx);
  for (void* p : disc->allocs) cudaFreeAsync(p, disc->ctx->stream);
  cudaStreamSynchronize(disc->ctx->stream);
  dd_disc_free(disc);
  delete disc;
}

// ------------------------------------------------------------------------------------ batch
extern "C" int aks_batch_reset(aks_batch* bt);
static int window_full(aks_batch* bt, int p0, int np);
static int ensure_complete(aks_batch* bt);
static void refresh_lanes(aks_batch* bt);
static void main_side(aks_batch* bt);
static void enter_main(aks_batch* bt) { refresh_lanes(bt); main_side(bt); }

static int batch_create_impl(aks_batch* bt, aks_disc* disc, int32_t nprob, const double* Z, const double* N,
                             const double* hartree, const int32_t* ex_id, const int32_t* ec_id,
                             const int32_t* lh_p, const int32_t* nh_p);

extern "C" int aks_batch_create(aks_disc* disc, int32_t nprob, const double* Z, const double* N,
                                const double* hartree, const int32_t* ex_id, const int32_t* ec_id,
                                const int32_t* lh_p, const int32_t* nh_p, aks_batch** out) {
  if (!disc || !out || nprob < 1 || !Z || !N) return AKS_EINVAL;
  *out = nullptr;
  GUARD(disc->ctx);
  if (disc->dd) return dd_batch_create(disc, nprob, Z, N, hartree, ex_id, ec_id, lh_p, nh_p, out);
  aks_batch* bt = new aks_batch();
  bt->disc = disc;
  bt->B = nprob;
  const int rc = batch_create_impl(bt, disc, nprob, Z, N, hartree, ex_id, ec_id, lh_p, nh_p);
  if (rc != AKS_OK) { aks_batch_destroy(bt); return rc; }   // every error path releases what was allocated
  *out = bt;
  return AKS_OK;
}

static int batch_create_impl(aks_batch* bt, aks_disc* disc, int32_t nprob, const double* Z, const double* N,
                             const double* hartree, const int32_t* ex_id, const int32_t* ec_id,
                             const int32_t* lh_p, const int32_t* nh_p) {
  aks_ctx* ctx = disc->ctx;
  const DiscDev& dv = disc->dev;
  BatchDev& b = bt->dev;
  b.B = nprob; b.s = dv.nspin; b.Lmax = dv.L; b.nhmax = dv.nh;
  b.ntile = (dv.Nh + MIX_TS - 1) / MIX_TS;
  b.ntri = b.ntile * (b.ntile + 1) / 2;
  bt->hZ.assign(Z, Z + nprob);
  bt->hN.assign(N, N + nprob);
  bt->hHar.assign(nprob, 1.0);
  if (hartree) bt->hHar.assign(hartree, hartree + nprob);
  std::vector<int> ex(nprob, 0), ec(nprob, 0);
  bt->hLp.assign(nprob, dv.L);
  bt->hnhp.assign(nprob, dv.nh);
  for (int i = 0; i < nprob; ++i) {
    if (ex_id) ex[i] = ex_id[i];
    if (ec_id) ec[i] = ec_id[i];
    if (lh_p) bt->hLp[i] = lh_p[i] + 1;
    if (nh_p) bt->hnhp[i] = nh_p[i];
    if ((ex[i] != AKS_XC_NONE && ex[i] != AKS_LDA_X) || (ec[i] != AKS_XC_NONE && ec[i] != AKS_LDA_C_PW)) {
      ctx->err = "unsupported functional id (LDA only: ex in {0,1}, ec in {0,12})"; return AKS_EINVAL;
    }
    if (bt->hLp[i] < 1 || bt->hLp[i] > dv.L || bt->hnhp[i] < 1 || bt->hnhp[i] > dv.nh || !(N[i] > 0) || !(Z[i] > 0)) {
      ctx->err = "bad per-problem lh/nh/Z/N"; return AKS_EINVAL;
    }
  }
  int rc;
#define UPB(vec, field) if ((rc = dev_upload(ctx, bt->allocs, vec, &b.field)) != AKS_OK) return rc; \
  bt->fields.push_back({offsetof(BatchDev, field), sizeof(*b.field)});
  UPB(bt->hZ, Z) UPB(bt->hN, N) UPB(bt->hHar, hartree) UPB(ex, ex) UPB(ec, ec) UPB(bt->hLp, Lp) UPB(bt->hnhp, nhp)
#undef UPB
  const size_t B = nprob, s = b.s, Nh = dv.Nh, C = dv.C, Q = dv.Q, G = dv.G, L = b.Lmax, nh = b.nhmax, n = dv.n;
#define ZB(T, field, count) if ((rc = dev_zeros<T>(ctx, bt->allocs, (count), &b.field)) != AKS_OK) return rc; \
  bt->fields.push_back({offsetof(BatchDev, field), sizeof(T) * ((size_t)(count) / B)});
  ZB(double, Dd, B * s * Nh * Nh) ZB(double, rho_q, B * s * C * Q) ZB(double, rho_y, B * s * G)
  ZB(double, Bv, B * Nh) ZB(double, Wv, B * Nh) ZB(double, E, B * 5) ZB(double, t, B) ZB(double, stop, B)
  ZB(int, niter, B) ZB(int, status, B) ZB(double, fxc, B * s * C * Q) ZB(double, Kxc, B * s * C * n * n) ZB(double, Hpk, B * s * L * C * dv.npk) ZB(double, Hfull, B * s * L * C * n * n)
  ZB(double, eps, B * s * L * nh) ZB(double, eps_new, B * s * L * nh) ZB(double, U, B * s * L * nh * Nh)
  ZB(int, eig_info, B * s * L * nh) ZB(long long, eig_dbg, B * s * L * nh * 8) ZB(double, red, B * s * L * C * (6 * (n - 2) + 4))
  ZB(double, Pm, B * s * L * C * (n - 2) * (n - 2)) ZB(double, Pt, B * s * L * C * (n - 2) * (n - 2)) ZB(double, Xt, B * s * L * nh * ((n - 2) * C + C)) ZB(double, eig_part, B * s * L * nh * C * 6) ZB(int, warm, B) ZB(int, kact, B * s * L) ZB(int, kdone, B * s * L) ZB(int, redo, B) ZB(int, kreq, B * s * L) ZB(int, full_refill, B) ZB(int, cert_cnt, B * s * L) ZB(int, redo_cnt, B) ZB(double, e_cert, B) ZB(int, degen3, B) ZB(double, nfix, B * s * L * nh) ZB(int, khist, B * s * L * AKS_KHIST) ZB(double, ekin_orb, B * s * L * nh)
  ZB(double, ecou_orb, B * s * L * nh) ZB(double, occ, B * 2 * s * L * nh) ZB(int, degen, B) ZB(int, occ_cnt, B * 2) ZB(int, occ_elk, B * 2 * AKS_MAXORB) ZB(double, occ_n, B * 2 * AKS_MAXORB)
  ZB(int, degen_idx, B * 2) ZB(double, degen_n, B * 4) ZB(double, rho_q_new, B * 2 * s * C * Q)
  ZB(double, rho_y_new, B * 2 * s * G) ZB(double, corr_part, B * 2 * C) ZB(double, Bloc_new, B * 2 * C * n)
  ZB(double, B_new, B * 2 * Nh) ZB(double, W_new, B * 2 * Nh) ZB(double, scal, B * 2 * 4) ZB(double, td, B)
  ZB(double, tmix, B) ZB(double, norm_part, B * s * b.ntri) ZB(aks_iter_record, rec, B)
  ZB(double, Enew, B * 5) ZB(double, scftol_p, B) ZB(int, frozen_p, B) ZB(int, maxiter_p, B) ZB(int, eig_fail, B)
  ZB(int, niter0, B)
#undef ZB
  // device history of aks_scf_run: allocated by the first run that asks for it; a lane sees it from its first problem
  b.hist = nullptr; b.hist_cap = 0; b.hist_stride = nprob;
  bt->fields.push_back({offsetof(BatchDev, hist), sizeof(aks_iter_record)});
  // ODA defaults (oda/types.jl:38-44, optimized.jl:23-26)
  b.scftol = 1e-10; b.tinit = 0.6; b.frozen_t = 0; b.maxiter_ls = 100;
  b.abstol_ls = b.reltol_ls = 1e4 * 2.220446049250313e-16;
  b.max_degen = 2; b.degen_tol = 1e-3; b.handle_spin = 1;
  b.eig_margin = 0;
  b.aufbau_kind = 0; b.temp = 0.0;
  // launch configuration
  bt->pot_tc = ctx->pot_tc && potential_tc_ok(dv.n, dv.Q, dv.Qs, b.s);
  bt->smem_pot = bt->pot_tc ? potential_tc_smem_bytes() : potential_smem_bytes(dv.n, dv.Q, b.s);
  bt->smem_eig = eig_smem_bytes(dv.C, dv.nb, dv.Nh);
  bt->smem_eig_main = bt->smem_eig + eig_red_smem_bytes(dv.C, dv.nb);
  bt->smem_red = reduce_smem_bytes(dv.nb);
  bt->smem_chan = eig_chan_smem_bytes(dv.n, dv.nb);
  bt->smem_dens = sizeof(double) * ((size_t)AKS_MAXORB * disc->nmax_tpl + (size_t)std::max<int>(dv.n * dv.n * b.s, dv.n * dv.n) + AKS_NW + 8);
  bt->smem_poisson = sizeof(double) * ((size_t)2 * dv.Nh + 3 * dv.C + 16);
  if (bt->smem_pot > 220 * 1024 || bt->smem_eig_main > 220 * 1024 || bt->smem_dens > 220 * 1024) {
    ctx->err = "discretization too large for the shared-memory tiling of this build"; return AKS_EINVAL;
  }
  // (the kernels' dynamic shared-memory limits were raised once per ctx: set_kernel_attributes)
  {
    const size_t need = 2 * sizeof(aks_iter_record) * B;   // two slots: aks_scf_run polls one chunk while the next runs
    for (size_t i = 0; i < ctx->pinned_free.size(); ++i)
      if (ctx->pinned_free[i].second >= need) {
        bt->h_rec = (aks_iter_record*)ctx->pinned_free[i].first;
        bt->h_rec_bytes = ctx->pinned_free[i].second;
        ctx->pinned_free.erase(ctx->pinned_free.begin() + i);
        break;
      }
    if (!bt->h_rec) { CK(cudaMallocHost((void**)&bt->h_rec, need)); bt->h_rec_bytes = need; }
  }
  const char* tenv = getenv("AKS_PHASE_TIMERS");
  bt->timers = tenv && tenv[0] == '1';
  if (bt->timers) for (auto& e : bt->ev) CK(cudaEventCreate(&e));
  // stream lanes (AKS_STREAM_LANES, default 8: measured 1.57 / 1.56 / 1.52 / 1.60 / 1.66 ms per 86-atom step
  // with 4 / 6 / 8 / 12 / 16 lanes); small batches use fewer
  {
    const char* lenv = getenv("AKS_STREAM_LANES");
    int nl = lenv ? atoi(lenv) : 8;
    if (nl < 1) nl = 1;
    if (nl > 16) nl = 16;
    while (nl > 1 && bt->B < 4 * nl) nl /= 2;   // at least 4 problems (one k_potential CTA column) per lane
    const int NP = potential_np(dv.n, b.s);
    if (!ctx->event_free.empty()) { bt->fork_ev = ctx->event_free.back(); ctx->event_free.pop_back(); }
    else CK(cudaEventCreateWithFlags(&bt->fork_ev, cudaEventDisableTiming));
    if (nl > 1) {
      int prev = 0;
      for (int k = 0; k < nl; ++k) {
        int end = (k == nl - 1) ? bt->B : (int)(((long long)bt->B * (k + 1) / nl) / NP * NP);
        if (end <= prev) continue;
        Lane ln;
        ln.p0 = prev; ln.B = end - prev;
        if (!ctx->stream_free.empty()) { ln.stream = ctx->stream_free.back(); ctx->stream_free.pop_back(); }
        else CK(cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking));
        if (!ctx->event_free.empty()) { ln.done = ctx->event_free.back(); ctx->event_free.pop_back(); }
        else CK(cudaEventCreateWithFlags(&ln.done, cudaEventDisableTiming));
        bt->lanes.push_back(ln);
        prev = end;
      }
    }
  }
  // ODA defaults for every problem
  bt->hScftol.assign(nprob, b.scftol);
  bt->hTinit.assign(nprob, b.tinit);
  bt->hFrozen.assign(nprob, 0);
  bt->hMaxiter.assign(nprob, 0);
  CK(cudaMemcpyAsync(b.scftol_p, bt->hScftol.data(), sizeof(double) * nprob, cudaMemcpyHostToDevice, ctx->stream));
  return aks_batch_reset(bt);
}

extern "C" void aks_batch_destroy(aks_batch* bt) {
  if (!bt) return;
  GUARD(bt->disc->ctx);
  {
    aks_ctx* cx = bt->disc->ctx;
    for (Lane& ln : bt->lanes) {
      if (ln.stream) { cudaStreamSynchronize(ln.stream); cx->stream_free.push_back(ln.stream); }
      if (ln.done) cx->event_free.push_back(ln.done);
      if (ln.gexec) cudaGraphExecDestroy(ln.gexec);
    }
    if (bt->whole.gexec) { cudaStreamSynchronize(cx->stream); cudaGraphExecDestroy(bt->whole.gexec); }
    if (bt->fork_ev) cx->event_free.push_back(bt->fork_ev);
    for (auto& e : bt->poll_ev) if (e) cx->event_free.push_back(e);
  }
  cudaStream_t st = bt->disc->ctx->stream;
  for (void* p : bt->allocs) cudaFreeAsync(p, st);
  cudaStreamSynchronize(st);
  if (bt->h_rec) bt->disc->ctx->pinned_free.push_back({bt->h_rec, bt->h_rec_bytes});
  for (auto& e : bt->ev) if (e) cudaEventDestroy(e);
  for (auto& e : bt->kev) if (e) cudaEventDestroy(e);
  dd_batch_free(bt);
  delete bt;
}

extern "C" int aks_batch_set_oda(aks_batch* bt, double scftol, double tinit, int32_t frozen_t,
                                 int32_t maxiter_ls, double abstol_ls, double reltol_ls, int32_t max_degen,
                                 double degen_tol, int32_t handle_degen_spin_1iter) {
  if (!bt) return AKS_EINVAL;
  GUARD(bt->disc->ctx);
  enter_main(bt);
  aks_ctx* ctx = bt->disc->ctx;
  if (!(scftol >= 0) || !(tinit >= 0 && tinit <= 1) || maxiter_ls < 1 || max_degen < 1 || !(degen_tol >= 0)) {
    ctx->err = "aks_batch_set_oda: bad option"; return AKS_EINVAL;
  }
  if (bt->dd) return dd_set_oda(bt, scftol, tinit, frozen_t, max_degen, degen_tol, maxiter_ls, abstol_ls, reltol_ls);
  BatchDev& b = bt->dev;
  b.scftol = scftol; b.tinit = tinit; b.frozen_t = frozen_t; b.maxiter_ls = maxiter_ls;
  b.abstol_ls = abstol_ls; b.reltol_ls = reltol_ls; b.max_degen = max_degen; b.degen_tol = degen_tol;
  b.handle_spin = handle_degen_spin_1iter;
  bt->graphs_valid = false;   // the options are kernel arguments, baked into the captured step
  // every problem back to the batch-wide algorithm (aks_batch_set_oda_problem overrides one)
  bt->hScftol.assign(bt->B, scftol);
  bt->hTinit.assign(bt->B, tinit);
  bt->hFrozen.assign(bt->B, frozen_t ? 1 : 0);
  bt->hMaxiter.assign(bt->B, 0);
  CK(cudaMemcpyAsync(b.t, bt->hTinit.data(), sizeof(double) * bt->B, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(b.scftol_p, bt->hScftol.data(), sizeof(double) * bt->B, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(b.frozen_p, bt->hFrozen.data(), sizeof(int) * bt->B, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(b.maxiter_p, bt->hMaxiter.data(), sizeof(int) * bt->B, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return AKS_OK;
}

extern "C" int aks_batch_set_oda_problem(aks_batch* bt, int32_t prob, double scftol, double tinit, int32_t frozen_t,
                                         int32_t maxiter) {
  if (!bt || prob < 0 || prob >= bt->B) return AKS_EINVAL;
  GUARD(bt->disc->ctx);
  DD_UNSUPPORTED(bt, "aks_batch_set_oda_problem");
  enter_main(bt);
  aks_ctx* ctx = bt->disc->ctx;
  if (!(scftol >= 0) || !(tinit >= 0 && tinit <= 1) || maxiter < 0) {
    ctx->err = "aks_batch_set_oda_problem: need scftol >= 0, 0 <= tinit <= 1, maxiter >= 0"; return AKS_EINVAL;
  }
  BatchDev& b = bt->dev;
  bt->hScftol[prob] = scftol; bt->hTinit[prob] = tinit; bt->hFrozen[prob] = frozen_t ? 1 : 0; bt->hMaxiter[prob] = maxiter;
  CK(cudaMemcpyAsync(b.t + prob, &bt->hTinit[prob], sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(b.scftol_p + prob, &bt->hScftol[prob], sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(b.frozen_p + prob, &bt->hFrozen[prob], sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(b.maxiter_p + prob, &bt->hMaxiter[prob], sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return AKS_OK;
}

extern "C" int aks_batch_reset(aks_batch* bt) {
  if (!bt) return AKS_EINVAL;
  GUARD(bt->disc->ctx);
  if (bt->dd) return dd_batch_reset(bt);
  enter_main(bt);
  aks_ctx* ctx = bt->disc->ctx;
  BatchDev& b = bt->dev;
  const DiscDev& dv = bt->disc->dev;
  const size_t B = bt->B, s = b.s;
  CK(cudaMemsetAsync(b.Dd, 0, sizeof(double) * B * s * dv.Nh * dv.Nh, ctx->stream));
  CK(cudaMemsetAsync(b.rho_q, 0, sizeof(double) * B * s * dv.C * dv.Q, ctx->stream));
  CK(cudaMemsetAsync(b.rho_y, 0, sizeof(double) * B * s * dv.G, ctx->stream));
  CK(cudaMemsetAsync(b.Bv, 0, sizeof(double) * B * dv.Nh, ctx->stream));
  CK(cudaMemsetAsync(b.Wv, 0, sizeof(double) * B * dv.Nh, ctx->stream));
  CK(cudaMemsetAsync(b.E, 0, sizeof(double) * B * 5, ctx->stream));
  CK(cudaMemsetAsync(b.stop, 0, sizeof(double) * B, ctx->stream));
  CK(cudaMemsetAsync(b.niter, 0, sizeof(int) * B, ctx->stream));
  CK(cudaMemsetAsync(b.status, 0, sizeof(int) * B, ctx->stream));
  CK(cudaMemsetAsync(b.warm, 0, sizeof(int) * B, ctx->stream));
  CK(cudaMemsetAsync(b.degen, 0, sizeof(int) * B, ctx->stream));
  CK(cudaMemsetAsync(b.eig_fail, 0, sizeof(int) * B, ctx->stream));
  CK(cudaMemsetAsync(b.rec, 0, sizeof(aks_iter_record) * B, ctx->stream));
  CK(cudaMemcpyAsync(b.t, bt->hTinit.data(), sizeof(double) * B, cudaMemcpyHostToDevice, ctx->stream));
  int rc = window_full(bt, 0, bt->B);
  if (rc != AKS_OK) return rc;
  bt->eig_dirty = false;
  CK(cudaStreamSynchronize(ctx->stream));
  return AKS_OK;
}

extern "C" int aks_batch_set_aufbau(aks_batch* bt, int32_t kind, double temp, const double* nfix) {
  if (!bt) return AKS_EINVAL;
  GUARD(bt->disc->ctx);
  bt->graphs_valid = false;
  if (bt->dd) {
    if (kind < 0 || kind > 2 || (kind == 1 && !(temp > 0.0)) || (kind == 2 && !nfix)) {
      bt->disc->ctx->err = "aks_batch_set_aufbau: kind in {0,1,2}, temp > 0 for 1 (SmearedAufbau), nfix for 2 (FrozenAufbau)";
      return AKS_EINVAL;
    }
    return dd_set_aufbau(bt, kind, temp, nfix);
  }
  enter_main(bt);
  aks_ctx* ctx = bt->disc->ctx;
  BatchDev& b = bt->dev;
  if (kind < 0 || kind > 2 || (kind == 1 && !(temp > 0.0)) || (kind == 2 && !nfix)) {
    ctx->err = "aks_batch_set_aufbau: kind in {0,1,2}, temp > 0 for 1 (SmearedAufbau), nfix for 2 (FrozenAufbau)";
    return AKS_EINVAL;
  }
  int rc = ensure_complete(bt);
  if (rc != AKS_OK) return rc;
  if (kind == 2) {
    // caller: per problem (L, nh[, 2]) column-major  ->  device [s][Lmax][nhmax]
    const size_t s = b.s, L = b.Lmax, nh = b.nhmax, per = s * L * nh;
    std::vector<double> h((size_t)bt->B * per, 0.0);
    for (size_t p = 0; p < (size_t)bt->B; ++p)
      for (size_t e = 0; e < s; ++e)
        for (size_t l = 0; l < L; ++l)
          for (size_t k = 0; k < nh; ++k) h[p * per + (e * L + l) * nh + k] = nfix[p * per + l + L * (k + nh * e)];
    CK(cudaMemcpyAsync(b.nfix, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  b.aufbau_kind = kind;
  b.temp = temp;
  return window_full(bt, 0, bt->B);
}

extern "C" int aks_batch_set_eig_window(aks_batch* bt, int32_t margin) {
  if (!bt) return AKS_EINVAL;
  GUARD(bt->disc->ctx);
  bt->graphs_valid = false;
  if (bt->dd) return dd_set_window(bt, margin);
  enter_main(bt);
  int rc = ensure_complete(bt);
  if (rc != AKS_OK) return rc;
  bt->dev.eig_margin = margin;
  return window_full(bt, 0, bt->B);
}

// ------------------------------------------------------------------------------------ SCF step
#define LAUNCH_CHECK()                                                                        \
  do {                                                                                        \
    ctx->launches += 1;                                                                       \
    cudaError_t e_ = cudaGetLastError();                                                      \
    if (e_ != cudaSuccess) {                                                                  \
      char buf_[256];                                                                         \
      snprintf(buf_, sizeof buf_, "%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      ctx->err = buf_;                                                                        \
      return AKS_ECUDA;                                                                       \
    }                                                                                         \
  } while (0)

// AKS_PHASE_TIMERS: an event after every kernel of the step (name = the kernel just launched; nullptr = start)
static void ktick(aks_batch* bt, const char* name) {
  if (!(bt->timers && bt->cur == &bt->whole)) return;
  if (name == nullptr) bt->kcount = 0;
  if (bt->kcount == bt->kev.size()) {
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    bt->kev.push_back(e);
    bt->knames.push_back(nullptr);
  }
  cudaEventRecord(bt->kev[bt->kcount], bt->cur->stream);
  bt->knames[bt->kcount] = name;
  bt->kcount += 1;
}

// problems [p0, p0 + np); a CTA covers one cell and potential_np() problems
static int launch_potential_range(aks_batch* bt, double* vxc_out, double* har_out, int p0, int np, int force) {
  aks_ctx* ctx = bt->disc->ctx;
  const DiscDev& dv = bt->disc->dev;
  const int NP = potential_np(dv.n, bt->cur->dev.s);
  const bool tm = bt->timers && bt->cur == &bt->whole;
  if (tm) CK(cudaEventRecord(bt->ev[11], bt->cur->stream));
  k_xc<<<dim3((dv.Q + 255) / 256, dv.C, np), 256, 0, bt->cur->stream>>>(dv, bt->cur->dev, p0, np, force);
  LAUNCH_CHECK();
  ktick(bt, "k_xc");
  if (tm) CK(cudaEventRecord(bt->ev[12], bt->cur->stream));
  const dim3 grid(dv.C, (np + NP - 1) / NP);
  if (bt->pot_tc) {
    const int npc = POT_TC_NU / bt->cur->dev.s;
    const dim3 gtc(dv.C, (np + npc - 1) / npc);
    k_potential_tc<<<gtc, POT_TC_NT, bt->smem_pot, bt->cur->stream>>>(dv, bt->cur->dev, p0, np, force);
    LAUNCH_CHECK();
    ktick(bt, "k_potential_tc");
    k_hassemble<<<gtc, AKS_NT, hassemble_smem_bytes(dv.n), bt->cur->stream>>>(dv, bt->cur->dev, vxc_out, har_out, p0, np, force);
    LAUNCH_CHECK();
    ktick(bt, "k_hassemble");
    if (tm) CK(cudaEventRecord(bt->ev[13], bt->cur->stream));
    return AKS_OK;
  }
  if (potential_nts(dv.n) == 1)
    k_potential<1, 4><<<grid, AKS_NT, bt->smem_pot, bt->cur->stream>>>(dv, bt->cur->dev, vxc_out, har_out, p0, np, force);
  else
    k_potential<2, 2><<<grid, AKS_NT, bt->smem_pot, bt->cur->stream>>>(dv, bt->cur->dev, vxc_out, har_out, p0, np, force);
  LAUNCH_CHECK();
  ktick(bt, "k_potential");
  if (tm) CK(cudaEventRecord(bt->ev[13], bt->cur->stream));
  return AKS_OK;
}
static int launch_potential(aks_batch* bt, double* vxc_out, double* har_out) {
  return launch_potential_range(bt, vxc_out, har_out, 0, bt->cur->B, 0);
}
// one pass of the eigensolver over the levels its mode selects (k_eig: 0 window, 1 refill, 2 completion)
static int launch_eig_pass(aks_batch* bt, int mode) {
  aks_ctx* ctx = bt->disc->ctx;
  const DiscDev& dv = bt->disc->dev;
  const BatchDev& b = bt->cur->dev;
  static const char* const nm[3][4] = {{"k_eig_init", "k_eig", "k_eig_final", "k_eig_scale"},
                                       {"k_eig_init (refill)", "k_eig (refill)", "k_eig_final (refill)", "k_eig_scale (refill)"},
                                       {"k_eig_init (completion)", "k_eig (completion)", "k_eig_final (completion)", "k_eig_scale (completion)"}};
  k_eig_init<<<dim3((dv.C + EIGC_CHUNK - 1) / EIGC_CHUNK, b.Lmax * b.s, bt->cur->B), EIGI_NT, bt->smem_chan, bt->cur->stream>>>(dv, b, mode);
  LAUNCH_CHECK();
  ktick(bt, nm[mode][0]);
  k_eig<<<dim3(b.nhmax, b.Lmax * b.s, bt->cur->B), EIG_NT, bt->smem_eig_main, bt->cur->stream>>>(dv, b, mode);
  LAUNCH_CHECK();
  ktick(bt, nm[mode][1]);
  const int nchunk = (dv.C + EIGC_CHUNK - 1) / EIGC_CHUNK;
  k_eig_final<<<dim3(nchunk, b.Lmax * b.s, bt->cur->B), EIGF_NT, 0, bt->cur->stream>>>(dv, b, mode);
  LAUNCH_CHECK();
  ktick(bt, nm[mode][2]);
  k_eig_scale<<<dim3(b.Lmax * b.s, bt->cur->B), AKS_NT, 0, bt->cur->stream>>>(dv, b, dv.C, mode);
  LAUNCH_CHECK();
  ktick(bt, nm[mode][3]);
  return AKS_OK;
}
static int launch_eig(aks_batch* bt) {
  aks_ctx* ctx = bt->disc->ctx;
  const DiscDev& dv = bt->disc->dev;
  const BatchDev& b = bt->cur->dev;
  {
    const dim3 grid((dv.C + RED_NW - 1) / RED_NW, b.Lmax * b.s, bt->cur->B);
    switch (reduce_nbm(dv.nb)) {
      case 8: k_reduce<8><<<grid, RED_NT, bt->smem_red, bt->cur->stream>>>(dv, b); break;
      case 12: k_reduce<12><<<grid, RED_NT, bt->smem_red, bt->cur->stream>>>(dv, b); break;
      case 20: k_reduce<20><<<grid, RED_NT, bt->smem_red, bt->cur->stream>>>(dv, b); break;
      default: k_reduce<28><<<grid, RED_NT, bt->smem_red, bt->cur->stream>>>(dv, b); break;
    }
  }
  LAUNCH_CHECK();
  ktick(bt, "k_reduce");
  return launch_eig_pass(bt, 0);
}
// every level current again (the reference returns all nh eigenpairs of the last Hamiltonian)
static int ensure_complete(aks_batch* bt) {
  if (!bt->eig_dirty) return AKS_OK;
  aks_ctx* ctx = bt->disc->ctx;
  k_set_redo<<<(bt->cur->B + 127) / 128, 128, 0, bt->cur->stream>>>(bt->cur->dev, 1);
  LAUNCH_CHECK();
  int rc = launch_eig_pass(bt, 2);
  if (rc != AKS_OK) return rc;
  k_set_redo<<<(bt->cur->B + 127) / 128, 128, 0, bt->cur->stream>>>(bt->cur->dev, 0);
  LAUNCH_CHECK();
  bt->eig_dirty = false;
  return AKS_OK;
}
static int window_full(aks_batch* bt, int p0, int np) {
  aks_ctx* ctx = bt->disc->ctx;
  k_window_full<<<(np * bt->cur->dev.s * bt->cur->dev.Lmax + 127) / 128, 128, 0, bt->cur->stream>>>(bt->cur->dev, p0, np);
  LAUNCH_CHECK();
  return AKS_OK;
}
static int launch_dens_cells(aks_batch* bt) {
  aks_ctx* ctx = bt->disc->ctx;
  const DiscDev& dv = bt->disc->dev;
  const dim3 grid(dv.C, bt->cur->B, 2);
  if (ctx->dens_tc && dens_tc_ok(dv.n, dv.Q, dv.Qs)) {
    const BatchDev& bd = bt->cur->dev;
    const dim3 gtc((dv.C + DENS_TC_NC - 1) / DENS_TC_NC, bt->cur->B, 1);   // both slots of a degenerate problem in one pass
    if (bd.s == 1) k_dens_cells_tc<1><<<gtc, DENS_TC_NT, dens_tc_smem_bytes(), bt->cur->stream>>>(dv, bd);
    else k_dens_cells_tc<2><<<gtc, DENS_TC_NT, dens_tc_smem_bytes(), bt->cur->stream>>>(dv, bd);
    LAUNCH_CHECK();
    ktick(bt, "k_dens_cells_tc");
    k_fd<<<dim3(dv.C, (bt->cur->B + FD_NP - 1) / FD_NP, 2), AKS_NT, fd_smem_bytes(dv.n), bt->cur->stream>>>(dv, bt->cur->dev);
    LAUNCH_CHECK();
    ktick(bt, "k_fd");
  } else {
  switch (bt->disc->nmax_tpl) {
    case 12: k_dens_cells<12><<<grid, DENS_NT, bt->smem_dens, bt->cur->stream>>>(dv, bt->cur->dev); break;
    case 24: k_dens_cells<24><<<grid, DENS_NT, bt->smem_dens, bt->cur->stream>>>(dv, bt->cur->dev); break;
    default: k_dens_cells<32><<<grid, DENS_NT, bt->smem_dens, bt->cur->stream>>>(dv, bt->cur->dev); break;
  }
  LAUNCH_CHECK();
  ktick(bt, "k_dens_cells");
  }
  const dim3 g2((dv.G + AKS_NT - 1) / AKS_NT, bt->cur->B, 2);
  switch (bt->disc->nmax_tpl) {
    case 12: k_dens_grid<12><<<g2, AKS_NT, 0, bt->cur->stream>>>(dv, bt->cur->dev); break;
    case 24: k_dens_grid<24><<<g2, AKS_NT, 0, bt->cur->stream>>>(dv, bt->cur->dev); break;
    default: k_dens_grid<32><<<g2, AKS_NT, 0, bt->cur->stream>>>(dv, bt->cur->dev); break;
  }
  LAUNCH_CHECK();
  ktick(bt, "k_dens_grid");
  return AKS_OK;
}

// the kernel sequence of one SCF step for the problems of bt->cur, on its stream
// Segments of the step: the multi-lane driver enqueues segment by segment across the lanes, so that every
// lane has work queued within a few launches (a lane enqueued as a whole would start 60 launches late).
#define STEP_SEGMENTS 6
static int step_segment(aks_batch* bt, int seg) {
  aks_ctx* ctx = bt->disc->ctx;
  const DiscDev& dv = bt->disc->dev;
  const BatchDev& b = bt->cur->dev;
  int rc;
  const bool tm = bt->timers && bt->cur == &bt->whole;
#define TICK(i) do { if (tm) CK(cudaEventRecord(bt->ev[i], bt->cur->stream)); } while (0)
  switch (seg) {
    case 0:
      TICK(0);
      ktick(bt, nullptr);
#ifdef AKS_EIG_TIMING
      CK(cudaMemsetAsync(b.eig_dbg, 0, sizeof(long long) * (size_t)bt->cur->B * b.s * b.Lmax * b.nhmax * 8, bt->cur->stream));
#endif
      if ((rc = launch_potential(bt, nullptr, nullptr)) != AKS_OK) return rc;
      TICK(1);
      break;
    case 1:
      if ((rc = launch_eig(bt)) != AKS_OK) return rc;
      TICK(2);
      break;
    case 2:
      k_aufbau<<<bt->cur->B, 32, 0, bt->cur->stream>>>(dv, b, 0);
      LAUNCH_CHECK();
      ktick(bt, "k_aufbau");
      if (b.eig_margin >= 0) {   // problems whose window could not be certified: solve the rest, repeat
        k_eig_cert<<<dim3(b.Lmax * b.s, bt->cur->B), EIG_NT, bt->smem_eig, bt->cur->stream>>>(dv, b);
        LAUNCH_CHECK();
        ktick(bt, "k_eig_cert");
        if ((rc = launch_eig_pass(bt, 1)) != AKS_OK) return rc;
        k_aufbau<<<bt->cur->B, 32, 0, bt->cur->stream>>>(dv, b, 1);
        LAUNCH_CHECK();
        ktick(bt, "k_aufbau (refill)");
        bt->eig_dirty = true;
      }
      TICK(3);
      break;
    case 3:
      if ((rc = launch_dens_cells(bt)) != AKS_OK) return rc;
      TICK(4);
      break;
    case 4:
      k_poisson<<<dim3(bt->cur->B, 2), 128, bt->smem_poisson, bt->cur->stream>>>(dv, b, 0, 0);
      LAUNCH_CHECK();
      ktick(bt, "k_poisson");
      TICK(5);
      if (dv.G > LS_NT * LS_MAXP) { ctx->err = "more than 2048 grid points: not supported by k_linesearch"; return AKS_EINVAL; }
      if (b.s == 1) {
        if (dv.G <= 2 * LS_NT) k_linesearch<2, 1><<<bt->cur->B, LS_NT, 0, bt->cur->stream>>>(dv, b);
        else k_linesearch<4, 1><<<bt->cur->B, LS_NT, 0, bt->cur->stream>>>(dv, b);
      } else {
        if (dv.G <= 2 * LS_NT) k_linesearch<2, 2><<<bt->cur->B, LS_NT, 0, bt->cur->stream>>>(dv, b);
        else k_linesearch<4, 2><<<bt->cur->B, LS_NT, 0, bt->cur->stream>>>(dv, b);
      }
      LAUNCH_CHECK();
      ktick(bt, "k_linesearch");
      TICK(6);
      break;
    default: {
      const size_t total = (size_t)b.s * dv.C * dv.Q + (size_t)b.s * dv.G + 2 * (size_t)dv.Nh;
      const int gx = (int)std::min<size_t>((total + 255) / 256, 1024);
      k_mix_state<<<dim3(gx, bt->cur->B), 256, 0, bt->cur->stream>>>(dv, b);
      LAUNCH_CHECK();
      ktick(bt, "k_mix_state");
      TICK(9);
      if (ctx->mix_pipe) {
        const long long items = (long long)b.ntri * b.s * bt->cur->B;
        const int grid = (int)std::min<long long>(items, 2LL * ctx->sm_count);
        k_mix_dense_pipe<<<grid, AKS_NT, mix_pipe_smem_bytes(), bt->cur->stream>>>(dv, b);
      } else {
        k_mix_dense<<<dim3(b.ntri, b.s, bt->cur->B), AKS_NT, 0, bt->cur->stream>>>(dv, b);
      }
      LAUNCH_CHECK();
      ktick(bt, "k_mix_dense");
      TICK(10);
      k_finalize<<<(bt->cur->B + 127) / 128, 128, 0, bt->cur->stream>>>(dv, b);
      LAUNCH_CHECK();
      ktick(bt, "k_finalize");
      TICK(7);
    }
  }
#undef TICK
  return AKS_OK;
}
// the kernel sequence of one SCF step for the problems of bt->cur, on its stream
static int step_on_lane(aks_batch* bt) {
  for (int seg = 0; seg < STEP_SEGMENTS; ++seg) {
    const int rc = step_segment(bt, seg);
    if (rc != AKS_OK) return rc;
  }
  return AKS_OK;
}
// lane views of the batch: options and pointers are taken from bt->dev at every step (they are cheap)
static void refresh_lanes(aks_batch* bt) {
  bt->whole.dev = bt->dev;
  bt->whole.B = bt->B;
  bt->whole.p0 = 0;
  bt->whole.stream = bt->disc->ctx->stream;
  for (Lane& ln : bt->lanes) {
    ln.dev = bt->dev;
    ln.dev.B = ln.B;
    for (const FieldInfo& f : bt->fields) {
      char** ptr = reinterpret_cast<char**>(reinterpret_cast<char*>(&ln.dev) + f.offset);
      if (*ptr) *ptr += (size_t)ln.p0 * f.bytes_per_problem;
    }
  }
}
// anything enqueued on the ctx stream from now on is ordered after the lanes' work (every step ends with
// the ctx stream waiting for the lanes); the next step must make the lanes wait for the ctx stream again
static void main_side(aks_batch* bt) { bt->forked = false; bt->cur = &bt->whole; }

// The step sequence of a lane (21 kernels, all with fixed grids: finished problems exit at block entry) is
// captured ONCE per set of options into a CUDA graph and replayed: one graph launch per lane and step instead of
// 21 kernel launches, and the kernels of a replay follow each other without the per-launch gap of the stream.
static void drop_graphs(aks_batch* bt) {
  for (Lane& ln : bt->lanes) if (ln.gexec) { cudaGraphExecDestroy(ln.gexec); ln.gexec = nullptr; }
  if (bt->whole.gexec) { cudaGraphExecDestroy(bt->whole.gexec); bt->whole.gexec = nullptr; }
}
static int capture_lane(aks_batch* bt, Lane* ln) {
  aks_ctx* ctx = bt->disc->ctx;
  const int64_t l0 = ctx->launches;
  bt->cur = ln;
  CK(cudaStreamBeginCapture(ln->stream, cudaStreamCaptureModeThreadLocal));
  const int rc = step_on_lane(bt);
  cudaGraph_t graph = nullptr;
  const cudaError_t e = cudaStreamEndCapture(ln->stream, &graph);
  ln->gkernels = (int)(ctx->launches - l0);
  ctx->launches = l0;   // a capture launches nothing
  if (rc != AKS_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
  if (e != cudaSuccess) { ctx->err = std::string("cudaStreamEndCapture -> ") + cudaGetErrorString(e); return AKS_ECUDA; }
  CK(cudaGraphInstantiate(&ln->gexec, graph, 0));
  CK(cudaGraphDestroy(graph));
  return AKS_OK;
}
// Graph replay pays where launch latency matters: measured (profiles/r02_e2e_breakdown.txt) 0.439 vs 0.468 ms per
// step for one atom, 0.582 vs 0.600 for 11, nothing for 86 (the step is GPU bound) -- while capturing and
// instantiating eight lane graphs costs 1.4 ms per batch.  Large batches therefore launch plainly.
static bool graphs_on(const aks_batch* bt) { return bt->disc->ctx->use_graphs && bt->B <= AKS_GRAPH_MAX_BATCH; }
static int run_lane(aks_batch* bt, Lane* ln) {
  aks_ctx* ctx = bt->disc->ctx;
  if (!graphs_on(bt)) { bt->cur = ln; return step_on_lane(bt); }
  if (!ln->gexec) { const int rc = capture_lane(bt, ln); if (rc != AKS_OK) return rc; }
  CK(cudaGraphLaunch(ln->gexec, ln->stream));
  ctx->launches += ln->gkernels;
  if (bt->dev.eig_margin >= 0) bt->eig_dirty = true;   // what step_segment notes when it runs the window pass
  return AKS_OK;
}

static int step_launch(aks_batch* bt) {
  aks_ctx* ctx = bt->disc->ctx;
  if (bt->dd) { refresh_lanes(bt); main_side(bt); return dd_step_launch(bt); }
  refresh_lanes(bt);
  if (!bt->graphs_valid) { drop_graphs(bt); bt->graphs_valid = true; }
  if (bt->timers) {   // per-phase events: plain launches on the ctx stream
    bt->cur = &bt->whole;
    bt->forked = false;
    return step_on_lane(bt);
  }
  if (bt->lanes.size() < 2) {
    bt->forked = false;
    const int rc = run_lane(bt, &bt->whole);
    bt->cur = &bt->whole;
    return rc;
  }
  if (!bt->forked) {
    CK(cudaEventRecord(bt->fork_ev, ctx->stream));
    for (Lane& ln : bt->lanes) CK(cudaStreamWaitEvent(ln.stream, bt->fork_ev, 0));
    bt->forked = true;
  }
  int rc = AKS_OK;
  if (graphs_on(bt)) {
    for (Lane& ln : bt->lanes)
      if ((rc = run_lane(bt, &ln)) != AKS_OK) break;
  } else {
    for (int seg = 0; seg < STEP_SEGMENTS && rc == AKS_OK; ++seg)
      for (Lane& ln : bt->lanes) {
        bt->cur = &ln;
        if ((rc = step_segment(bt, seg)) != AKS_OK) break;
      }
  }
  if (rc == AKS_OK)
    for (Lane& ln : bt->lanes) {
      CK(cudaEventRecord(ln.done, ln.stream));
      CK(cudaStreamWaitEvent(ctx->stream, ln.done, 0));
    }
  bt->cur = &bt->whole;
  return rc;
}

static int fetch_records(aks_batch* bt, aks_iter_record* out) {
  aks_ctx* ctx = bt->disc->ctx;
  main_side(bt);
  CK(cudaMemcpyAsync(bt->h_rec, bt->dev.rec, sizeof(aks_iter_record) * bt->B, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (bt->timers) {
    for (int i = 0; i < 7; ++i) {
      float ms = 0;
      cudaEventElapsedTime(&ms, bt->ev[i], bt->ev[i + 1]);
      bt->phase_ms[i] = ms;
    }
    float msd = 0;
    cudaEventElapsedTime(&msd, bt->ev[9], bt->ev[10]);
    bt->phase_ms[7] = msd;   // k_mix_dense alone (part of phase 6)
    cudaEventElapsedTime(&msd, bt->ev[12], bt->ev[13]);
    bt->phase_ms[8] = msd;   // k_potential alone (part of phase 0)
    cudaEventElapsedTime(&msd, bt->ev[11], bt->ev[12]);
    bt->phase_ms[9] = msd;   // k_xc alone
    bt->kms.assign(bt->kcount, 0.0);
    for (size_t i = 1; i < bt->kcount; ++i) {
      float ms = 0;
      cudaEventElapsedTime(&ms, bt->kev[i - 1], bt->kev[i]);
      bt->kms[i] = ms;
    }
  }
  if (out) memcpy(out, bt->h_rec, sizeof(aks_iter_record) * bt->B);
  return AKS_OK;
}

extern "C" int aks_scf_step(aks_batch* bt, aks_iter_record* out) {
  if (!bt) return AKS_EINVAL;
  GUARD(bt->disc->ctx);
  int rc = step_launch(bt);
  if (rc != AKS_OK) return rc;
  if (!out && !bt->timers) return AKS_OK;  // enqueue only: the caller synchronises the ctx stream
  return fetch_records(bt, out);
}

// solve! (groundstate.jl:53-62).  The loop condition lives on the device: k_finalize sets the status word of a
// problem (SUCCESS / MAXITERS of its own cap / error) and every kernel skips finished problems, so the host only
// has to look at the status words every `check_every` steps; the per-step records the reference's LogBook wants
// are written into a device-side history by k_finalize and copied once at the end.
extern "C" int aks_scf_run(aks_batch* bt, int32_t maxiter, aks_iter_record* history, aks_iter_record* final_records) {
  if (!bt || maxiter < 1) return AKS_EINVAL;
  GUARD(bt->disc->ctx);
  enter_main(bt);
  aks_ctx* ctx = bt->disc->ctx;
  const size_t B = bt->B;
  int rc;
  const bool dev_hist = history && !bt->dd;
  if (dev_hist) {
    BatchDev& b = bt->dev;
    if (bt->hist_cap_alloc < (size_t)maxiter) {
      void* ptr = nullptr;
      CK(cudaMallocAsync(&ptr, sizeof(aks_iter_record) * B * (size_t)maxiter, ctx->stream));
      bt->allocs.push_back(ptr);   // an older, smaller buffer stays in the list until the batch is destroyed
      b.hist = (aks_iter_record*)ptr;
      bt->hist_cap_alloc = (size_t)maxiter;
      bt->graphs_valid = false;
    }
    if (b.hist_cap != maxiter) { b.hist_cap = maxiter; bt->graphs_valid = false; }
    CK(cudaMemsetAsync(b.hist, 0, sizeof(aks_iter_record) * B * (size_t)maxiter, ctx->stream));
    CK(cudaMemcpyAsync(b.niter0, b.niter, sizeof(int) * B, cudaMemcpyDeviceToDevice, ctx->stream));
  } else if (!bt->dd && bt->dev.hist_cap != 0) {
    bt->dev.hist_cap = 0;   // no history wanted: k_finalize's slot test fails for every step
    bt->graphs_valid = false;
  }
  std::vector<aks_iter_record> last(B);
  // records persist on the device between steps; start from the current ones
  if ((rc = fetch_records(bt, last.data())) != AKS_OK) return rc;
  const std::vector<aks_iter_record> first = last;
  if (bt->dd && history) {
    // a Double64 batch keeps its history on the host: one poll per step
    for (int it = 0; it < maxiter; ++it) {
      if ((rc = step_launch(bt)) != AKS_OK) return rc;
      if ((rc = fetch_records(bt, nullptr)) != AKS_OK) return rc;
      bool running = false;
      for (size_t i = 0; i < B; ++i) {
        const aks_iter_record& r = bt->h_rec[i];
        if (r.niter != last[i].niter || r.status != last[i].status) last[i] = r;
        history[(size_t)it * B + i] = last[i];
        running = running || last[i].status == AKS_RUNNING;
      }
      if (!running) {   // a problem that was not running keeps its old record
        for (int k = it + 1; k < maxiter; ++k) for (size_t i = 0; i < B; ++i) history[(size_t)k * B + i] = last[i];
        break;
      }
    }
  } else {
    // chunks of `check_every` steps; the status words of chunk i are read back while chunk i + 1 already runs, so
    // the device never waits for the host.  A chunk enqueued after every problem has finished only launches kernels
    // that return at block entry.
    const int every = std::max(1, ctx->check_every);
    for (auto& e : bt->poll_ev)
      if (!e) {
        if (!ctx->event_free.empty()) { e = ctx->event_free.back(); ctx->event_free.pop_back(); }
        else CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      }
    int enq = 0;
    auto enqueue_chunk = [&](int slot) -> int {
      const int nstep = std::min(every, maxiter - enq);
      for (int q = 0; q < nstep; ++q) { const int r_ = step_launch(bt); if (r_ != AKS_OK) return r_; }
      enq += nstep;
      main_side(bt);
      CK(cudaMemcpyAsync(bt->h_rec + (size_t)slot * B, bt->dev.rec, sizeof(aks_iter_record) * B, cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaEventRecord(bt->poll_ev[slot], ctx->stream));
      return AKS_OK;
    };
    int slot = 0;
    if ((rc = enqueue_chunk(slot)) != AKS_OK) return rc;
    while (true) {
      const bool more = enq < maxiter;
      if (more && (rc = enqueue_chunk(slot ^ 1)) != AKS_OK) return rc;
      CK(cudaEventSynchronize(bt->poll_ev[slot]));
      bool running = false;
      for (size_t i = 0; i < B; ++i) {
        last[i] = bt->h_rec[(size_t)slot * B + i];
        running = running || last[i].status == AKS_RUNNING;
      }
      if (!running || !more) break;
      slot ^= 1;
    }
    CK(cudaStreamSynchronize(ctx->stream));   // a chunk may still be in flight (no-ops)
  }
  if (dev_hist) {
    CK(cudaMemcpyAsync(history, bt->dev.hist, sizeof(aks_iter_record) * B * (size_t)maxiter, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    // slots a problem did not step in (finished earlier) repeat its last record, as the per-step host loop did
    for (size_t i = 0; i < B; ++i) {
      aks_iter_record cur = first[i];
      for (int k = 0; k < maxiter; ++k) {
        aks_iter_record& h = history[(size_t)k * B + i];
        if (h.niter > 0) cur = h; else h = cur;
      }
    }
  }
  // problems still running have exhausted maxiter
  bool notconv = false, any_error = false;
  std::vector<int> st(B);
  for (size_t i = 0; i < B; ++i) {
    st[i] = last[i].status;
    if (st[i] == AKS_RUNNING) { st[i] = AKS_MAXITERS; last[i].status = AKS_MAXITERS; }
    if (st[i] == AKS_MAXITERS) notconv = true;
    if (st[i] < 0) any_error = true;
  }
  if (notconv) {
    CK(cudaMemcpyAsync(bt->dev.status, st.data(), sizeof(int) * B, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  if (final_records) memcpy(final_records, last.data(), sizeof(aks_iter_record) * B);
  if (any_error)
    for (size_t i = 0; i < B; ++i) if (last[i].status < 0) return last[i].status;
  return notconv ? AKS_ENOTCONV : AKS_OK;
}

extern "C" int aks_last_kernel_ms(aks_batch* bt, int32_t cap, double* ms, char* names, int32_t name_stride, int32_t* count) {
  if (!bt || !count) return AKS_EINVAL;
  if (!bt->timers) return AKS_EINVAL;
  const int n = bt->kms.empty() ? 0 : (int)bt->kms.size() - 1;   // entry 0 is the start marker
  *count = n;
  for (int i = 0; i < n && i < cap; ++i) {
    if (ms) ms[i] = bt->kms[i + 1];
    if (names && name_stride > 0) {
      snprintf(names + (size_t)i * name_stride, (size_t)name_stride, "%s", bt->knames[i + 1] ? bt->knames[i + 1] : "?");
    }
  }
  return AKS_OK;
}

extern "C" int aks_last_phase_ms(aks_batch* bt, double out[10]) {
  if (!bt || !out) return AKS_EINVAL;
  GUARD(bt->disc->ctx);
  for (int i = 0; i < 10; ++i) out[i] = bt->phase_ms[i];
  return bt->timers ? AKS_OK : AKS_EINVAL;
}

// ------------------------------------------------------------------------------------ state I/O
extern "C" int aks_get_state(aks_batch* bt, int32_t prob, double* D, double* U, double* eps, double* nocc, double* W) {
  if (!bt || prob < 0 || prob >= bt->B) return AKS_EINVAL;
  GUARD(bt->disc->ctx);
  if (bt->dd) return dd_get_state(bt, prob, D, U, eps, nocc, W);
  enter_main(bt);
  aks_ctx* ctx = bt->disc->ctx;
  const DiscDev& dv = bt->disc->dev;
  const BatchDev& b = bt->dev;
  const size_t s = b.s, Nh = dv.Nh, L = b.Lmax, nh = b.nhmax;
  if (U || eps) { int rc = ensure_complete(bt); if (rc != AKS_OK) return rc; }
  if (D) {   // the iterations keep only the lower triangle of D current
    const int nb32 = (int)((Nh + 31) / 32);
    k_mirror_dense<<<dim3(nb32, nb32, (unsigned)s), AKS_NT, 0, ctx->stream>>>(dv, b, prob);
    LAUNCH_CHECK();
  }
  if (D) CK(cudaMemcpyAsync(D, b.Dd + (size_t)prob * s * Nh * Nh, sizeof(double) * s * Nh * Nh, cudaMemcpyDeviceToHost, ctx->stream));
  if (U) CK(cudaMemcpyAsync(U, b.U + (size_t)prob * s * L * nh * Nh, sizeof(double) * s * L * nh * Nh, cudaMemcpyDeviceToHost, ctx->stream));
  std::vector<double> he(s * L * nh), hn(s * L * nh), hw(Nh);
  if (eps) CK(cudaMemcpyAsync(he.data(), b.eps + (size_t)prob * s * L * nh, sizeof(double) * s * L * nh, cudaMemcpyDeviceToHost, ctx->stream));
  if (nocc) CK(cudaMemcpyAsync(hn.data(), b.occ + (size_t)prob * 2 * s * L * nh, sizeof(double) * s * L * nh, cudaMemcpyDeviceToHost, ctx->stream));
  if (W) CK(cudaMemcpyAsync(hw.data(), b.Wv + (size_t)prob * Nh, sizeof(double) * Nh, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  // device [s][L][nh]  ->  Julia (L, nh[, s]) column-major
  for (size_t e = 0; e < s; ++e)
    for (size_t l = 0; l < L; ++l)
      for (size_t k = 0; k < nh; ++k) {
        if (eps) eps[l + L * (k + nh * e)] = he[(e * L + l) * nh + k];
        if (nocc) nocc[l + L * (k + nh * e)] = hn[(e * L + l) * nh + k];
      }
  if (W) for (size_t i = 0; i < Nh; ++i) W[i] = hw[i] * bt->hHar[prob];
  return AKS_OK;
}

// eps, n, W of EVERY problem in three copies (a sweep reads these for all atoms; per-problem calls cost a
// synchronisation each).  Layout per problem as in aks_get_state: eps, n (L, nh[, 2]) column-major, W (Nh).
extern "C" int aks_get_state_all(aks_batch* bt, double* eps, double* nocc, double* W) {
  if (!bt) return AKS_EINVAL;
  GUARD(bt->disc->ctx);
  if (bt->dd) {   // pairs: eps, n [nprob][(L, nh)][2], W [nprob][Nh][2]
    const size_t per = 2 * (size_t)bt->dev.Lmax * bt->dev.nhmax, Nh2 = 2 * (size_t)bt->disc->dev.Nh;
    for (int p = 0; p < bt->B; ++p) {
      const int rc = dd_get_state(bt, p, nullptr, nullptr, eps ? eps + p * per : nullptr, nocc ? nocc + p * per : nullptr,
                                  W ? W + p * Nh2 : nullptr);
      if (rc != AKS_OK) return rc;
    }
    return AKS_OK;
  }
  enter_main(bt);
  aks_ctx* ctx = bt->disc->ctx;
  const DiscDev& dv = bt->disc->dev;
  const BatchDev& b = bt->dev;
  const size_t s = b.s, Nh = dv.Nh, L = b.Lmax, nh = b.nhmax, B = bt->B, per = s * L * nh;
  if (eps) { int rc = ensure_complete(bt); if (rc != AKS_OK) return rc; }
  std::vector<double> he(eps ? B * per : 0), hn(nocc ? B * 2 * per : 0);
  if (eps) CK(cudaMemcpyAsync(he.data(), b.eps, sizeof(double) * B * per, cudaMemcpyDeviceToHost, ctx->stream));
  if (nocc) CK(cudaMemcpyAsync(hn.data(), b.occ, sizeof(double) * B * 2 * per, cudaMemcpyDeviceToHost, ctx->stream));
  if (W) CK(cudaMemcpyAsync(W, b.Wv, sizeof(double) * B * Nh, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  for (size_t p = 0; p < B; ++p) {
    for (size_t e = 0; e < s; ++e)
      for (size_t l = 0; l < L; ++l)
        for (size_t k = 0; k < nh; ++k) {
          if (eps) eps[p * per + l + L * (k + nh * e)] = he[p * per + (e * L + l) * nh + k];
          if (nocc) nocc[p * per + l + L * (k + nh * e)] = hn[p * 2 * per + (e * L + l) * nh + k];   // slot 0
        }
    if (W) for (size_t i = 0; i < Nh; ++i) W[p * Nh + i] *= bt->hHar[p];
  }
  return AKS_OK;
}

static int launch_state_from_dense(aks_batch* bt, int prob) {
  aks_ctx* ctx = bt->disc->ctx;
  const DiscDev& dv = bt->disc->dev;
  const BatchDev& b = bt->dev;
  switch (bt->disc->nmax_tpl) {
    case 12: k_cells_from_dense<12><<<dv.C, AKS_NT, bt->smem_dens, ctx->stream>>>(dv, b, prob); break;
    case 24: k_cells_from_dense<24><<<dv.C, AKS_NT, bt->smem_dens, ctx->stream>>>(dv, b, prob); break;
    default: k_cells_from_dense<32><<<dv.C, AKS_NT, bt->smem_dens, ctx->stream>>>(dv, b, prob); break;
  }
  LAUNCH_CHECK();
  k_grid_from_dense<<<(dv.G + AKS_NT - 1) / AKS_NT, AKS_NT, 0, ctx->stream>>>(dv, b, prob);
  LAUNCH_CHECK();
  return AKS_OK;
}

extern "C" int aks_set_density(aks_batch* bt, int32_t prob, const double* D, const double* eprev, int32_t niter) {
  if (!bt || prob < 0 || prob >= bt->B || !D || niter < 0) return AKS_EINVAL;
  GUARD(bt->disc->ctx);
  DD_UNSUPPORTED(bt, "aks_set_density");
  enter_main(bt);
  aks_ctx* ctx = bt->disc->ctx;
  const DiscDev& dv = bt->disc->dev;
  BatchDev& b = bt->dev;
  const size_t s = b.s, Nh = dv.Nh;
  CK(cudaMemcpyAsync(b.Dd + (size_t)prob * s * Nh * Nh, D, sizeof(double) * s * Nh * Nh, cudaMemcpyHostToDevice, ctx->stream));
  double e5[5] = {0, 0, 0, 0, 0};
  if (eprev) memcpy(e5, eprev, sizeof e5);
  CK(cudaMemcpyAsync(b.E + (size_t)prob * 5, e5, sizeof e5, cudaMemcpyHostToDevice, ctx->stream));
  const int zero = 0, nit = niter;
  CK(cudaMemcpyAsync(b.status + prob, &zero, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(b.degen + prob, &zero, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(b.warm + prob, &zero, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(b.niter + prob, &nit, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  int rc = ensure_complete(bt);   // the other problems keep a complete set of eigenpairs
  if (rc != AKS_OK) return rc;
  if ((rc = window_full(bt, prob, 1)) != AKS_OK) return rc;
  if ((rc = launch_state_from_dense(bt, prob)) != AKS_OK) return rc;
  // Poisson solve of slot 0 for this problem only, then commit B, W
  k_poisson<<<dim3(1, 1), 128, bt->smem_poisson, ctx->stream>>>(dv, b, prob, 1);
  LAUNCH_CHECK();
  CK(cudaMemcpyAsync(b.Bv + (size_t)prob * Nh, b.B_new + ((size_t)prob * 2) * Nh, sizeof(double) * Nh, cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemcpyAsync(b.Wv + (size_t)prob * Nh, b.W_new + ((size_t)prob * 2) * Nh, sizeof(double) * Nh, cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return AKS_OK;
}

// ------------------------------------------------------------------------------------ hooks
static void assemble_global(const aks_disc* disc, const std::vector<double>& loc, double* out /*Nh x Nh*/) {
  const DiscDev& dv = disc->dev;
  const int n = dv.n, Nh = dv.Nh;
  std::fill(out, out + (size_t)Nh * Nh, 0.0);
  for (int c = 0; c < dv.C; ++c)
    for (int a = 0; a < n; ++a) {
      const int I = disc->l2g[(size_t)c * n + a];
      if (I < 0) continue;
      for (int bq = 0; bq < n; ++bq) {
        const int J = disc->l2g[(size_t)c * n + bq];
        if (J < 0) continue;
        out[(size_t)I + (size_t)Nh * J] += loc[((size_t)c * n + a) * n + bq];
      }
    }
}

static int potential_hook(aks_batch* bt, int prob, double* vxc, double* har) {
  aks_ctx* ctx = bt->disc->ctx;
  const DiscDev& dv = bt->disc->dev;
  const BatchDev& b = bt->dev;
  const size_t cn2 = (size_t)dv.C * dv.n * dv.n;
  double *dvxc = nullptr, *dhar = nullptr;
  CK(cudaMalloc((void**)&dvxc, sizeof(double) * b.s * cn2));
  CK(cudaMalloc((void**)&dhar, sizeof(double) * cn2));
  // run the kernel for this problem only
  { int rc = launch_potential_range(bt, dvxc, dhar, prob, 1, 1); if (rc != AKS_OK) return rc; }
  std::vector<double> hv(b.s * cn2), hh(cn2);
  CK(cudaMemcpyAsync(hv.data(), dvxc, sizeof(double) * b.s * cn2, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(hh.data(), dhar, sizeof(double) * cn2, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  cudaFree(dvxc);
  cudaFree(dhar);
  if (vxc)
    for (int e = 0; e < b.s; ++e) {
      std::vector<double> one(hv.begin() + (size_t)e * cn2, hv.begin() + (size_t)(e + 1) * cn2);
      assemble_global(bt->disc, one, vxc + (size_t)e * dv.Nh * dv.Nh);
    }
  if (har) assemble_global(bt->disc, hh, har);
  return AKS_OK;
}

extern "C" int aks_assemble_hartree(aks_batch* bt, int32_t prob, double* Har) {
  if (!bt || prob < 0 || prob >= bt->B || !Har) return AKS_EINVAL;
  GUARD(bt->disc->ctx);
  DD_UNSUPPORTED(bt, "aks_assemble_hartree");
  enter_main(bt);
  return potential_hook(bt, prob, nullptr, Har);
}
extern "C" int aks_assemble_vxc(aks_batch* bt, int32_t prob, double* Vxc) {
  if (!bt || prob < 0 || prob >= bt->B || !Vxc) return AKS_EINVAL;
  GUARD(bt->disc->ctx);
  DD_UNSUPPORTED(bt, "aks_assemble_vxc");
  enter_main(bt);
  return potential_hook(bt, prob, Vxc, nullptr);
}

extern "C" int aks_eig_lowest(aks_batch* bt, int32_t prob, double* eps, double* U, double* residuals) {
  if (!bt || prob < 0 || prob >= bt->B) return AKS_EINVAL;
  GUARD(bt->disc->ctx);
  DD_UNSUPPORTED(bt, "aks_eig_lowest");
  enter_main(bt);
  aks_ctx* ctx = bt->disc->ctx;
  const DiscDev& dv = bt->disc->dev;
  const BatchDev& b = bt->dev;
  int rc;
  if ((rc = window_full(bt, 0, bt->B)) != AKS_OK) return rc;
  bt->eig_dirty = false;
  if ((rc = launch_potential(bt, nullptr, nullptr)) != AKS_OK) return rc;
  if ((rc = launch_eig(bt)) != AKS_OK) return rc;
  const size_t s = b.s, L = b.Lmax, nh = b.nhmax, Nh = dv.Nh;
  std::vector<double> he(s * L * nh), hr(s * L * nh, 0.0);
  CK(cudaMemcpyAsync(he.data(), b.eps_new + (size_t)prob * s * L * nh, sizeof(double) * s * L * nh, cudaMemcpyDeviceToHost, ctx->stream));
  if (U) CK(cudaMemcpyAsync(U, b.U + (size_t)prob * s * L * nh * Nh, sizeof(double) * s * L * nh * Nh, cudaMemcpyDeviceToHost, ctx->stream));
  if (residuals) {
    double* dres = nullptr;
    CK(cudaMalloc((void**)&dres, sizeof(double) * s * L * nh));
    CK(cudaMemsetAsync(dres, 0, sizeof(double) * s * L * nh, ctx->stream));
    k_eig_residual<<<dim3(nh, L * s), AKS_NT, sizeof(double) * (3 * Nh + 2 * dv.C + AKS_NW + 8) + sizeof(int) * 2 * dv.C, ctx->stream>>>(dv, b, prob, dres);
    LAUNCH_CHECK();
    CK(cudaMemcpyAsync(hr.data(), dres, sizeof(double) * s * L * nh, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(dres);
  }
  CK(cudaStreamSynchronize(ctx->stream));
  for (size_t e = 0; e < s; ++e)
    for (size_t l = 0; l < L; ++l)
      for (size_t k = 0; k < nh; ++k) {
        if (eps) eps[l + L * (k + nh * e)] = he[(e * L + l) * nh + k];
        if (residuals) residuals[l + L * (k + nh * e)] = hr[(e * L + l) * nh + k];
      }
  return AKS_OK;
}

extern "C" int aks_density(aks_batch* bt, int32_t prob, double* D) {
  if (!bt || prob < 0 || prob >= bt->B || !D) return AKS_EINVAL;
  GUARD(bt->disc->ctx);
  DD_UNSUPPORTED(bt, "aks_density");
  enter_main(bt);
  aks_ctx* ctx = bt->disc->ctx;
  const DiscDev& dv = bt->disc->dev;
  const size_t s = bt->dev.s, Nh = dv.Nh;
  double* dout = nullptr;
  CK(cudaMallocAsync((void**)&dout, sizeof(double) * s * Nh * Nh, ctx->stream));
  k_density_dense<<<dim3((unsigned)((Nh + 15) / 16), (unsigned)((Nh + 15) / 16), (unsigned)s), AKS_NT, 0, ctx->stream>>>(dv, bt->dev, prob, dout);
  LAUNCH_CHECK();
  CK(cudaMemcpyAsync(D, dout, sizeof(double) * s * Nh * Nh, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaFreeAsync(dout, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return AKS_OK;
}

extern "C" int aks_eval_grid(aks_batch* bt, int32_t prob, int32_t kind, int32_t l, int32_t k, int32_t sigma, int64_t npts,
                             const double* x, double* out) {
  if (!bt || prob < 0 || prob >= bt->B || npts < 0 || (npts > 0 && (!x || !out))) return AKS_EINVAL;
  GUARD(bt->disc->ctx);
  DD_UNSUPPORTED(bt, "aks_eval_grid");
  aks_ctx* ctx = bt->disc->ctx;
  const DiscDev& dv = bt->disc->dev;
  const BatchDev& b = bt->dev;
  if (kind < 0 || kind > 5 || dv.n > EVAL_NMAX) { ctx->err = "aks_eval_grid: unknown kind (0..5) or order above 31"; return AKS_EINVAL; }
  if ((kind == 0 || kind == 5) && (l < 0 || l >= b.Lmax || k < 0 || k >= b.nhmax || sigma < 0 || sigma >= b.s)) {
    ctx->err = "aks_eval_grid: orbital (l, k, sigma) outside the batch"; return AKS_EINVAL;
  }
  if ((kind == 3 || kind == 4) && (sigma < 0 || sigma >= b.s)) { ctx->err = "aks_eval_grid: sigma outside 0..nspin-1"; return AKS_EINVAL; }
  if (kind == 1 && sigma >= b.s) { ctx->err = "aks_eval_grid: sigma outside -1..nspin-1"; return AKS_EINVAL; }
  if (npts == 0) return AKS_OK;
  enter_main(bt);
  double* dbuf = nullptr;
  CK(cudaMallocAsync((void**)&dbuf, sizeof(double) * 2 * (size_t)npts, ctx->stream));
  CK(cudaMemcpyAsync(dbuf, x, sizeof(double) * (size_t)npts, cudaMemcpyHostToDevice, ctx->stream));
  k_eval_grid<<<(unsigned)((npts + 127) / 128), 128, 0, ctx->stream>>>(dv, b, prob, kind, l, k, sigma, (long long)npts, dbuf, dbuf + npts);
  LAUNCH_CHECK();
  CK(cudaMemcpyAsync(out, dbuf + npts, sizeof(double) * (size_t)npts, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaFreeAsync(dbuf, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return AKS_OK;
}

extern "C" int aks_exc_energy(aks_batch* bt, int32_t prob, double* Exc) {
  if (!bt || prob < 0 || prob >= bt->B || !Exc) return AKS_EINVAL;
  GUARD(bt->disc->ctx);
  DD_UNSUPPORTED(bt, "aks_exc_energy");
  enter_main(bt);
  aks_ctx* ctx = bt->disc->ctx;
  double* dout = nullptr;
  CK(cudaMalloc((void**)&dout, sizeof(double)));
  k_exc_state<<<1, AKS_NT, 0, ctx->stream>>>(bt->disc->dev, bt->dev, prob, dout);
  LAUNCH_CHECK();
  CK(cudaMemcpyAsync(Exc, dout, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  cudaFree(dout);
  return AKS_OK;
}

extern "C" int aks_hartree_energy(aks_batch* bt, int32_t prob, double* Ehar) {
  if (!bt || prob < 0 || prob >= bt->B || !Ehar) return AKS_EINVAL;
  GUARD(bt->disc->ctx);
  DD_UNSUPPORTED(bt, "aks_hartree_energy");
  enter_main(bt);
  aks_ctx* ctx = bt->disc->ctx;
  const DiscDev& dv = bt->disc->dev;
  if (bt->hHar[prob] == 0.0) { *Ehar = 0.0; return AKS_OK; }
  std::vector<double> hb(dv.Nh), hw(dv.Nh);
  CK(cudaMemcpyAsync(hb.data(), bt->dev.Bv + (size_t)prob * dv.Nh, sizeof(double) * dv.Nh, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(hw.data(), bt->dev.Wv + (size_t)prob * dv.Nh, sizeof(double) * dv.Nh, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  double acc = 0.0;
  for (int i = 0; i < dv.Nh; ++i) acc += hb[i] * hw[i];
  *Ehar = (acc + bt->hN[prob] * bt->hN[prob] / dv.Rmax) / 2.0;
  return AKS_OK;
}

extern "C" int aks_debug_eig_cycles(aks_batch* bt, int64_t* out) {
  if (!bt || !out) return AKS_EINVAL;
  GUARD(bt->disc->ctx);
  DD_UNSUPPORTED(bt, "aks_debug_eig_cycles");
  enter_main(bt);
  aks_ctx* ctx = bt->disc->ctx;
  const BatchDev& b = bt->dev;
  const size_t cnt = (size_t)bt->B * b.s * b.Lmax * b.nhmax * 8;
  CK(cudaMemcpyAsync(out, b.eig_dbg, sizeof(int64_t) * cnt, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return AKS_OK;
}

extern "C" int aks_debug_eig_refills(aks_batch* bt, int32_t* counts) {
  if (!bt || !counts) return AKS_EINVAL;
  GUARD(bt->disc->ctx);
  if (bt->dd) return dd_eig_cold(bt, counts);   // Double64: eigenpairs of the last step that took the cold path
  enter_main(bt);
  aks_ctx* ctx = bt->disc->ctx;
  CK(cudaMemcpyAsync(counts, bt->dev.redo_cnt, sizeof(int) * bt->B, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return AKS_OK;
}

extern "C" int aks_debug_eig_info(aks_batch* bt, int32_t* info) {
  if (!bt || !info) return AKS_EINVAL;
  GUARD(bt->disc->ctx);
  DD_UNSUPPORTED(bt, "aks_debug_eig_info");
  enter_main(bt);
  aks_ctx* ctx = bt->disc->ctx;
  const BatchDev& b = bt->dev;
  const size_t cnt = (size_t)bt->B * b.s * b.Lmax * b.nhmax;
  CK(cudaMemcpyAsync(info, b.eig_info, sizeof(int) * cnt, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return AKS_OK;
}

// ------------------------------------------------------------------------------------ FP64 peak
extern "C" int aks_measure_dmma_peak(aks_ctx* ctx, double* tflops) {
  if (!ctx || !tflops) return AKS_EINVAL;
  GUARD(ctx);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, ctx->device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 14;
  double* dout = nullptr;
  CK(cudaMalloc((void**)&dout, sizeof(double) * blocks * threads));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0, ctx->stream));
    k_dmma_peak<<<blocks, threads, 0, ctx->stream>>>(dout, iters);
    LAUNCH_CHECK();
    CK(cudaEventRecord(e1, ctx->stream));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    // per warp and iteration: 8 MMAs of 8*8*4 FMAs
    const double flops = 2.0 * 256.0 * 8.0 * (double)iters * blocks * (threads / 32);
    if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(dout);
  *tflops = best;
  return AKS_OK;
}

extern "C" int aks_measure_fp64_peak(aks_ctx* ctx, double* tflops) {
  if (!ctx || !tflops) return AKS_EINVAL;
  GUARD(ctx);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, ctx->device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
  double* dout = nullptr;
  CK(cudaMalloc((void**)&dout, sizeof(double) * blocks * threads));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(e0, ctx->stream));
    k_fp64_peak<<<blocks, threads, 0, ctx->stream>>>(dout, iters);
    LAUNCH_CHECK();
    CK(cudaEventRecord(e1, ctx->stream));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
    if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(dout);
  *tflops = best;
  return AKS_OK;
}

#include "aks_dd.cuh"
