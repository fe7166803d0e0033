// Host side of the Double64 (dtype = 1) path: table splitting in double-double, device memory, the launch
// sequence of k_dd.cuh.  Included at the end of aks_b200.cu (one translation unit); the C-ABI entry points
// dispatch here when the discretization was created with dtype = 1.  Every floating-point table of
// aks_disc_desc and every floating-point output buffer is then an array of interleaved (hi, lo) pairs --
// the memory layout of Julia's Vector{Double64}.
#pragma once

struct DDDiscHost {
  DDDiscDev dev{};
};
struct DDBatchHost {
  DDBatchDev dev{};
  int max_orb_launch = DD_MAXORB;
  bool any_xc = false;
  bool eig_dirty = false;   // levels outside the eigenpair window are stale
  size_t eig_smem = 0;   // dynamic shared memory of ddk_eig (0: the packed cell blocks go to global scratch)
};

static inline dd ddh_pair(const double* a, size_t idx) { return dd_make(a[2 * idx], a[2 * idx + 1]); }

static int dd_disc_create(aks_ctx* ctx, const aks_disc_desc* D, aks_disc** out) {
  const int p = D->p, C = D->C, Nh = D->Nh, n = p + 1, nb = p - 1;
  if (p < 2 || p > DD_MAXN - 1 || C < 2 || C > DD_NT || Nh != p * C - 1 || D->lh < 0 || D->nh < 1) {
    ctx->err = "aks_disc_create(dtype=1): need 2 <= p <= 29, 2 <= C <= 64, Nh == p*C-1";
    return AKS_EINVAL;
  }
  if (D->nspin != 1) { ctx->err = "dtype=1 (Double64): nspin = 2 is not built"; return AKS_ENOSYS; }
  if ((D->lh + 1) * D->nh > DD_MAXORB) { ctx->err = "(lh+1)*nh exceeds DD_MAXORB"; return AKS_EINVAL; }
  if (!D->mesh || !D->M0 || !D->A || !D->Mm1 || !D->cell_len || !D->cell_indices || !D->cell_generators) {
    ctx->err = "aks_disc_create(dtype=1): NULL table";
    return AKS_EINVAL;
  }
  if (D->lh > 0 && !D->Mm2) { ctx->err = "Mm2 required when lh > 0"; return AKS_EINVAL; }
  aks_disc* disc = new aks_disc();
  disc->ctx = ctx;
  disc->dd = new DDDiscHost();
  DDDiscDev& dv = disc->dd->dev;
  dv.p = p; dv.n = n; dv.nb = nb; dv.C = C; dv.Nh = Nh; dv.L = D->lh + 1; dv.nh = D->nh;
  dv.Rmax = ddh_pair(D->mesh, C);
  dv.Q = D->Q; dv.G = D->G;
  if (D->Q < 0 || D->G < 0 || (D->Q > 0) != (D->G > 0)) { ctx->err = "aks_disc_create(dtype=1): Q and G must both be 0 or both positive"; dd_disc_free(disc); delete disc; return AKS_EINVAL; }
  if (D->Q > 0 && (!D->x || !D->w || !D->Pgenx || !D->invshifts || !D->y || !D->wy || !D->ycell || !D->Pgeny)) {
    ctx->err = "aks_disc_create(dtype=1): Q > 0 needs x, w, Pgenx, invshifts, y, wy, ycell, Pgeny"; dd_disc_free(disc); delete disc; return AKS_EINVAL;
  }
  {  // the Float64 view carries the sizes only
    DiscDev& d0 = disc->dev;
    d0.p = p; d0.n = n; d0.nb = nb; d0.C = C; d0.Nh = Nh; d0.L = dv.L; d0.nh = dv.nh; d0.nspin = 1;
    d0.Q = D->Q; d0.G = D->G; d0.Rmax = dv.Rmax.hi;
  }
  std::vector<int>& l2g = disc->l2g;
  l2g.assign((size_t)C * n, -1);
  for (int c = 0; c < C; ++c) {
    const int len = (int)D->cell_len[c];
    if (len < 1 || len > n) { ctx->err = "bad cell_len"; aks_disc_destroy(disc); return AKS_EINVAL; }
    for (int j = 0; j < len; ++j) {
      const int64_t I = D->cell_indices[(size_t)c * n + j], g = D->cell_generators[(size_t)c * n + j];
      if (I < 1 || I > Nh || g < 1 || g > n) { ctx->err = "bad cell index/generator"; aks_disc_destroy(disc); return AKS_EINVAL; }
      l2g[(size_t)c * n + (g - 1)] = (int)(I - 1);
    }
  }
  bool chain = (l2g[1] == -1 && l2g[(size_t)(C - 1) * n] == -1);
  for (int c = 0; c + 1 < C && chain; ++c)
    chain = l2g[(size_t)c * n] >= 0 && l2g[(size_t)c * n] == l2g[(size_t)(c + 1) * n + 1];
  for (int c = 0; c < C && chain; ++c)
    for (int i = 0; i < nb; ++i)
      if (l2g[(size_t)c * n + 2 + i] != l2g[(size_t)c * n + 2] + i) chain = false;
  if (!chain) { ctx->err = "cell maps are not a P1+bubble chain with Dirichlet ends"; aks_disc_destroy(disc); return AKS_EINVAL; }

  // global symmetric matrices (pairs, column-major) -> cell-owned blocks: an entry shared by two cells
  // belongs to the lower one
  auto split = [&](const double* Gm, std::vector<dd>& loc) {
    loc.assign((size_t)C * n * n, dd_make(0.0));
    if (!Gm) return;
    std::vector<char> owned((size_t)Nh * Nh, 0);
    for (int c = 0; c < C; ++c)
      for (int a = 0; a < n; ++a) {
        const int I = l2g[(size_t)c * n + a];
        if (I < 0) continue;
        for (int bq = 0; bq < n; ++bq) {
          const int J = l2g[(size_t)c * n + bq];
          if (J < 0 || owned[(size_t)I * Nh + J]) continue;
          owned[(size_t)I * Nh + J] = 1;
          loc[((size_t)c * n + a) * n + bq] =
              (ddh_pair(Gm, (size_t)I + (size_t)Nh * J) + ddh_pair(Gm, (size_t)J + (size_t)Nh * I)) * 0.5;
        }
      }
  };
  std::vector<dd> M0l, Al, M1l, M2l;
  split(D->M0, M0l);
  split(D->A, Al);
  split(D->Mm1, M1l);
  split(D->Mm2, M2l);
  std::vector<dd> Fl((size_t)C * n * n * n, dd_make(0.0));
  if (D->nF > 0) {
    if (!D->Fi || !D->Fj || !D->Fm || !D->Fv) { ctx->err = "NULL F arrays"; aks_disc_destroy(disc); return AKS_EINVAL; }
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
      const int I = (int)D->Fi[t] - 1, J = (int)D->Fj[t] - 1, Mi = (int)D->Fm[t] - 1;
      if (I < 0 || J < 0 || Mi < 0 || I >= Nh || J >= Nh || Mi >= Nh) { ctx->err = "bad F index"; aks_disc_destroy(disc); return AKS_EINVAL; }
      int owner = -1, la = 0, lb = 0, lm = 0;
      for (int si = 0; si < 2 && owner < 0; ++si) {
        const int c = cell_of[2 * I + si];
        if (c < 0) continue;
        int fj = -1, fm = -1;
        for (int sj = 0; sj < 2; ++sj) if (cell_of[2 * J + sj] == c) fj = loc_of[2 * J + sj];
        for (int sm2 = 0; sm2 < 2; ++sm2) if (cell_of[2 * Mi + sm2] == c) fm = loc_of[2 * Mi + sm2];
        if (fj >= 0 && fm >= 0) { owner = c; la = loc_of[2 * I + si]; lb = fj; lm = fm; }
      }
      if (owner < 0) continue;
      const dd half = ddh_pair(D->Fv, (size_t)t) * 0.5;   // symmetrised in (i, j)
      dd& e1 = Fl[(((size_t)owner * n + lm) * n + la) * n + lb];
      e1 = e1 + half;
      dd& e2 = Fl[(((size_t)owner * n + lm) * n + lb) * n + la];
      e2 = e2 + half;
    }
  }
  // Gauss tables (pairs; Pgenx / Pgeny column-major (p+1, Q) -> [g][q])
  const int Q = D->Q, G = D->G;
  std::vector<dd> Pg((size_t)n * Q), Pgy((size_t)n * G), xv(Q), wv(Q), yv(G), wyv(G), inv1(C), inv2(C);
  std::vector<int> ycell(G);
  if (Q > 0) {
    for (int g = 0; g < n; ++g) {
      for (int q = 0; q < Q; ++q) Pg[(size_t)g * Q + q] = ddh_pair(D->Pgenx, (size_t)g + (size_t)n * q);
      for (int q = 0; q < G; ++q) Pgy[(size_t)g * G + q] = ddh_pair(D->Pgeny, (size_t)g + (size_t)n * q);
    }
    for (int q = 0; q < Q; ++q) { xv[q] = ddh_pair(D->x, q); wv[q] = ddh_pair(D->w, q); }
    for (int q = 0; q < G; ++q) {
      yv[q] = ddh_pair(D->y, q); wyv[q] = ddh_pair(D->wy, q);
      const int64_t cc = D->ycell[q];
      if (cc < 1 || cc > C) { ctx->err = "ycell out of range"; aks_disc_destroy(disc); return AKS_EINVAL; }
      ycell[q] = (int)cc - 1;
    }
    for (int c = 0; c < C; ++c) { inv1[c] = ddh_pair(D->invshifts, 2 * (size_t)c); inv2[c] = ddh_pair(D->invshifts, 2 * (size_t)c + 1); }
  }
  int rc;
#define UPD(vec, field)                                                                          \
  if ((rc = dev_upload(ctx, disc->allocs, vec, &dv.field)) != AKS_OK) { aks_disc_destroy(disc); return rc; }
  UPD(Al, A_loc) UPD(M0l, M0_loc) UPD(M1l, Mm1_loc) UPD(M2l, Mm2_loc) UPD(Fl, F_loc) UPD(l2g, l2g)
  UPD(Pg, Pg) UPD(Pgy, Pgy) UPD(xv, xq) UPD(wv, wq) UPD(yv, y) UPD(wyv, wy) UPD(ycell, ycell) UPD(inv1, inv1) UPD(inv2, inv2)
#undef UPD
  if ((rc = dev_zeros<dd>(ctx, disc->allocs, (size_t)C * dd_npk(nb), &dv.Afac)) != AKS_OK) { aks_disc_destroy(disc); return rc; }
  if ((rc = dev_zeros<dd>(ctx, disc->allocs, (size_t)2 * DD_NT, &dv.Atri)) != AKS_OK) { aks_disc_destroy(disc); return rc; }
  ddk_setup<<<1, DD_NT, 0, ctx->stream>>>(dv);
  ctx->launches += 1;
  {
    cudaError_t e_ = cudaStreamSynchronize(ctx->stream);
    if (e_ == cudaSuccess) e_ = cudaGetLastError();
    if (e_ != cudaSuccess) { ctx->err = std::string("ddk_setup -> ") + cudaGetErrorString(e_); aks_disc_destroy(disc); return AKS_ECUDA; }
  }
  *out = disc;
  return AKS_OK;
}

static int dd_batch_reset(aks_batch* bt) {
  aks_ctx* ctx = bt->disc->ctx;
  DDBatchDev& b = bt->dd->dev;
  const DDDiscDev& dv = bt->disc->dd->dev;
  const size_t B = bt->B, Nh = dv.Nh;
  CK(cudaMemsetAsync(b.D, 0, sizeof(dd) * B * Nh * Nh, ctx->stream));
  CK(cudaMemsetAsync(b.Bv, 0, sizeof(dd) * B * Nh, ctx->stream));
  CK(cudaMemsetAsync(b.Wv, 0, sizeof(dd) * B * Nh, ctx->stream));
  CK(cudaMemsetAsync(b.E, 0, sizeof(dd) * B * 5, ctx->stream));
  CK(cudaMemsetAsync(b.niter, 0, sizeof(int) * B, ctx->stream));
  CK(cudaMemsetAsync(b.status, 0, sizeof(int) * B, ctx->stream));
  CK(cudaMemsetAsync(b.warm, 0, sizeof(int) * B, ctx->stream));
  CK(cudaMemsetAsync(b.eig_fail, 0, sizeof(int) * B, ctx->stream));
  if (bt->dd->any_xc) {
    CK(cudaMemsetAsync(b.rho_q, 0, sizeof(dd) * B * dv.C * dv.Q, ctx->stream));
    CK(cudaMemsetAsync(b.rho_y, 0, sizeof(dd) * B * dv.G, ctx->stream));
  }
  CK(cudaMemsetAsync(b.rec, 0, sizeof(dd) * B * 7, ctx->stream));
  CK(cudaMemsetAsync(b.rec64, 0, sizeof(aks_iter_record) * B, ctx->stream));
  std::vector<dd> tv(B, dd_make(b.tinit));
  CK(cudaMemcpyAsync(b.t, tv.data(), sizeof(dd) * B, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemsetAsync(b.redo, 0, sizeof(int) * B, ctx->stream));
  ddk_window_set<<<(int)((B * dv.L + 127) / 128), 128, 0, ctx->stream>>>(b, 1, -1);
  LAUNCH_CHECK();
  bt->dd->eig_dirty = false;
  CK(cudaStreamSynchronize(ctx->stream));
  return AKS_OK;
}

static int dd_batch_create(aks_disc* disc, int32_t nprob, const double* Z, const double* N, const double* hartree,
                           const int32_t* ex_id, const int32_t* ec_id, const int32_t* lh_p, const int32_t* nh_p,
                           aks_batch** out) {
  aks_ctx* ctx = disc->ctx;
  const DDDiscDev& dv = disc->dd->dev;
  std::vector<int> ex(nprob, 0), ec(nprob, 0);
  for (int i = 0; i < nprob; ++i) {
    if (ex_id) ex[i] = ex_id[i];
    if (ec_id) ec[i] = ec_id[i];
    if ((ex[i] != AKS_XC_NONE && ex[i] != AKS_LDA_X) || (ec[i] != AKS_XC_NONE && ec[i] != AKS_LDA_C_PW)) {
      ctx->err = "unsupported functional id (LDA only: ex in {0,1}, ec in {0,12})"; return AKS_EINVAL;
    }
    if ((ex[i] || ec[i]) && dv.Q == 0) {
      ctx->err = "dtype=1 (Double64): this discretization was created without Gauss tables (Q = 0), a functional "
                 "needs them";
      return AKS_EINVAL;
    }
  }
  aks_batch* bt = new aks_batch();
  bt->disc = disc;
  bt->B = nprob;
  bt->dd = new DDBatchHost();
  DDBatchDev& b = bt->dd->dev;
  b.B = nprob; b.L = dv.L; b.nh = dv.nh;
  bt->hZ.assign(Z, Z + nprob);
  bt->hN.assign(N, N + nprob);
  bt->hHar.assign(nprob, 1.0);
  if (hartree) bt->hHar.assign(hartree, hartree + nprob);
  bt->hLp.assign(nprob, dv.L);
  bt->hnhp.assign(nprob, dv.nh);
  for (int i = 0; i < nprob; ++i) {
    if (lh_p) bt->hLp[i] = lh_p[i] + 1;
    if (nh_p) bt->hnhp[i] = nh_p[i];
    if (bt->hLp[i] < 1 || bt->hLp[i] > dv.L || bt->hnhp[i] < 1 || bt->hnhp[i] > dv.nh || !(N[i] > 0) || !(Z[i] > 0)) {
      ctx->err = "bad per-problem lh/nh/Z/N"; aks_batch_destroy(bt); return AKS_EINVAL;
    }
  }
  int rc;
#define UPB(vec, field) if ((rc = dev_upload(ctx, bt->allocs, vec, &b.field)) != AKS_OK) { aks_batch_destroy(bt); return rc; }
  UPB(bt->hZ, Z) UPB(bt->hN, N) UPB(bt->hHar, hartree) UPB(bt->hLp, Lp) UPB(bt->hnhp, nhp) UPB(ex, ex) UPB(ec, ec)
#undef UPB
  const size_t B = nprob, Nh = dv.Nh, C = dv.C, L = dv.L, nh = dv.nh, n = dv.n, Q = dv.Q, G = dv.G;
  {
    const size_t need = sizeof(dd) * C * dd_npk((int)n - 2);
    bt->dd->eig_smem = (need <= 200 * 1024) ? need : 0;
    if (bt->dd->eig_smem > 40 * 1024)
      CK(cudaFuncSetAttribute(ddk_eig<DD_EIG_LN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bt->dd->eig_smem));
  }
  bool any_xc = false;
  for (int i = 0; i < nprob; ++i) any_xc = any_xc || ex[i] || ec[i];
  bt->dd->any_xc = any_xc;
  const size_t xq = any_xc ? Q : 0, xg = any_xc ? G : 0;   // batches without a functional hold no node arrays
  b.nparts = (int)std::min<size_t>((Nh * Nh + 255) / 256, 148);
#define ZB(T, field, count) if ((rc = dev_zeros<T>(ctx, bt->allocs, (count), &b.field)) != AKS_OK) { aks_batch_destroy(bt); return rc; }
  ZB(dd, D, B * Nh * Nh) ZB(dd, Bv, B * Nh) ZB(dd, Wv, B * Nh) ZB(dd, E, B * 5) ZB(dd, t, B) ZB(int, niter, B)
  ZB(int, status, B) ZB(int, warm, B) ZB(dd, H, B * L * C * n * n) ZB(dd, eps, B * L * nh) ZB(dd, U, B * L * nh * Nh)
  ZB(dd, vtmp, B * L * nh * Nh) ZB(dd, ekin_orb, B * L * nh) ZB(dd, ecou_orb, B * L * nh) ZB(dd, occ, B * L * nh)
  ZB(int, eig_fail, B) ZB(double, nfix, B * L * nh) ZB(int, kact, B * L) ZB(int, kdone, B * L) ZB(int, redo, B) ZB(int, eig_cold, B) ZB(long long, eig_cyc, B * L * nh * 6) ZB(int, eig_code, B * L * nh) ZB(dd, scratch, (bt->dd->eig_smem ? 1 : B * L * nh * C * dd_npk((int)n - 2))) ZB(int, norb, B) ZB(int, orb_lk, B * DD_MAXORB)
  ZB(dd, orb_n0, B * DD_MAXORB) ZB(dd, orb_n1, B * DD_MAXORB) ZB(dd, orb_n, B * DD_MAXORB) ZB(int, degen, B)
  ZB(dd, bloc, B * DD_MAXORB * C * n) ZB(dd, borb, B * DD_MAXORB * Nh) ZB(dd, worb, B * DD_MAXORB * Nh)
  ZB(dd, Bnew, B * Nh) ZB(dd, Wnew, B * Nh) ZB(dd, Enew, B * 5) ZB(dd, tmix, B) ZB(dd, tnew, B)
  ZB(dd, norm_part, B * b.nparts) ZB(dd, rec, B * 7) ZB(aks_iter_record, rec64, B)
  ZB(dd, rho_q, B * C * xq) ZB(dd, rho_y, B * xg) ZB(dd, fxc, B * C * xq) ZB(dd, rho_q_new, B * 2 * C * xq)
  ZB(dd, rho_y_new, B * 2 * xg) ZB(dd, corr_part, B * 2 * C) ZB(dd, td, B)
#undef ZB
  b.aufbau_kind = 0; b.temp = 0.0;
  b.scftol = 1e-10; b.tinit = 0.6; b.frozen_t = 0; b.max_degen = 2; b.degen_tol = 1e-3;
  b.abstol_ls = b.reltol_ls = 4.930380657631324e-28; b.maxiter_ls = 100;
  {
    const char* wenv = getenv("AKS_DD_EIG_WINDOW");
    b.eig_window = wenv ? atoi(wenv) : 1;
  }   // 1e4 eps(Double64), oda/types.jl:39-40
  // the generic driver (aks_scf_run, fetch_records) reads the records and status words through the Float64 view
  bt->dev.B = nprob; bt->dev.s = 1; bt->dev.Lmax = dv.L; bt->dev.nhmax = dv.nh;
  bt->dev.rec = b.rec64;
  bt->dev.status = b.status;
  {
    const size_t need = 2 * sizeof(aks_iter_record) * B;
    CK(cudaMallocHost((void**)&bt->h_rec, need));
    bt->h_rec_bytes = need;
  }
  // at most this many orbitals carry weight: every level of every channel
  bt->dd->max_orb_launch = (int)std::min<size_t>(L * nh, DD_MAXINV);
  *out = bt;
  return dd_batch_reset(bt);
}

static int dd_set_oda(aks_batch* bt, double scftol, double tinit, int frozen_t, int max_degen, double degen_tol,
                      int maxiter_ls, double abstol_ls, double reltol_ls) {
  aks_ctx* ctx = bt->disc->ctx;
  DDBatchDev& b = bt->dd->dev;
  b.scftol = scftol; b.tinit = tinit; b.frozen_t = frozen_t; b.max_degen = max_degen; b.degen_tol = degen_tol;
  b.maxiter_ls = maxiter_ls; b.abstol_ls = abstol_ls; b.reltol_ls = reltol_ls;
  std::vector<dd> tv(bt->B, dd_make(tinit));
  CK(cudaMemcpyAsync(b.t, tv.data(), sizeof(dd) * bt->B, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return AKS_OK;
}

static int dd_step_launch(aks_batch* bt) {
  aks_ctx* ctx = bt->disc->ctx;
  const DDDiscDev& dv = bt->disc->dd->dev;
  const DDBatchDev& b = bt->dd->dev;
  cudaStream_t st = ctx->stream;
  const int B = bt->B, no = bt->dd->max_orb_launch;
  CK(cudaMemsetAsync(b.eig_fail, 0, sizeof(int) * B, st));
  CK(cudaMemsetAsync(b.eig_cold, 0, sizeof(int) * B, st));
  CK(cudaMemsetAsync(b.eig_cyc, 0, sizeof(long long) * (size_t)B * b.L * b.nh * 6, st));
  if (bt->dd->any_xc) {
    ddk_xc<<<dim3(dv.C, B), 128, 0, st>>>(dv, b);
    LAUNCH_CHECK();
  }
  ddk_H<<<dim3(dv.C, B), 256, 0, st>>>(dv, b);
  LAUNCH_CHECK();
  ddk_eig<DD_EIG_LN><<<dim3(dv.nh, dv.L, B), DD_EIG_LN * DD_NT, bt->dd->eig_smem, st>>>(dv, b, bt->dd->eig_smem ? 1 : 0, 0);
  LAUNCH_CHECK();
  ddk_aufbau<<<(B + 31) / 32, 32, 0, st>>>(dv, b, 0);
  LAUNCH_CHECK();
  if (b.eig_window) {   // problems whose window could not be certified: the remaining levels, then the aufbau again
    ddk_eig<DD_EIG_LN><<<dim3(dv.nh, dv.L, B), DD_EIG_LN * DD_NT, bt->dd->eig_smem, st>>>(dv, b, bt->dd->eig_smem ? 1 : 0, 1);
    LAUNCH_CHECK();
    ddk_aufbau<<<(B + 31) / 32, 32, 0, st>>>(dv, b, 1);
    LAUNCH_CHECK();
    bt->dd->eig_dirty = true;
  }
  ddk_orbB<<<dim3(dv.C, no, B), 32, 0, st>>>(dv, b);
  LAUNCH_CHECK();
  ddk_poisson<<<dim3(no, B), DD_NT, 0, st>>>(dv, b);
  LAUNCH_CHECK();
  if (bt->dd->any_xc) {
    ddk_rho_nodes<<<dim3(dv.C, B), 128, 0, st>>>(dv, b);
    LAUNCH_CHECK();
    ddk_rho_grid<<<dim3((dv.G + 127) / 128, B), 128, 0, st>>>(dv, b);
    LAUNCH_CHECK();
  }
  ddk_decide<<<dim3(DD_DECIDE_CL, B), DD_DECIDE_NT, 0, st>>>(dv, b);
  LAUNCH_CHECK();
  ddk_mix<<<dim3(b.nparts, B), 256, 0, st>>>(dv, b);
  LAUNCH_CHECK();
  ddk_finish<<<B, 128, 0, st>>>(dv, b);
  LAUNCH_CHECK();
  return AKS_OK;
}

// outputs in the reference's layouts, every number an interleaved (hi, lo) pair:
// D (Nh,Nh)  U (Nh,nh,L)  eps, n (L,nh)  W (Nh)
static int dd_get_state(aks_batch* bt, int prob, double* D, double* U, double* eps, double* nocc, double* W) {
  aks_ctx* ctx = bt->disc->ctx;
  const DDDiscDev& dv = bt->disc->dd->dev;
  const DDBatchDev& b = bt->dd->dev;
  const size_t Nh = dv.Nh, L = dv.L, nh = dv.nh;
  if ((U || eps) && bt->dd->eig_dirty) {
    // the caller receives all nh eigenpairs of the last Hamiltonian: solve the levels outside the window now
    ddk_eig<DD_EIG_LN><<<dim3(dv.nh, dv.L, bt->B), DD_EIG_LN * DD_NT, bt->dd->eig_smem, ctx->stream>>>(dv, b, bt->dd->eig_smem ? 1 : 0, 2);
    LAUNCH_CHECK();
    ddk_window_set<<<(int)((bt->B * dv.L + 127) / 128), 128, 0, ctx->stream>>>(b, 0, 1);
    LAUNCH_CHECK();
    bt->dd->eig_dirty = false;
  }
  if (D) CK(cudaMemcpyAsync(D, b.D + (size_t)prob * Nh * Nh, sizeof(dd) * Nh * Nh, cudaMemcpyDeviceToHost, ctx->stream));
  if (U) CK(cudaMemcpyAsync(U, b.U + (size_t)prob * L * nh * Nh, sizeof(dd) * L * nh * Nh, cudaMemcpyDeviceToHost, ctx->stream));
  std::vector<dd> he(L * nh), hn(L * nh), hw(Nh);
  if (eps) CK(cudaMemcpyAsync(he.data(), b.eps + (size_t)prob * L * nh, sizeof(dd) * L * nh, cudaMemcpyDeviceToHost, ctx->stream));
  if (nocc) CK(cudaMemcpyAsync(hn.data(), b.occ + (size_t)prob * L * nh, sizeof(dd) * L * nh, cudaMemcpyDeviceToHost, ctx->stream));
  if (W) CK(cudaMemcpyAsync(hw.data(), b.Wv + (size_t)prob * Nh, sizeof(dd) * Nh, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  for (size_t l = 0; l < L; ++l)
    for (size_t k = 0; k < nh; ++k) {
      const size_t o = l + L * k, i = l * nh + k;
      if (eps) { eps[2 * o] = he[i].hi; eps[2 * o + 1] = he[i].lo; }
      if (nocc) { nocc[2 * o] = hn[i].hi; nocc[2 * o + 1] = hn[i].lo; }
    }
  if (W)
    for (size_t i = 0; i < Nh; ++i) {
      const dd v = hw[i] * bt->hHar[prob];
      W[2 * i] = v.hi; W[2 * i + 1] = v.lo;
    }
  return AKS_OK;
}

extern "C" int aks_get_records_dd(aks_batch* bt, double* out) {
  if (!bt || !out) return AKS_EINVAL;
  GUARD(bt->disc->ctx);
  aks_ctx* ctx = bt->disc->ctx;
  if (!bt->dd) { ctx->err = "aks_get_records_dd: not a Double64 batch"; return AKS_EINVAL; }
  CK(cudaMemcpyAsync(out, bt->dd->dev.rec, sizeof(dd) * 7 * bt->B, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return AKS_OK;
}

static void dd_disc_free(aks_disc* disc) { delete disc->dd; disc->dd = nullptr; }
static void dd_batch_free(aks_batch* bt) { delete bt->dd; bt->dd = nullptr; }

// test hook: pointwise vxc / exc per particle of the Double64 functionals; rho, vrho, zk: npts (hi, lo) pairs
extern "C" int aks_dd_xc_eval(aks_ctx* ctx, int32_t npts, const double* rho, int32_t ex_id, int32_t ec_id,
                              double* vrho, double* zk) {
  if (!ctx || npts < 1 || !rho || !vrho || !zk) return AKS_EINVAL;
  GUARD(ctx);
  dd *dr = nullptr, *dv_ = nullptr, *dz = nullptr;
  const size_t bytes = sizeof(dd) * (size_t)npts;
  CK(cudaMallocAsync((void**)&dr, bytes, ctx->stream));
  CK(cudaMallocAsync((void**)&dv_, bytes, ctx->stream));
  CK(cudaMallocAsync((void**)&dz, bytes, ctx->stream));
  CK(cudaMemcpyAsync(dr, rho, bytes, cudaMemcpyHostToDevice, ctx->stream));
  ddk_xc_eval<<<(npts + 127) / 128, 128, 0, ctx->stream>>>(dr, npts, ex_id, ec_id, dv_, dz);
  LAUNCH_CHECK();
  CK(cudaMemcpyAsync(vrho, dv_, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(zk, dz, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  cudaFreeAsync(dr, ctx->stream); cudaFreeAsync(dv_, ctx->stream); cudaFreeAsync(dz, ctx->stream);
  return AKS_OK;
}

static int dd_eig_cold(aks_batch* bt, int32_t* counts) {
  aks_ctx* ctx = bt->disc->ctx;
  if (getenv("AKS_DD_EIG_CYCLES")) {   // debug: average cycle split of the CTAs of the last step (thread 0's clock)
    const DDBatchDev& b = bt->dd->dev;
    std::vector<long long> h((size_t)bt->B * b.L * b.nh * 6);
    CK(cudaMemcpyAsync(h.data(), b.eig_cyc, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    double sum[6] = {0, 0, 0, 0, 0, 0}, mx = 0; int cnt = 0;
    for (size_t i = 0; i < h.size() / 6; ++i) {
      if (h[i * 6 + 4] <= 0) continue;
      for (int q = 0; q < 6; ++q) sum[q] += (double)h[i * 6 + q];
      mx = std::max(mx, (double)h[i * 6 + 4]);
      ++cnt;
    }
    if (cnt) fprintf(stderr, "[dd eig cycles] CTAs %d  avg: dd-factor %.0f  dd-solve %.0f  matvec %.0f  total %.0f  (dd factorisations %.2f)  max total %.0f\n",
                     cnt, sum[0] / cnt, sum[1] / cnt, sum[2] / cnt, sum[4] / cnt, sum[5] / cnt, mx);
  }
  if (getenv("AKS_DD_EIG_CODES")) {   // debug: per-pair reason codes of problem 0 instead ([L][nh] words)
    const DDBatchDev& b = bt->dd->dev;
    std::vector<int> h((size_t)b.L * b.nh);
    CK(cudaMemcpyAsync(h.data(), b.eig_code, sizeof(int) * h.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < h.size(); ++i) if (h[i]) fprintf(stderr, "[dd eig] l=%zu k=%zu code=%d cl=%d ch=%d\n", i / b.nh, i % b.nh, h[i] & 0xff, (h[i] >> 8) & 0xff, (h[i] >> 16) & 0xff);
  }
  CK(cudaMemcpyAsync(counts, bt->dd->dev.eig_cold, sizeof(int) * bt->B, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return AKS_OK;
}

// aks_batch_set_eig_window on a Double64 batch: margin >= 0 switches the (certified) eigenpair window on,
// margin < 0 makes every step solve all nh levels; the window size itself is chosen by ddk_aufbau
static int dd_set_window(aks_batch* bt, int32_t margin) {
  aks_ctx* ctx = bt->disc->ctx;
  DDBatchDev& b = bt->dd->dev;
  b.eig_window = (margin >= 0 && b.aufbau_kind == 0) ? 1 : 0;
  ddk_window_set<<<(int)(((size_t)bt->B * b.L + 127) / 128), 128, 0, ctx->stream>>>(b, 1, 0);
  LAUNCH_CHECK();
  CK(cudaStreamSynchronize(ctx->stream));
  return AKS_OK;
}

// aks_batch_set_aufbau on a Double64 batch (kinds as in the Float64 path; nfix per problem (L, nh) column-major,
// Float64 values: occupations are small integers or user-given fractions)
static int dd_set_aufbau(aks_batch* bt, int32_t kind, double temp, const double* nfix) {
  aks_ctx* ctx = bt->disc->ctx;
  DDBatchDev& b = bt->dd->dev;
  if (kind == 2) {
    const size_t L = b.L, nh = b.nh, per = L * nh;
    std::vector<double> h((size_t)bt->B * per, 0.0);
    for (size_t p = 0; p < (size_t)bt->B; ++p)
      for (size_t l = 0; l < L; ++l)
        for (size_t k = 0; k < nh; ++k) h[p * per + l * nh + k] = nfix[p * per + l + L * k];
    CK(cudaMemcpyAsync(b.nfix, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  b.aufbau_kind = kind;
  b.temp = temp;
  if (kind != 0) b.eig_window = 0;
  ddk_window_set<<<(int)(((size_t)bt->B * b.L + 127) / 128), 128, 0, ctx->stream>>>(b, 1, 0);
  LAUNCH_CHECK();
  CK(cudaStreamSynchronize(ctx->stream));
  return AKS_OK;
}


// ------------------------------------------------------------------------------------ Double64 tables on the device
// SURVEY 8(f) rank 1.  Results in caller-owned host buffers of (hi, lo) pairs -- exactly what aks_disc_create
// (dtype = 1) consumes, so a caller that has no Double64 tables of its own builds them here.
extern "C" int aks_dd_gauss_legendre(aks_ctx* ctx, int32_t n, double* x, double* w) {
  if (!ctx || n < 1 || !x || !w) return AKS_EINVAL;
  GUARD(ctx);
  dd *dx = nullptr, *dw = nullptr;
  CK(cudaMallocAsync((void**)&dx, sizeof(dd) * n, ctx->stream));
  CK(cudaMallocAsync((void**)&dw, sizeof(dd) * n, ctx->stream));
  ddk_tab_gauss<<<(n + 63) / 64, 64, 0, ctx->stream>>>(n, dx, dw);
  LAUNCH_CHECK();
  CK(cudaMemcpyAsync(x, dx, sizeof(dd) * n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(w, dw, sizeof(dd) * n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  cudaFreeAsync(dx, ctx->stream); cudaFreeAsync(dw, ctx->stream);
  return AKS_OK;
}

extern "C" int aks_dd_build_operators(aks_ctx* ctx, int32_t p, int32_t C, const double* mesh, const int32_t* present,
                                      int32_t nq, double* M0_loc, double* A_loc, double* Mm1_loc, double* Mm2_loc,
                                      double* F_loc) {
  if (!ctx || !mesh || !present || !M0_loc || !A_loc || !Mm1_loc || !Mm2_loc || !F_loc) return AKS_EINVAL;
  if (p < 1 || p >= DDT_PMAX || C < 1 || nq < 1) { ctx->err = "aks_dd_build_operators: need 1 <= p <= 29, C >= 1, nq >= 1"; return AKS_EINVAL; }
  GUARD(ctx);
  const size_t n = (size_t)p + 1, n2 = (size_t)C * n * n, n3 = n2 * n;
  double* dmesh = nullptr; int* dpres = nullptr;
  dd *dx, *dw, *dG, *ddG, *dscr, *dM0, *dA, *dM1, *dM2, *dF;
  std::vector<void*> tmp;
  auto alloc = [&](void** ptr, size_t bytes) -> cudaError_t { cudaError_t e = cudaMallocAsync(ptr, bytes, ctx->stream); if (e == cudaSuccess) tmp.push_back(*ptr); return e; };
  CK(alloc((void**)&dmesh, sizeof(double) * (C + 1)));
  CK(alloc((void**)&dpres, sizeof(int) * C * n));
  CK(alloc((void**)&dx, sizeof(dd) * nq)); CK(alloc((void**)&dw, sizeof(dd) * nq));
  CK(alloc((void**)&dG, sizeof(dd) * n * nq)); CK(alloc((void**)&ddG, sizeof(dd) * n * nq));
  CK(alloc((void**)&dscr, sizeof(dd) * (size_t)C * 3 * nq));
  CK(alloc((void**)&dM0, sizeof(dd) * n2)); CK(alloc((void**)&dA, sizeof(dd) * n2));
  CK(alloc((void**)&dM1, sizeof(dd) * n2)); CK(alloc((void**)&dM2, sizeof(dd) * n2));
  CK(alloc((void**)&dF, sizeof(dd) * n3));
  CK(cudaMemcpyAsync(dmesh, mesh, sizeof(double) * (C + 1), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(dpres, present, sizeof(int) * C * n, cudaMemcpyHostToDevice, ctx->stream));
  ddk_tab_gauss<<<(nq + 63) / 64, 64, 0, ctx->stream>>>(nq, dx, dw);
  LAUNCH_CHECK();
  ddk_tab_genvals<<<(nq + 63) / 64, 64, 0, ctx->stream>>>(p, nq, dx, dG, ddG);
  LAUNCH_CHECK();
  ddk_tab_cells<<<C, 256, 0, ctx->stream>>>(p, C, nq, dmesh, dpres, dx, dw, dG, ddG, dscr, dM0, dA, dM1, dM2, dF);
  LAUNCH_CHECK();
  CK(cudaMemcpyAsync(M0_loc, dM0, sizeof(dd) * n2, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(A_loc, dA, sizeof(dd) * n2, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(Mm1_loc, dM1, sizeof(dd) * n2, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(Mm2_loc, dM2, sizeof(dd) * n2, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(F_loc, dF, sizeof(dd) * n3, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  for (void* q : tmp) cudaFreeAsync(q, ctx->stream);
  return AKS_OK;
}
