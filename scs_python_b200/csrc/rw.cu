// rw.cu -- problem data files and the CSV trace header (host code only).
//
// Reference behaviour: S/src/rw.c.  SCS(write_data) (rw.c:240-260) dumps cone, data and settings of
// a problem in the library's native binary layout; SCS(read_data) (rw.c:262-315) reads such a file
// back, converting the integer width when the file was written by a build with another scs_int;
// SCS(log_data_to_csv) (rw.c:317-476) appends one row of solver state per iteration.  The file
// format is the drop-in surface here: files written by either library must be readable by the other
// (tests/test_rw.py checks both directions byte for byte against the compiled reference).
//
//   file := u32 sizeof(scs_int) | u32 sizeof(scs_float) | u32 len | version[len]
//           | cone | data | settings
//   cone := z l bsize | bl[bsize-1] bu[bsize-1] | qsize q[] | ssize s[] | ep ed | psize p[]
//           (complex PSD sizes `cs` are not part of the format, as in the reference)
//   data := m n | b[m] c[n] | A | has_p [P]          matrix := m n | p[n+1] | x[nnz] | i[nnz]
//   settings := normalize scale rho_x max_iters eps_abs eps_rel eps_infeas alpha verbose warm_start(=0)
//               acceleration_lookback acceleration_interval acceleration_type_1
//               acceleration_regularization acceleration_relaxation adaptive_scale
//           (time_limit_secs and the two file names are not stored)
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"

extern "C" const char *scs_version(void);

namespace b200 {
namespace {

struct Writer {
  FILE *f;
  bool ok = true;
  template <class T>
  void put(const T *p, size_t n) {
    if (n && fwrite(p, sizeof(T), n, f) != n) ok = false;
  }
  template <class T>
  void one(const T &v) { put(&v, 1); }
  void matrix(const ScsMatrix *M) {
    const scs_int nnz = M->p[M->n];
    one(M->m); one(M->n);
    put(M->p, (size_t)M->n + 1);
    put(M->x, (size_t)nnz);
    put(M->i, (size_t)nnz);
  }
};

struct Reader {
  FILE *f;
  size_t int_sz;
  bool ok = true;
  // lenient: a short read leaves the (zero-initialised) destination as it is and is not an error.  The
  // reference reads every field this way (rw.c:60-101: checked_fread only prints); it matters for the
  // settings block, which grew over the versions -- S/test/problems/random_prob (written by 3.0.0) ends
  // three fields early and the reference's own tests solve it with those fields at zero.
  bool lenient = false;
  void fail(size_t n) {
    if (lenient) B200_PRINTF("Error: fread expected %lu items\n", (unsigned long)n);
    else ok = false;
  }
  // integers of the file's width -> scs_int
  void ints(scs_int *dst, size_t n) {
    if (!n) return;
    if (int_sz == sizeof(scs_int)) {
      if (fread(dst, sizeof(scs_int), n, f) != n) fail(n);
      return;
    }
    if (int_sz == 8) {
      std::vector<long long> tmp(n);
      if (fread(tmp.data(), 8, n, f) != n) { fail(n); return; }
      for (size_t k = 0; k < n; ++k) dst[k] = (scs_int)tmp[k];
    } else if (int_sz == 4) {
      std::vector<int> tmp(n);
      if (fread(tmp.data(), 4, n, f) != n) { fail(n); return; }
      for (size_t k = 0; k < n; ++k) dst[k] = (scs_int)tmp[k];
    } else {
      ok = false;
    }
  }
  scs_int one_int() { scs_int v = 0; ints(&v, 1); return v; }
  void floats(scs_float *dst, size_t n) {
    if (n && fread(dst, sizeof(scs_float), n, f) != n) {
      if (!lenient) B200_PRINTF("Error: fread expected %lu items\n", (unsigned long)n);
      fail(n);
    }
  }
  scs_float one_float() { scs_float v = 0; floats(&v, 1); return v; }
  ScsMatrix *matrix() {
    ScsMatrix *M = (ScsMatrix *)calloc(1, sizeof(ScsMatrix));
    M->m = one_int(); M->n = one_int();
    if (!ok || M->n < 0) { ok = false; return M; }
    M->p = (scs_int *)calloc((size_t)M->n + 1, sizeof(scs_int));
    ints(M->p, (size_t)M->n + 1);
    const scs_int nnz = ok ? M->p[M->n] : 0;
    if (nnz < 0) { ok = false; return M; }
    M->x = (scs_float *)calloc(nnz > 0 ? nnz : 1, sizeof(scs_float));
    M->i = (scs_int *)calloc(nnz > 0 ? nnz : 1, sizeof(scs_int));
    floats(M->x, (size_t)nnz);
    ints(M->i, (size_t)nnz);
    return M;
  }
};

void free_matrix(ScsMatrix *M) {
  if (!M) return;
  free(M->x); free(M->i); free(M->p); free(M);
}

}  // namespace
}  // namespace b200

using namespace b200;

// util.c SCS(free_data) / cones.c free_cone: everything scs_b200_read_data allocated
extern "C" void scs_b200_free_data(ScsData *d, ScsCone *k, ScsSettings *stgs) {
  if (d) {
    free(d->b); free(d->c);
    free_matrix(d->A); free_matrix(d->P);
    free(d);
  }
  if (k) {
    free(k->bu); free(k->bl); free(k->q); free(k->s); free(k->cs); free(k->p);
    free(k);
  }
  free(stgs);
}

// SCS(write_data), rw.c:240-260.  The target is stgs->write_data_filename.  Returns 0, < 0 on I/O errors
// (the reference prints and carries on; scs_init does the same with this return value).
extern "C" scs_int scs_b200_write_data(const ScsData *d, const ScsCone *k, const ScsSettings *stgs) {
  if (!d || !k || !stgs || !stgs->write_data_filename || !d->A) return -1;
  FILE *f = fopen(stgs->write_data_filename, "wb");
  if (!f) {
    B200_PRINTF("Error: could not open %s for writing\n", stgs->write_data_filename);
    return -1;
  }
  Writer w{f};
  const uint32_t isz = (uint32_t)sizeof(scs_int), fsz = (uint32_t)sizeof(scs_float);
  const char *ver = scs_version();
  const uint32_t vlen = (uint32_t)strlen(ver);
  w.one(isz); w.one(fsz); w.one(vlen);
  w.put(ver, vlen);
  // cone (rw.c:42-58)
  const size_t nb = k->bsize > 1 ? (size_t)k->bsize - 1 : 0;
  w.one(k->z); w.one(k->l); w.one(k->bsize);
  w.put(k->bl, nb); w.put(k->bu, nb);
  w.one(k->qsize); w.put(k->q, (size_t)k->qsize);
  w.one(k->ssize); w.put(k->s, (size_t)k->ssize);
  w.one(k->ep); w.one(k->ed);
  w.one(k->psize); w.put(k->p, (size_t)k->psize);
  // data (rw.c:208-220)
  w.one(d->m); w.one(d->n);
  w.put(d->b, (size_t)d->m); w.put(d->c, (size_t)d->n);
  w.matrix(d->A);
  const scs_int has_p = d->P ? 1 : 0;
  w.one(has_p);
  if (d->P) w.matrix(d->P);
  // settings (rw.c:136-157); warm_start is written as 0
  const scs_int zero = 0;
  w.one(stgs->normalize); w.one(stgs->scale); w.one(stgs->rho_x); w.one(stgs->max_iters);
  w.one(stgs->eps_abs); w.one(stgs->eps_rel); w.one(stgs->eps_infeas); w.one(stgs->alpha);
  w.one(stgs->verbose); w.one(zero);
  w.one(stgs->acceleration_lookback); w.one(stgs->acceleration_interval); w.one(stgs->acceleration_type_1);
  w.one(stgs->acceleration_regularization); w.one(stgs->acceleration_relaxation);
  w.one(stgs->adaptive_scale);
  const bool ok = w.ok;
  fclose(f);
  return ok ? 0 : -1;
}

// SCS(read_data), rw.c:262-315.  Allocates *d, *k, *stgs (release with scs_b200_free_data).
extern "C" scs_int scs_b200_read_data(const char *filename, ScsData **d, ScsCone **k, ScsSettings **stgs) {
  if (!filename || !d || !k || !stgs) return -1;
  *d = nullptr; *k = nullptr; *stgs = nullptr;
  errno = 0;
  FILE *f = fopen(filename, "rb");
  if (!f) {
    B200_PRINTF("Error reading file %s\n", filename);
    B200_PRINTF("errno:%i:%s\n", errno, strerror(errno));
    return -1;
  }
  B200_PRINTF("Reading data from %s\n", filename);
  uint32_t isz = 0, fsz = 0, vlen = 0;
  char ver[16];
  if (fread(&isz, 4, 1, f) != 1 || fread(&fsz, 4, 1, f) != 1) { fclose(f); return -1; }
  if (isz != (uint32_t)sizeof(scs_int))
    B200_PRINTF("Warning, sizeof(file int) is %lu, but scs expects sizeof(int) %lu. SCS will attempt to cast the data, "
                "which may be slow. This message can be avoided by recompiling with the correct flags.\n",
                (unsigned long)isz, (unsigned long)sizeof(scs_int));
  if (fsz != (uint32_t)sizeof(scs_float)) {
    B200_PRINTF("Error, sizeof(file float) is %lu, but scs expects sizeof(float) %lu, scs should be recompiled with "
                "the correct flags.\n", (unsigned long)fsz, (unsigned long)sizeof(scs_float));
    fclose(f);
    return -1;
  }
  if (fread(&vlen, 4, 1, f) != 1 || vlen >= sizeof(ver)) {
    B200_PRINTF("Error: file version string length %lu exceeds buffer size\n", (unsigned long)vlen);
    fclose(f);
    return -1;
  }
  if (fread(ver, 1, vlen, f) != vlen) { fclose(f); return -1; }
  ver[vlen] = '\0';
  if (strcmp(ver, scs_version()) != 0)
    B200_PRINTF("************************************************************\n"
                "Warning: SCS file version %s, this is SCS version %s.\n"
                "The file reading / writing logic might have changed.\n"
                "************************************************************\n", ver, scs_version());
  Reader r{f, (size_t)isz};
  // cone (rw.c:103-134)
  ScsCone *K = (ScsCone *)calloc(1, sizeof(ScsCone));
  K->z = r.one_int(); K->l = r.one_int(); K->bsize = r.one_int();
  if (r.ok && K->bsize > 1) {
    const size_t nb = (size_t)K->bsize - 1;
    K->bl = (scs_float *)calloc(nb, sizeof(scs_float));
    K->bu = (scs_float *)calloc(nb, sizeof(scs_float));
    r.floats(K->bl, nb); r.floats(K->bu, nb);
  }
  K->qsize = r.one_int();
  if (r.ok && K->qsize > 0) { K->q = (scs_int *)calloc(K->qsize, sizeof(scs_int)); r.ints(K->q, (size_t)K->qsize); }
  K->ssize = r.one_int();
  if (r.ok && K->ssize > 0) { K->s = (scs_int *)calloc(K->ssize, sizeof(scs_int)); r.ints(K->s, (size_t)K->ssize); }
  K->ep = r.one_int(); K->ed = r.one_int();
  K->psize = r.one_int();
  if (r.ok && K->psize > 0) { K->p = (scs_float *)calloc(K->psize, sizeof(scs_float)); r.floats(K->p, (size_t)K->psize); }
  // data (rw.c:222-238)
  ScsData *D = (ScsData *)calloc(1, sizeof(ScsData));
  D->m = r.one_int(); D->n = r.one_int();
  if (r.ok && D->m >= 0 && D->n >= 0) {
    D->b = (scs_float *)calloc(D->m > 0 ? D->m : 1, sizeof(scs_float));
    D->c = (scs_float *)calloc(D->n > 0 ? D->n : 1, sizeof(scs_float));
    r.floats(D->b, (size_t)D->m); r.floats(D->c, (size_t)D->n);
    D->A = r.matrix();
    // has_p: absent (end of file) in files of old versions -> no P (rw.c:235)
    scs_int has_p = 0;
    if (r.ok) {
      Reader probe{f, (size_t)isz};
      has_p = probe.one_int();
      if (!probe.ok) has_p = 0;
    }
    if (r.ok && has_p) D->P = r.matrix();
  } else {
    r.ok = false;
  }
  // settings (rw.c:159-180); fields missing at the end of an older file stay zero, as in the reference
  ScsSettings *S = (ScsSettings *)calloc(1, sizeof(ScsSettings));
  r.lenient = true;
  S->normalize = r.one_int(); S->scale = r.one_float(); S->rho_x = r.one_float(); S->max_iters = r.one_int();
  S->eps_abs = r.one_float(); S->eps_rel = r.one_float(); S->eps_infeas = r.one_float(); S->alpha = r.one_float();
  S->verbose = r.one_int(); S->warm_start = r.one_int();
  S->acceleration_lookback = r.one_int(); S->acceleration_interval = r.one_int(); S->acceleration_type_1 = r.one_int();
  S->acceleration_regularization = r.one_float(); S->acceleration_relaxation = r.one_float();
  S->adaptive_scale = r.one_int();
  fclose(f);
  if (!r.ok) {
    scs_b200_free_data(D, K, S);
    return -1;
  }
  *d = D; *k = K; *stgs = S;
  B200_PRINTF("Finished reading data.\n");
  return 0;
}

// Column names of the CSV trace, rw.c:333-402 (USE_LAPACK builds append five spectral-cone columns; this
// backend has no spectral cones and writes the 62 common columns).
extern "C" const char *scs_b200_csv_header(void) {
  return "iter,res_pri,res_dual,gap,x_nrm_inf,y_nrm_inf,s_nrm_inf,x_nrm_2,y_nrm_2,s_nrm_2,"
         "x_nrm_inf_normalized,y_nrm_inf_normalized,s_nrm_inf_normalized,x_nrm_2_normalized,y_nrm_2_normalized,"
         "s_nrm_2_normalized,ax_s_btau_nrm_inf,px_aty_ctau_nrm_inf,ax_s_btau_nrm_2,px_aty_ctau_nrm_2,res_infeas,"
         "res_unbdd_a,res_unbdd_p,pobj,dobj,tau,kap,res_pri_normalized,res_dual_normalized,gap_normalized,"
         "ax_s_btau_nrm_inf_normalized,px_aty_ctau_nrm_inf_normalized,ax_s_btau_nrm_2_normalized,"
         "px_aty_ctau_nrm_2_normalized,res_infeas_normalized,res_unbdd_a_normalized,res_unbdd_p_normalized,"
         "pobj_normalized,dobj_normalized,tau_normalized,kap_normalized,ax_nrm_inf,ax_s_nrm_inf,px_nrm_inf,"
         "aty_nrm_inf,xt_p_x,xt_p_x_tau,ctx,ctx_tau,bty,bty_tau,b_nrm_inf,c_nrm_inf,scale,diff_u_ut_nrm_2,"
         "diff_v_v_prev_nrm_2,diff_u_ut_nrm_inf,diff_v_v_prev_nrm_inf,aa_norm,accepted_accel_steps,"
         "rejected_accel_steps,time,";
}
