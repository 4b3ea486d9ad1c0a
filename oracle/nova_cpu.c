/* nova_cpu.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the CPU path nova-snark 0.23.0 runs for one NIFS fold step, as driven by
 * zero-savvy/vimz (hot loop entered at /root/reference/vimz/src/nova_snark_backend/folding.rs:35).
 * PARITY UNPINNED at the nova-snark boundary: the crate is an un-vendored dependency
 * (vimz/Cargo.toml:51, Cargo.lock:3577) and the reference has no golden vector for this path
 * (SURVEY.md section 8c).  This file is cross-checked against the independent big-int model in
 * oracle/pyref.py (tests/test_oracle.py); results are canonical values, so any correct
 * implementation is bit-identical.
 *
 * Restates:
 *   - 4 x 64-bit Montgomery field arithmetic (halo2curves 0.1.0 / pasta_curves 0.5.1 layout)
 *   - Jacobian group law (a = 0)
 *   - provider::cpu_best_multiexp / cpu_multiexp_serial  [A.6]: unsigned windows, c = ceil(ln n),
 *     one chunk per thread, buckets summed by the running-sum trick, chunk results added
 *   - R1CSShape::multiply_vec over COO triples, three matrices in parallel  [A.3]
 *   - R1CSShape::commit_T cross term  [A.4] and RelaxedR1CSWitness::fold  [A.5]
 * It is also bench.py's cpu_baseline / --impl reference leg ("port": the Rust crate cannot be built
 * here -- no cargo/rustc in the image).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fe;

typedef struct {
  fe p;          /* modulus */
  uint64_t inv;  /* -p^-1 mod 2^64 */
  fe one;        /* R mod p */
  fe r2;         /* R^2 mod p */
} field_t;

typedef struct {
  field_t fb;  /* coordinate field */
  field_t fs;  /* scalar field */
} curve_t;

static curve_t CURVES[4];
static int g_init = 0;

/* ---- 256-bit helpers ------------------------------------------------------------------------ */
static int fe_is_zero(const fe* a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static int fe_eq(const fe* a, const fe* b) {
  return ((a->l[0] ^ b->l[0]) | (a->l[1] ^ b->l[1]) | (a->l[2] ^ b->l[2]) | (a->l[3] ^ b->l[3])) == 0;
}
static int fe_geq(const fe* a, const fe* b) {
  for (int i = 3; i >= 0; i--) {
    if (a->l[i] > b->l[i]) return 1;
    if (a->l[i] < b->l[i]) return 0;
  }
  return 1;
}
static uint64_t add256(fe* r, const fe* a, const fe* b) {
  u128 c = 0;
  for (int i = 0; i < 4; i++) {
    c += (u128)a->l[i] + b->l[i];
    r->l[i] = (uint64_t)c;
    c >>= 64;
  }
  return (uint64_t)c;
}
static uint64_t sub256(fe* r, const fe* a, const fe* b) {
  uint64_t borrow = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a->l[i] - b->l[i] - borrow;
    r->l[i] = (uint64_t)d;
    borrow = (uint64_t)(d >> 64) & 1;
  }
  return borrow;
}

static void f_add(const field_t* F, fe* r, const fe* a, const fe* b) {
  fe t;
  uint64_t c = add256(&t, a, b);
  if (c || fe_geq(&t, &F->p)) sub256(&t, &t, &F->p);
  *r = t;
}
static void f_sub(const field_t* F, fe* r, const fe* a, const fe* b) {
  fe t;
  if (sub256(&t, a, b)) add256(&t, &t, &F->p);
  *r = t;
}
static void f_neg(const field_t* F, fe* r, const fe* a) {
  if (fe_is_zero(a)) { *r = *a; return; }
  sub256(r, &F->p, a);
}
/* Montgomery product a*b*R^-1 mod p (CIOS, 4 x 64-bit limbs) */
static void f_mul(const field_t* F, fe* r, const fe* a, const fe* b) {
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) {
      c += (u128)a->l[j] * b->l[i] + t[j];
      t[j] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[4] = (uint64_t)c;
    t[5] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * F->inv;
    c = (u128)m * F->p.l[0] + t[0];
    c >>= 64;
    for (int j = 1; j < 4; j++) {
      c += (u128)m * F->p.l[j] + t[j];
      t[j - 1] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[3] = (uint64_t)c;
    t[4] = t[5] + (uint64_t)(c >> 64);
  }
  fe out = {{t[0], t[1], t[2], t[3]}};
  if (t[4] || fe_geq(&out, &F->p)) sub256(&out, &out, &F->p);
  *r = out;
}
static void f_sqr(const field_t* F, fe* r, const fe* a) { f_mul(F, r, a, a); }
static void f_from_mont(const field_t* F, fe* r, const fe* a) {
  fe one = {{1, 0, 0, 0}};
  f_mul(F, r, a, &one);
}
static void f_inv(const field_t* F, fe* r, const fe* a) { /* a^(p-2) */
  fe e = F->p, two = {{2, 0, 0, 0}};
  sub256(&e, &e, &two);
  fe acc = F->one;
  for (int i = 255; i >= 0; i--) {
    f_sqr(F, &acc, &acc);
    if ((e.l[i >> 6] >> (i & 63)) & 1) f_mul(F, &acc, &acc, a);
  }
  *r = acc;
}

/* ---- constants ---------------------------------------------------------------------------- */
static void field_init(field_t* F, const uint64_t p[4]) {
  memcpy(F->p.l, p, 32);
  uint64_t inv = 1; /* Newton: inv = p^-1 mod 2^64 */
  for (int i = 0; i < 6; i++) inv *= 2 - p[0] * inv;
  F->inv = (uint64_t)0 - inv;
  /* R mod p by 256 doublings of 1; R^2 mod p by 256 more */
  fe x = {{1, 0, 0, 0}};
  for (int i = 0; i < 512; i++) {
    fe t;
    uint64_t c = add256(&t, &x, &x);
    if (c || fe_geq(&t, &F->p)) sub256(&t, &t, &F->p);
    x = t;
    if (i == 255) F->one = x;
  }
  F->r2 = x;
}
static void oracle_init(void) {
  if (g_init) return;
  static const uint64_t PALLAS_P[4] = {0x992d30ed00000001ull, 0x224698fc094cf91bull, 0x0ull, 0x4000000000000000ull};
  static const uint64_t VESTA_P[4] = {0x8c46eb2100000001ull, 0x224698fc0994a8ddull, 0x0ull, 0x4000000000000000ull};
  static const uint64_t BN_P[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
  static const uint64_t BN_R[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
  field_init(&CURVES[0].fb, PALLAS_P); field_init(&CURVES[0].fs, VESTA_P);
  field_init(&CURVES[1].fb, VESTA_P);  field_init(&CURVES[1].fs, PALLAS_P);
  field_init(&CURVES[2].fb, BN_P);     field_init(&CURVES[2].fs, BN_R);
  field_init(&CURVES[3].fb, BN_R);     field_init(&CURVES[3].fs, BN_P);
  g_init = 1;
}

/* ---- Jacobian group law (a = 0), identity Z = 0 --------------------------------------------- */
typedef struct { fe x, y, z; } jac;
typedef struct { fe x, y; } aff;

static void jac_set_identity(const field_t* F, jac* r) {
  memset(r, 0, sizeof(*r));
  r->y = F->one;
}
static int aff_is_identity(const aff* a) { return fe_is_zero(&a->x) && fe_is_zero(&a->y); }

static void jac_double(const field_t* F, jac* r, const jac* p) {
  if (fe_is_zero(&p->z) || fe_is_zero(&p->y)) { jac_set_identity(F, r); return; }
  fe a, b, c, d, e, f, t, x3, y3, z3;
  f_sqr(F, &a, &p->x);
  f_sqr(F, &b, &p->y);
  f_sqr(F, &c, &b);
  f_add(F, &t, &p->x, &b); f_sqr(F, &t, &t); f_sub(F, &t, &t, &a); f_sub(F, &t, &t, &c);
  f_add(F, &d, &t, &t);
  f_add(F, &e, &a, &a); f_add(F, &e, &e, &a);
  f_sqr(F, &f, &e);
  f_add(F, &t, &d, &d); f_sub(F, &x3, &f, &t);
  f_sub(F, &t, &d, &x3); f_mul(F, &y3, &e, &t);
  f_add(F, &t, &c, &c); f_add(F, &t, &t, &t); f_add(F, &t, &t, &t);
  f_sub(F, &y3, &y3, &t);
  f_mul(F, &z3, &p->y, &p->z); f_add(F, &z3, &z3, &z3);
  r->x = x3; r->y = y3; r->z = z3;
}
static void jac_add(const field_t* F, jac* r, const jac* p, const jac* q) {
  if (fe_is_zero(&p->z)) { *r = *q; return; }
  if (fe_is_zero(&q->z)) { *r = *p; return; }
  fe z1z1, z2z2, u1, u2, s1, s2, h, rr, hh, hhh, v, t, x3, y3, z3;
  f_sqr(F, &z1z1, &p->z); f_sqr(F, &z2z2, &q->z);
  f_mul(F, &u1, &p->x, &z2z2); f_mul(F, &u2, &q->x, &z1z1);
  f_mul(F, &s1, &p->y, &q->z); f_mul(F, &s1, &s1, &z2z2);
  f_mul(F, &s2, &q->y, &p->z); f_mul(F, &s2, &s2, &z1z1);
  if (fe_eq(&u1, &u2)) {
    if (fe_eq(&s1, &s2)) { jac_double(F, r, p); return; }
    jac_set_identity(F, r);
    return;
  }
  f_sub(F, &h, &u2, &u1); f_sub(F, &rr, &s2, &s1);
  f_sqr(F, &hh, &h); f_mul(F, &hhh, &h, &hh); f_mul(F, &v, &u1, &hh);
  f_sqr(F, &x3, &rr); f_sub(F, &x3, &x3, &hhh); f_add(F, &t, &v, &v); f_sub(F, &x3, &x3, &t);
  f_sub(F, &t, &v, &x3); f_mul(F, &y3, &rr, &t); f_mul(F, &t, &s1, &hhh); f_sub(F, &y3, &y3, &t);
  f_mul(F, &z3, &p->z, &q->z); f_mul(F, &z3, &z3, &h);
  r->x = x3; r->y = y3; r->z = z3;
}
static void jac_add_affine(const field_t* F, jac* r, const jac* p, const aff* q) {
  if (aff_is_identity(q)) { *r = *p; return; }
  if (fe_is_zero(&p->z)) { r->x = q->x; r->y = q->y; r->z = F->one; return; }
  fe z1z1, u2, s2, h, rr, hh, hhh, v, t, x3, y3, z3;
  f_sqr(F, &z1z1, &p->z);
  f_mul(F, &u2, &q->x, &z1z1);
  f_mul(F, &s2, &q->y, &p->z); f_mul(F, &s2, &s2, &z1z1);
  if (fe_eq(&p->x, &u2)) {
    if (fe_eq(&p->y, &s2)) { jac_double(F, r, p); return; }
    jac_set_identity(F, r);
    return;
  }
  f_sub(F, &h, &u2, &p->x); f_sub(F, &rr, &s2, &p->y);
  f_sqr(F, &hh, &h); f_mul(F, &hhh, &h, &hh); f_mul(F, &v, &p->x, &hh);
  f_sqr(F, &x3, &rr); f_sub(F, &x3, &x3, &hhh); f_add(F, &t, &v, &v); f_sub(F, &x3, &x3, &t);
  f_sub(F, &t, &v, &x3); f_mul(F, &y3, &rr, &t); f_mul(F, &t, &p->y, &hhh); f_sub(F, &y3, &y3, &t);
  f_mul(F, &z3, &p->z, &h);
  r->x = x3; r->y = y3; r->z = z3;
}
static void jac_to_affine(const field_t* F, aff* r, const jac* p) {
  if (fe_is_zero(&p->z)) { memset(r, 0, sizeof(*r)); return; }
  fe zi, zi2, zi3;
  f_inv(F, &zi, &p->z);
  f_sqr(F, &zi2, &zi);
  f_mul(F, &zi3, &zi2, &zi);
  f_mul(F, &r->x, &p->x, &zi2);
  f_mul(F, &r->y, &p->y, &zi3);
}

/* ---- cpu_multiexp_serial / cpu_best_multiexp  [A.6] ------------------------------------------- */
enum { B_NONE = 0, B_AFFINE = 1, B_PROJ = 2 };
typedef struct { int kind; aff a; jac j; } bucket_t;

static void bucket_add_assign(const field_t* F, bucket_t* b, const aff* other) {
  if (b->kind == B_NONE) { b->kind = B_AFFINE; b->a = *other; }
  else if (b->kind == B_AFFINE) {
    jac t; t.x = b->a.x; t.y = b->a.y; t.z = F->one;
    if (aff_is_identity(&b->a)) jac_set_identity(F, &t);
    jac_add_affine(F, &b->j, &t, other);
    b->kind = B_PROJ;
  } else {
    jac_add_affine(F, &b->j, &b->j, other);
  }
}
static void bucket_add_to(const field_t* F, const bucket_t* b, jac* acc) {
  if (b->kind == B_AFFINE) jac_add_affine(F, acc, acc, &b->a);
  else if (b->kind == B_PROJ) jac_add(F, acc, acc, &b->j);
}
static uint64_t get_at(int segment, int c, const fe* repr) {
  int skip = segment * c;
  if (skip >= 256) return 0;
  int limb = skip >> 6, off = skip & 63;
  uint64_t v = repr->l[limb] >> off;
  if (off && limb + 1 < 4) v |= repr->l[limb + 1] << (64 - off);
  return c >= 64 ? v : v & (((uint64_t)1 << c) - 1);
}
static int window_size(size_t n) {
  if (n < 4) return 1;
  if (n < 32) return 3;
  return (int)ceil(log((double)n));
}
/* scalars: canonical (non-Montgomery) representations */
static void cpu_multiexp_serial(const curve_t* C, const fe* repr, const aff* bases, size_t n, jac* acc) {
  const field_t* F = &C->fb;
  int c = window_size(n);
  int segments = 256 / c + 1;
  size_t nb = ((size_t)1 << c) - 1;
  bucket_t* buckets = (bucket_t*)malloc(nb * sizeof(bucket_t));
  jac_set_identity(F, acc);
  for (int seg = segments - 1; seg >= 0; seg--) {
    for (int k = 0; k < c; k++) jac_double(F, acc, acc);
    for (size_t b = 0; b < nb; b++) buckets[b].kind = B_NONE;
    for (size_t i = 0; i < n; i++) {
      uint64_t d = get_at(seg, c, &repr[i]);
      if (d != 0) bucket_add_assign(F, &buckets[d - 1], &bases[i]);
    }
    jac running;
    jac_set_identity(F, &running);
    for (size_t b = nb; b-- > 0;) {
      bucket_add_to(F, &buckets[b], &running);
      jac_add(F, acc, acc, &running);
    }
  }
  free(buckets);
}

typedef struct {
  const curve_t* C;
  const fe* scalars_mont;
  const aff* bases;
  size_t n;
  jac out;
} msm_job;

static void* msm_worker(void* arg) {
  msm_job* j = (msm_job*)arg;
  fe* repr = (fe*)malloc((j->n ? j->n : 1) * sizeof(fe));
  for (size_t i = 0; i < j->n; i++) f_from_mont(&j->C->fs, &repr[i], &j->scalars_mont[i]); /* coeffs.to_repr() */
  cpu_multiexp_serial(j->C, repr, j->bases, j->n, &j->out);
  free(repr);
  return NULL;
}

/* out = sum scalars[i] * bases[i]; Jacobian, Montgomery limbs.  nthreads plays rayon's
 * current_num_threads(): n > nthreads -> chunks of n / nthreads, else one serial call. */
int oracle_msm(int curve_id, const uint64_t* scalars_mont, const uint64_t* bases_aff, size_t n, int nthreads, uint64_t* out_jac) {
  if (curve_id < 0 || curve_id > 3) return -2;
  oracle_init();
  const curve_t* C = &CURVES[curve_id];
  if (nthreads < 1) nthreads = 1;
  jac acc;
  if (n > (size_t)nthreads) {
    size_t chunk = n / nthreads;
    size_t njobs = (n + chunk - 1) / chunk;
    msm_job* jobs = (msm_job*)malloc(njobs * sizeof(msm_job));
    pthread_t* th = (pthread_t*)malloc(njobs * sizeof(pthread_t));
    for (size_t k = 0; k < njobs; k++) {
      size_t s0 = k * chunk, len = (s0 + chunk <= n) ? chunk : n - s0;
      jobs[k].C = C;
      jobs[k].scalars_mont = (const fe*)scalars_mont + s0;
      jobs[k].bases = (const aff*)bases_aff + s0;
      jobs[k].n = len;
      pthread_create(&th[k], NULL, msm_worker, &jobs[k]);
    }
    jac_set_identity(&C->fb, &acc);
    for (size_t k = 0; k < njobs; k++) {
      pthread_join(th[k], NULL);
      jac_add(&C->fb, &acc, &acc, &jobs[k].out);
    }
    free(jobs);
    free(th);
  } else {
    msm_job j = {C, (const fe*)scalars_mont, (const aff*)bases_aff, n, {{{0}}, {{0}}, {{0}}}};
    msm_worker(&j);
    acc = j.out;
  }
  memcpy(out_jac, &acc, 96);
  return 0;
}

int oracle_to_affine(int curve_id, const uint64_t* jac_in, uint64_t* aff_out) {
  if (curve_id < 0 || curve_id > 3) return -2;
  oracle_init();
  jac p;
  memcpy(&p, jac_in, 96);
  aff a;
  jac_to_affine(&CURVES[curve_id].fb, &a, &p);
  memcpy(aff_out, &a, 64);
  return 0;
}

/* out = a + r*b on the curve (Jacobian in/out, r a Montgomery scalar): RelaxedR1CSInstance::fold */
int oracle_point_scale_add(int curve_id, const uint64_t* a_jac, const uint64_t* r_mont, const uint64_t* b_jac, uint64_t* out_jac) {
  if (curve_id < 0 || curve_id > 3) return -2;
  oracle_init();
  const curve_t* C = &CURVES[curve_id];
  jac a, b, acc;
  fe r;
  memcpy(&a, a_jac, 96);
  memcpy(&b, b_jac, 96);
  f_from_mont(&C->fs, &r, (const fe*)r_mont);
  jac_set_identity(&C->fb, &acc);
  for (int i = 255; i >= 0; i--) {
    jac_double(&C->fb, &acc, &acc);
    if ((r.l[i >> 6] >> (i & 63)) & 1) jac_add(&C->fb, &acc, &acc, &b);
  }
  jac_add(&C->fb, &acc, &acc, &a);
  memcpy(out_jac, &acc, 96);
  return 0;
}

/* ---- R1CS  [A.3 - A.5] ------------------------------------------------------------------------ */
typedef struct {
  const field_t* F;
  const uint32_t *row, *col;
  const fe* val;
  size_t nnz, m;
  const fe* z;
  fe* out;
} spmv_job;

static void* spmv_worker(void* arg) {
  spmv_job* j = (spmv_job*)arg;
  memset(j->out, 0, j->m * sizeof(fe));
  for (size_t k = 0; k < j->nnz; k++) {
    fe t;
    f_mul(j->F, &t, &j->val[k], &j->z[j->col[k]]);
    f_add(j->F, &j->out[j->row[k]], &j->out[j->row[k]], &t);
  }
  return NULL;
}

/* (Az, Bz, Cz); the three matrices run concurrently like the nested rayon::join. Returns -3 on
 * InvalidWitnessLength. */
int oracle_multiply_vec(int curve_id, size_t m, size_t num_vars, size_t num_io,
                        const uint32_t* rowA, const uint32_t* colA, const uint64_t* valA, size_t nnzA,
                        const uint32_t* rowB, const uint32_t* colB, const uint64_t* valB, size_t nnzB,
                        const uint32_t* rowC, const uint32_t* colC, const uint64_t* valC, size_t nnzC,
                        const uint64_t* z, size_t z_len, uint64_t* Az, uint64_t* Bz, uint64_t* Cz, int parallel) {
  if (curve_id < 0 || curve_id > 3) return -2;
  if (z_len != num_io + num_vars + 1) return -3;
  oracle_init();
  const field_t* F = &CURVES[curve_id].fs;
  spmv_job jobs[3] = {
      {F, rowA, colA, (const fe*)valA, nnzA, m, (const fe*)z, (fe*)Az},
      {F, rowB, colB, (const fe*)valB, nnzB, m, (const fe*)z, (fe*)Bz},
      {F, rowC, colC, (const fe*)valC, nnzC, m, (const fe*)z, (fe*)Cz},
  };
  if (parallel) {
    pthread_t th[3];
    for (int k = 0; k < 3; k++) pthread_create(&th[k], NULL, spmv_worker, &jobs[k]);
    for (int k = 0; k < 3; k++) pthread_join(th[k], NULL);
  } else {
    for (int k = 0; k < 3; k++) spmv_worker(&jobs[k]);
  }
  return 0;
}

typedef struct {
  const field_t* F;
  const fe *a1, *b1, *c1, *a2, *b2, *c2, *u1;
  fe* T;
  size_t lo, hi;
} ct_job;
static void* ct_worker(void* arg) {
  ct_job* j = (ct_job*)arg;
  for (size_t i = j->lo; i < j->hi; i++) {
    fe t, s;
    f_mul(j->F, &t, &j->a1[i], &j->b2[i]);
    f_mul(j->F, &s, &j->a2[i], &j->b1[i]);
    f_add(j->F, &t, &t, &s);
    f_mul(j->F, &s, j->u1, &j->c2[i]);
    f_sub(j->F, &t, &t, &s);
    f_sub(j->F, &j->T[i], &t, &j->c1[i]);
  }
  return NULL;
}
/* T = Az1 o Bz2 + Az2 o Bz1 - u1*Cz2 - Cz1 */
int oracle_cross_term(int curve_id, size_t m, const uint64_t* Az1, const uint64_t* Bz1, const uint64_t* Cz1,
                      const uint64_t* Az2, const uint64_t* Bz2, const uint64_t* Cz2, const uint64_t* u1, uint64_t* T, int nthreads) {
  if (curve_id < 0 || curve_id > 3) return -2;
  oracle_init();
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  ct_job jobs[256];
  pthread_t th[256];
  size_t per = (m + nthreads - 1) / nthreads;
  for (int k = 0; k < nthreads; k++) {
    size_t lo = (size_t)k * per, hi = lo + per > m ? m : lo + per;
    if (lo > m) lo = m;
    ct_job j = {&CURVES[curve_id].fs, (const fe*)Az1, (const fe*)Bz1, (const fe*)Cz1, (const fe*)Az2, (const fe*)Bz2,
                (const fe*)Cz2, (const fe*)u1, (fe*)T, lo, hi};
    jobs[k] = j;
    pthread_create(&th[k], NULL, ct_worker, &jobs[k]);
  }
  for (int k = 0; k < nthreads; k++) pthread_join(th[k], NULL);
  return 0;
}

typedef struct {
  const field_t* F;
  const fe *a, *b, *r;
  fe* out;
  size_t lo, hi;
} axpy_job;
static void* axpy_worker(void* arg) {
  axpy_job* j = (axpy_job*)arg;
  for (size_t i = j->lo; i < j->hi; i++) {
    fe t;
    f_mul(j->F, &t, j->r, &j->b[i]);
    f_add(j->F, &j->out[i], &j->a[i], &t);
  }
  return NULL;
}
/* out = a + r*b  (W = W1 + r*W2, E = E1 + r*T) */
int oracle_axpy(int curve_id, size_t len, const uint64_t* a, const uint64_t* b, const uint64_t* r, uint64_t* out, int nthreads) {
  if (curve_id < 0 || curve_id > 3) return -2;
  oracle_init();
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  axpy_job jobs[256];
  pthread_t th[256];
  size_t per = (len + nthreads - 1) / nthreads;
  for (int k = 0; k < nthreads; k++) {
    size_t lo = (size_t)k * per, hi = lo + per > len ? len : lo + per;
    if (lo > len) lo = len;
    axpy_job j = {&CURVES[curve_id].fs, (const fe*)a, (const fe*)b, (const fe*)r, (fe*)out, lo, hi};
    jobs[k] = j;
    pthread_create(&th[k], NULL, axpy_worker, &jobs[k]);
  }
  for (int k = 0; k < nthreads; k++) pthread_join(th[k], NULL);
  return 0;
}

/* element-wise field op: which 0 = coordinate field, 1 = scalar field; op 0 mul, 1 add, 2 sub */
int oracle_field_op(int curve_id, int which, int op, const uint64_t* a, const uint64_t* b, size_t n, uint64_t* out) {
  if (curve_id < 0 || curve_id > 3) return -2;
  oracle_init();
  const field_t* F = which == 0 ? &CURVES[curve_id].fb : &CURVES[curve_id].fs;
  for (size_t i = 0; i < n; i++) {
    const fe* x = (const fe*)a + i;
    const fe* y = (const fe*)b + i;
    fe* o = (fe*)out + i;
    if (op == 0) f_mul(F, o, x, y);
    else if (op == 1) f_add(F, o, x, y);
    else f_sub(F, o, x, y);
  }
  return 0;
}

/* bases[i] = (k0 + i*dk) * G as affine Montgomery points (same family as vimz_gen_bases_dev), with one
 * batched inversion.  gx/gy: generator in Montgomery form. */
int oracle_gen_bases(int curve_id, const uint64_t* g_aff, uint64_t k0, uint64_t dk, size_t n, uint64_t* out_aff) {
  if (curve_id < 0 || curve_id > 3) return -2;
  oracle_init();
  const field_t* F = &CURVES[curve_id].fb;
  aff g;
  memcpy(&g, g_aff, 64);
  jac G, step, cur;
  G.x = g.x; G.y = g.y; G.z = F->one;
  jac_set_identity(F, &step);
  jac_set_identity(F, &cur);
  for (int i = 63; i >= 0; i--) {
    jac_double(F, &step, &step);
    if ((dk >> i) & 1) jac_add(F, &step, &step, &G);
    jac_double(F, &cur, &cur);
    if ((k0 >> i) & 1) jac_add(F, &cur, &cur, &G);
  }
  jac* pts = (jac*)malloc((n ? n : 1) * sizeof(jac));
  fe* prefix = (fe*)malloc((n ? n : 1) * sizeof(fe));
  fe run = F->one;
  for (size_t i = 0; i < n; i++) {
    pts[i] = cur;
    prefix[i] = run;
    if (!fe_is_zero(&cur.z)) f_mul(F, &run, &run, &cur.z);
    jac_add(F, &cur, &cur, &step);
  }
  fe inv;
  f_inv(F, &inv, &run);
  aff* out = (aff*)out_aff;
  for (size_t i = n; i-- > 0;) {
    if (fe_is_zero(&pts[i].z)) { memset(&out[i], 0, sizeof(aff)); continue; }
    fe zi, zi2, zi3;
    f_mul(F, &zi, &inv, &prefix[i]);
    f_mul(F, &inv, &inv, &pts[i].z);
    f_sqr(F, &zi2, &zi);
    f_mul(F, &zi3, &zi2, &zi);
    f_mul(F, &out[i].x, &pts[i].x, &zi2);
    f_mul(F, &out[i].y, &pts[i].y, &zi3);
  }
  free(pts);
  free(prefix);
  return 0;
}

/* ---- circomlib Poseidon permutation over a curve's scalar field (oracle/poseidon.py is the definition-level
 * restatement and supplies the Grain-LFSR constants; this is its fast path for the 720-row image hashes of
 * /root/reference/marketplace/image-data/ (the .hash files).  x^5 S-box, rf full + rp partial rounds, plain form:
 * add round constants, S-box (all cells in full rounds, cell 0 in partial rounds), multiply by the MDS matrix.
 * consts: (rf + rp) * t, mds: t * t row-major, state: count * t in/out; all Montgomery form. */
int oracle_poseidon(int curve_id, int t, int rf, int rp, const uint64_t* consts, const uint64_t* mds, uint64_t* state, size_t count) {
  if (curve_id < 0 || curve_id > 3 || t < 2 || t > 17) return -2;
  oracle_init();
  const field_t* F = &CURVES[curve_id].fs;
  const fe* C = (const fe*)consts;
  const fe* M = (const fe*)mds;
  for (size_t n = 0; n < count; n++) {
    fe* s = (fe*)state + n * (size_t)t;
    fe tmp[17];
    for (int r = 0; r < rf + rp; r++) {
      for (int i = 0; i < t; i++) f_add(F, &s[i], &s[i], &C[r * t + i]);
      int full = (r < rf / 2) || (r >= rf / 2 + rp);
      for (int i = 0; i < (full ? t : 1); i++) {
        fe x2, x4;
        f_sqr(F, &x2, &s[i]);
        f_sqr(F, &x4, &x2);
        f_mul(F, &s[i], &x4, &s[i]);
      }
      for (int i = 0; i < t; i++) {
        fe acc = {{0, 0, 0, 0}};
        for (int j = 0; j < t; j++) {
          fe m;
          f_mul(F, &m, &M[i * t + j], &s[j]);
          f_add(F, &acc, &acc, &m);
        }
        tmp[i] = acc;
      }
      memcpy(s, tmp, (size_t)t * sizeof(fe));
    }
  }
  return 0;
}
