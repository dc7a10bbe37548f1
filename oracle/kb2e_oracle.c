/* TEST INFRASTRUCTURE ONLY -- see kb2e_oracle.h for scope, pinning and layout.
 *
 * Every "ref" function keeps the reference's exact operation order (fp64, left-to-right sums,
 * separate multiply and add -- build with -ffp-contract=off) so that it can be compared BITWISE
 * with the reference compiled here (oracle/_ref).  Citations are into /root/reference.
 */
#include "kb2e_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

static double sqr(double x) { return x * x; } /* common/utils.cpp:40-42 */

/* ============================== energies ====================================================== */

/* transe/transe.cpp:10-28: sum_i |t_i - h_i - r_i| or sum_i (t_i - h_i - r_i)^2, evaluated (t-h)-r. */
static double energy_transe(int distance, int D, const double* ent, const double* rel, int h, int t, int r) {
   const double* eh = ent + (size_t)h * D;
   const double* et = ent + (size_t)t * D;
   const double* er = rel + (size_t)r * D;
   double energy = 0;
   if (distance == 0) {
      for (int i = 0; i < D; i++) energy += fabs(et[i] - eh[i] - er[i]);
   } else {
      for (int i = 0; i < D; i++) energy += sqr(et[i] - eh[i] - er[i]);
   }
   return energy;
}

/* transh/transh.cpp:10-29 (L1 only; ignores -distance). */
static double energy_transh(int D, const double* ent, const double* rel, const double* w, int h, int t, int r) {
   const double* eh = ent + (size_t)h * D;
   const double* et = ent + (size_t)t * D;
   const double* er = rel + (size_t)r * D;
   const double* wr = w + (size_t)r * D;
   double headSum = 0, tailSum = 0;
   for (int i = 0; i < D; i++) {
      headSum += wr[i] * eh[i];
      tailSum += wr[i] * et[i];
   }
   double energy = 0;
   for (int i = 0; i < D; i++) {
      energy += fabs(et[i] - tailSum * wr[i] - (eh[i] - headSum * wr[i]) - er[i]);
   }
   return energy;
}

/* transr/transr.cpp:13-37 as shipped: headVec/tailVec are caller-owned accumulators. */
double orc_energy_transr_shipped(int distance, int D, const double* ent, const double* rel, const double* w,
                                 int h, int t, int r, double* headVec, double* tailVec) {
   const double* eh = ent + (size_t)h * D;
   const double* et = ent + (size_t)t * D;
   const double* er = rel + (size_t)r * D;
   const double* M = w + (size_t)r * D * D;
   for (int i = 0; i < D; i++) {
      for (int j = 0; j < D; j++) {
         headVec[i] += M[(size_t)j * D + i] * eh[j];
         tailVec[i] += M[(size_t)j * D + i] * et[j];
      }
   }
   double sum = 0;
   for (int i = 0; i < D; i++) {
      if (distance == 0) {
         sum += fabs(tailVec[i] - headVec[i] - er[i]);
      } else {
         sum += sqr(tailVec[i] - headVec[i] - er[i]);
      }
   }
   return sum;
}

static double energy_transr(int distance, int D, const double* ent, const double* rel, const double* w, int h, int t, int r) {
   double* work = (double*)calloc((size_t)2 * D, sizeof(double));
   double e = orc_energy_transr_shipped(distance, D, ent, rel, w, h, t, r, work, work + D);
   free(work);
   return e;
}

double orc_energy(int model, int distance, int D, const double* ent, const double* rel, const double* w,
                  int h, int t, int r) {
   if (model == 0) return energy_transe(distance, D, ent, rel, h, t, r);
   if (model == 1) return energy_transh(D, ent, rel, w, h, t, r);
   return energy_transr(distance, D, ent, rel, w, h, t, r);
}

void orc_energy_many(int model, int distance, int D, const double* ent, const double* rel, const double* w,
                     long n, const int* h, const int* t, const int* r, double* out) {
   for (long k = 0; k < n; k++) out[k] = orc_energy(model, distance, D, ent, rel, w, h[k], t[k], r[k]);
}

/* ============================== normalisations ================================================ */

double orc_vec_len(const double* a, int n) { /* common/utils.cpp:44-51 */
   double res = 0;
   for (int i = 0; i < n; i++) res += sqr(a[i]);
   return sqrt(res);
}

void orc_norm(double* a, int n, int ignoreShort) { /* common/utils.cpp:70-77 */
   double len = orc_vec_len(a, n);
   if (!ignoreShort || len > 1) {
      for (int i = 0; i < n; i++) a[i] /= len;
   }
}

/* common/utils.cpp:79-111.  `sum` is NOT reset between iterations (reference behaviour, kept). */
void orc_norm2(double* a, double* b, int n, double rate) {
   orc_norm(b, n, 0);
   double sum = 0;
   while (1) {
      for (int i = 0; i < n; i++) sum += sqr(b[i]);
      sum = sqrt(sum);
      for (int i = 0; i < n; i++) b[i] /= sum;
      double x = 0;
      for (int i = 0; i < n; i++) x += b[i] * a[i];
      if (x > 0.1) {
         for (int i = 0; i < n; i++) {
            a[i] -= rate * b[i];
            b[i] -= rate * a[i];
         }
      } else {
         break;
      }
   }
   orc_norm(b, n, 0);
}

/* transr/trainer.cpp:35-64; M = b[j][i] row-major [D][D], lambda = 1. */
void orc_transr_norm(double* a, double* M, int D, double lr) {
   while (1) {
      double x = 0;
      for (int i = 0; i < D; i++) {
         double tmp = 0;
         for (int j = 0; j < D; j++) tmp += M[(size_t)j * D + i] * a[j];
         x += sqr(tmp);
      }
      if (x <= 1) break;
      double lambda = 1;
      for (int i = 0; i < D; i++) {
         double tmp = 0;
         for (int j = 0; j < D; j++) tmp += M[(size_t)j * D + i] * a[j];
         tmp *= 2;
         for (int j = 0; j < D; j++) {
            M[(size_t)j * D + i] -= lr * lambda * tmp * a[j];
            a[j] -= lr * lambda * tmp * M[(size_t)j * D + i];
         }
      }
   }
}

/* ============================== gradient updates ============================================== */

/* `raw` != 0 stops before the normalisation tail of each gradientUpdate: the accumulation code is shared with the
 * bitwise-pinned normalising form, so (raw result, then the same tail) == the reference by construction. */
static void grad_transe(int distance, int D, double lr, const double* ent, const double* rel,
                        double* entN, double* relN, int h, int t, int r, int corrupted, int raw) {
   /* transe/trainer.cpp:25-46 */
   double modifier = corrupted ? 1.0 : -1.0;
   const double* eh = ent + (size_t)h * D;
   const double* et = ent + (size_t)t * D;
   const double* er = rel + (size_t)r * D;
   double* nh = entN + (size_t)h * D;
   double* nt = entN + (size_t)t * D;
   double* nr = relN + (size_t)r * D;
   for (int i = 0; i < D; i++) {
      double x = 2.0 * (et[i] - eh[i] - er[i]);
      if (distance == 0) x = (x > 0) ? 1 : -1;
      nr[i] -= modifier * lr * x;
      nh[i] -= modifier * lr * x;
      nt[i] += modifier * lr * x;
   }
   if (raw) return;
   orc_norm(nr, D, 1);
   orc_norm(nh, D, 1);
   orc_norm(nt, D, 1);
}

static void grad_transh(int D, double lr, const double* ent, const double* rel, const double* w,
                        double* entN, double* relN, double* wN, int h, int t, int r, int corrupted, int raw) {
   /* transh/trainer.cpp:11-59 */
   double beta = corrupted ? 1 : -1;
   const double* eh = ent + (size_t)h * D;
   const double* et = ent + (size_t)t * D;
   const double* er = rel + (size_t)r * D;
   const double* wr = w + (size_t)r * D;
   double* nh = entN + (size_t)h * D;
   double* nt = entN + (size_t)t * D;
   double* nr = relN + (size_t)r * D;
   double* nw = wN + (size_t)r * D;
   double headSum = 0, tailSum = 0, sum_x = 0;
   for (int i = 0; i < D; i++) {
      headSum += wr[i] * eh[i];
      tailSum += wr[i] * et[i];
   }
   for (int i = 0; i < D; i++) {
      double x = 2 * (et[i] - tailSum * wr[i] - (eh[i] - headSum * wr[i]) - er[i]);
      x = (x > 0) ? 1 : -1;
      sum_x += x * wr[i];
      nr[i] -= beta * lr * x;
      nh[i] -= beta * lr * x;
      nt[i] += beta * lr * x;
      nw[i] += beta * lr * x * headSum;
      nw[i] -= beta * lr * x * tailSum;
   }
   for (int i = 0; i < D; i++) {
      nw[i] += beta * lr * sum_x * eh[i];
      nw[i] -= beta * lr * sum_x * et[i];
   }
   if (raw) return;
   orc_norm(nr, D, 1);
   orc_norm(nh, D, 1);
   orc_norm(nt, D, 1);
   orc_norm(nw, D, 0);
   orc_norm2(nr, nw, D, lr);
   orc_norm2(nh, nw, D, lr);
   orc_norm2(nt, nw, D, lr);
}

static void grad_transr(int distance, int D, double lr, const double* ent, const double* rel, const double* w,
                        double* entN, double* relN, double* wN, int h, int t, int r, int corrupted, int raw) {
   /* transr/trainer.cpp:144-188 */
   double beta = corrupted ? 1.0 : -1.0;
   const double* eh = ent + (size_t)h * D;
   const double* et = ent + (size_t)t * D;
   const double* er = rel + (size_t)r * D;
   const double* M = w + (size_t)r * D * D;
   double* nh = entN + (size_t)h * D;
   double* nt = entN + (size_t)t * D;
   double* nr = relN + (size_t)r * D;
   double* nM = wN + (size_t)r * D * D;
   for (int i = 0; i < D; i++) {
      double headSum = 0, tailSum = 0;
      for (int j = 0; j < D; j++) {
         headSum += M[(size_t)j * D + i] * eh[j];
         tailSum += M[(size_t)j * D + i] * et[j];
      }
      double x = 2.0 * (tailSum - headSum - er[i]);
      if (distance == 0) x = (x > 0) ? 1 : -1;
      for (int j = 0; j < D; j++) {
         nM[(size_t)j * D + i] -= beta * lr * x * (eh[j] - et[j]);
         nh[j] -= beta * lr * x * M[(size_t)j * D + i];
         nt[j] += beta * lr * x * M[(size_t)j * D + i];
      }
      nr[i] -= beta * lr * x;
   }
   if (raw) return;
   orc_norm(nr, D, 0);
   orc_norm(nh, D, 0);
   orc_norm(nt, D, 0);
   for (int j = 0; j < D; j++) orc_norm(nM + (size_t)j * D, D, 0);
   orc_transr_norm(nh, nM, D, lr);
   orc_transr_norm(nt, nM, D, lr);
   /* transr/trainer.cpp:187 indexes the ENTITY table with the relation id (reference quirk, kept). */
   orc_transr_norm(entN + (size_t)r * D, nM, D, lr);
}

void orc_grad(int model, int distance, int D, double lr,
              const double* ent, const double* rel, const double* w,
              double* entN, double* relN, double* wN, int h, int t, int r, int corrupted) {
   if (model == 0) grad_transe(distance, D, lr, ent, rel, entN, relN, h, t, r, corrupted, 0);
   else if (model == 1) grad_transh(D, lr, ent, rel, w, entN, relN, wN, h, t, r, corrupted, 0);
   else grad_transr(distance, D, lr, ent, rel, w, entN, relN, wN, h, t, r, corrupted, 0);
}

/* The same accumulation WITHOUT the normalisation tail (transe/trainer.cpp:43-45, transh/trainer.cpp:48-58,
 * transr/trainer.cpp:174-187): entN/relN/wN - their inputs = the pre-normalisation update of one gradientUpdate. */
void orc_grad_raw(int model, int distance, int D, double lr,
                  const double* ent, const double* rel, const double* w,
                  double* entN, double* relN, double* wN, int h, int t, int r, int corrupted) {
   if (model == 0) grad_transe(distance, D, lr, ent, rel, entN, relN, h, t, r, corrupted, 1);
   else if (model == 1) grad_transh(D, lr, ent, rel, w, entN, relN, wN, h, t, r, corrupted, 1);
   else grad_transr(distance, D, lr, ent, rel, w, entN, relN, wN, h, t, r, corrupted, 1);
}

/* The normalisation tail alone, on rows that already hold the accumulated update. */
void orc_grad_tail(int model, int D, double lr, double* entN, double* relN, double* wN, int h, int t, int r) {
   double* nh = entN + (size_t)h * D;
   double* nt = entN + (size_t)t * D;
   double* nr = relN + (size_t)r * D;
   if (model == 0) {
      orc_norm(nr, D, 1); orc_norm(nh, D, 1); orc_norm(nt, D, 1);
   } else if (model == 1) {
      double* nw = wN + (size_t)r * D;
      orc_norm(nr, D, 1); orc_norm(nh, D, 1); orc_norm(nt, D, 1); orc_norm(nw, D, 0);
      orc_norm2(nr, nw, D, lr); orc_norm2(nh, nw, D, lr); orc_norm2(nt, nw, D, lr);
   } else {
      double* nM = wN + (size_t)r * D * D;
      orc_norm(nr, D, 0); orc_norm(nh, D, 0); orc_norm(nt, D, 0);
      for (int j = 0; j < D; j++) orc_norm(nM + (size_t)j * D, D, 0);
      orc_transr_norm(nh, nM, D, lr);
      orc_transr_norm(nt, nM, D, lr);
      orc_transr_norm(entN + (size_t)r * D, nM, D, lr);
   }
}

static size_t w_elems(int model, int D, int nR) {
   if (model == 1) return (size_t)nR * D;
   if (model == 2) return (size_t)nR * D * D;
   return 0;
}

double orc_train_batch_ref(int model, int distance, int D, int nE, int nR, double lr, double margin,
                           const double* ent, const double* rel, const double* w,
                           long n, const int* pairs,
                           double* entN, double* relN, double* wN, double* losses) {
   /* prebatch: next = cur (transe/trainer.cpp:53-56) */
   memcpy(entN, ent, sizeof(double) * (size_t)nE * D);
   memcpy(relN, rel, sizeof(double) * (size_t)nR * D);
   if (w_elems(model, D, nR)) memcpy(wN, w, sizeof(double) * w_elems(model, D, nR));
   double total = 0;
   for (long k = 0; k < n; k++) {
      const int* p = pairs + 6 * k;
      /* common/trainer.cpp:130-149 */
      double loss = 0;
      double normalEnergy = orc_energy(model, distance, D, ent, rel, w, p[0], p[1], p[2]);
      double corruptedEnergy = orc_energy(model, distance, D, ent, rel, w, p[3], p[4], p[5]);
      if (normalEnergy + margin > corruptedEnergy) {
         loss = margin + normalEnergy - corruptedEnergy;
         orc_grad(model, distance, D, lr, ent, rel, w, entN, relN, wN, p[0], p[1], p[2], 0);
         orc_grad(model, distance, D, lr, ent, rel, w, entN, relN, wN, p[3], p[4], p[5], 1);
      }
      if (losses) losses[k] = loss;
      total += loss;
   }
   return total;
}

/* ============================== bern statistics =============================================== */

typedef struct { int a, b; } pair2;
static int cmp_pair2(const void* x, const void* y) {
   const pair2* p = (const pair2*)x;
   const pair2* q = (const pair2*)y;
   if (p->a != q->a) return p->a < q->a ? -1 : 1;
   if (p->b != q->b) return p->b < q->b ? -1 : 1;
   return 0;
}

/* common/trainer.cpp:163-194: per relation, (#triples) / (#distinct heads) and / (#distinct tails);
 * 0 when the relation has no triples. */
void orc_bern(long n, const int* h, const int* t, const int* r, int nR, double* head_mean, double* tail_mean) {
   pair2* v = (pair2*)malloc(sizeof(pair2) * (size_t)(n > 0 ? n : 1));
   for (int side = 0; side < 2; side++) {
      double* out = side == 0 ? head_mean : tail_mean;
      for (long k = 0; k < n; k++) {
         v[k].a = r[k];
         v[k].b = side == 0 ? h[k] : t[k];
      }
      qsort(v, (size_t)n, sizeof(pair2), cmp_pair2);
      for (int i = 0; i < nR; i++) out[i] = 0;
      long k = 0;
      while (k < n) {
         int rr = v[k].a;
         long total = 0, distinct = 0;
         while (k < n && v[k].a == rr) {
            int e = v[k].b;
            distinct++;
            while (k < n && v[k].a == rr && v[k].b == e) { total++; k++; }
         }
         if (rr >= 0 && rr < nR) out[rr] = (double)total / (double)distinct;
      }
   }
   free(v);
}

/* ============================== triple set ==================================================== */

typedef struct { int h, r, t; } trip;
static int cmp_trip(const void* x, const void* y) {
   const trip* p = (const trip*)x;
   const trip* q = (const trip*)y;
   if (p->h != q->h) return p->h < q->h ? -1 : 1;
   if (p->r != q->r) return p->r < q->r ? -1 : 1;
   if (p->t != q->t) return p->t < q->t ? -1 : 1;
   return 0;
}
static int trip_in(const trip* set, long n, int h, int r, int t) {
   trip key = {h, r, t};
   return bsearch(&key, set, (size_t)n, sizeof(trip), cmp_trip) != NULL;
}

/* ============================== ranking ======================================================= */

void orc_rank(int model, int distance, int D, int nE, int nR,
              const double* ent, const double* rel, const double* w,
              long nTest, const int* th, const int* tt, const int* tr,
              long nFilter, const int* fh, const int* ft, const int* fr,
              int* raw_lo, int* raw_hi, int* filt_lo, int* filt_hi) {
   (void)nR;
   long nk = nTest + nFilter;
   trip* known = (trip*)malloc(sizeof(trip) * (size_t)(nk > 0 ? nk : 1));
   for (long i = 0; i < nTest; i++) { known[i].h = th[i]; known[i].r = tr[i]; known[i].t = tt[i]; }
   for (long i = 0; i < nFilter; i++) { known[nTest + i].h = fh[i]; known[nTest + i].r = fr[i]; known[nTest + i].t = ft[i]; }
   qsort(known, (size_t)nk, sizeof(trip), cmp_trip);
   double* en = (double*)malloc(sizeof(double) * (size_t)nE);
   for (long q = 0; q < nTest; q++) {
      for (int side = 0; side < 2; side++) {
         int head = th[q], tail = tt[q], r = tr[q];
         /* common/evaluation.cpp:129-135: score every entity, the true one included. */
         for (int c = 0; c < nE; c++) {
            en[c] = side == 0 ? orc_energy(model, distance, D, ent, rel, w, c, tail, r)
                              : orc_energy(model, distance, D, ent, rel, w, head, c, r);
         }
         int truth = side == 0 ? head : tail;
         double et = en[truth];
         int less = 0, eq = 0, fless = 0, feq = 0;
         for (int c = 0; c < nE; c++) {
            if (c == truth) continue;
            int isLess = en[c] < et, isEq = en[c] == et;
            if (!isLess && !isEq) continue;
            int knownTriple = side == 0 ? trip_in(known, nk, c, r, tail) : trip_in(known, nk, head, r, c);
            less += isLess;
            eq += isEq;
            if (!knownTriple) { fless += isLess; feq += isEq; }
         }
         long o = 2 * q + side;
         raw_lo[o] = 1 + less;
         raw_hi[o] = 1 + less + eq;
         filt_lo[o] = 1 + fless;
         filt_hi[o] = 1 + fless + feq;
      }
   }
   free(en);
   free(known);
}

/* ============================== counter RNG + sampler ========================================= */

void orc_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
   for (int round = 0; round < 10; round++) {
      uint64_t p0 = (uint64_t)0xD2511F53u * c0;
      uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
      uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
      uint32_t n1 = (uint32_t)p1;
      uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
      uint32_t n3 = (uint32_t)p0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
   }
   out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct orc_sampler {
   long n;
   int nE, nR;
   int *h, *t, *r;
   trip* set;
   double* pr;
   int mode; /* 0: uniform indices (mulhi); 1: the index DISTRIBUTION of the reference's randMax */
};

orc_sampler* orc_sampler_create(long n, const int* h, const int* t, const int* r, int nE, int nR, int method) {
   orc_sampler* s = (orc_sampler*)calloc(1, sizeof(orc_sampler));
   s->n = n; s->nE = nE; s->nR = nR;
   s->h = (int*)malloc(sizeof(int) * (size_t)n);
   s->t = (int*)malloc(sizeof(int) * (size_t)n);
   s->r = (int*)malloc(sizeof(int) * (size_t)n);
   memcpy(s->h, h, sizeof(int) * (size_t)n);
   memcpy(s->t, t, sizeof(int) * (size_t)n);
   memcpy(s->r, r, sizeof(int) * (size_t)n);
   s->set = (trip*)malloc(sizeof(trip) * (size_t)n);
   for (long i = 0; i < n; i++) { s->set[i].h = h[i]; s->set[i].r = r[i]; s->set[i].t = t[i]; }
   qsort(s->set, (size_t)n, sizeof(trip), cmp_trip);
   s->pr = (double*)malloc(sizeof(double) * (size_t)nR);
   double* hm = (double*)malloc(sizeof(double) * (size_t)nR);
   double* tm = (double*)malloc(sizeof(double) * (size_t)nR);
   orc_bern(n, h, t, r, nR, hm, tm);
   for (int i = 0; i < nR; i++) {
      /* common/trainer.cpp:82-86 */
      s->pr[i] = (method == 0) ? 500 : 1000 * tm[i] / (tm[i] + hm[i]);
   }
   free(hm); free(tm);
   return s;
}

void orc_sampler_destroy(orc_sampler* s) {
   if (!s) return;
   free(s->h); free(s->t); free(s->r); free(s->set); free(s->pr); free(s);
}

const double* orc_sampler_pr(const orc_sampler* s) { return s->pr; }
void orc_sampler_set_mode(orc_sampler* s, int mode) { s->mode = mode; }

/* common/utils.cpp:113-120 with two uniform 31-bit draws in place of the two std::rand() calls:
 * `(rand() * rand()) % x` in 32-bit int arithmetic (the product wraps), then `while (res < 0) res += x`.
 * Same index distribution as the reference (e.g. 75 % even values), from the counter RNG. */
static int randmax_from(uint32_t a, uint32_t b, int x) {
   int32_t res = (int32_t)((a >> 1) * (b >> 1));
   res = res % x; /* C remainder: sign of the dividend */
   if (res < 0) res += x;
   return (int)res;
}

static uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static uint64_t mulhi64(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) >> 64); }

/* n draws of the randMax emulation from the counter RNG (test hook: compared with the reference's own randMax). */
void orc_randmax_draws(uint64_t seed, int x, long n, int* out) {
   uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
   for (long i = 0; i < n; i++) {
      uint32_t y[4];
      orc_philox((uint32_t)i, (uint32_t)(i >> 32), 0, 7, k0, k1, y);
      out[i] = randmax_from(y[0], y[1], x);
   }
}

void orc_sample_batch(const orc_sampler* s, uint64_t seed, uint32_t gb, long count, int* out) {
   uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
   for (long k = 0; k < count; k++) {
      uint32_t x[4];
      orc_philox((uint32_t)k, gb, 0, 0, k0, k1, x);
      long i = (long)mulhi64(((uint64_t)x[0] << 32) | x[1], (uint64_t)s->n);
      int coin = (int)(x[2] % 1000u);
      int j = (int)mulhi32(x[3], (uint32_t)s->nE);
      if (s->mode == 1) {
         uint32_t y[4];
         orc_philox((uint32_t)k, gb, 0, 1, k0, k1, y);
         i = randmax_from(x[0], x[1], (int)s->n);
         j = randmax_from(y[0], y[1], s->nE);
      }
      int h = s->h[i], t = s->t[i], r = s->r[i];
      int* p = out + 6 * k;
      p[0] = h; p[1] = t; p[2] = r; p[5] = r;
      /* common/trainer.cpp:88-98: `rand() % 1000 < pr` -> corrupt the tail, else the head. */
      int corruptTail = (double)coin < s->pr[r];
      for (uint32_t a = 1; a < 64; a++) {
         int hit = corruptTail ? trip_in(s->set, s->n, h, r, j) : trip_in(s->set, s->n, j, r, t);
         if (!hit) break;
         orc_philox((uint32_t)k, gb, a, 0, k0, k1, x);
         j = s->mode == 1 ? randmax_from(x[0], x[1], s->nE) : (int)mulhi32(x[0], (uint32_t)s->nE);
      }
      if (corruptTail) { p[3] = h; p[4] = j; } else { p[3] = j; p[4] = t; }
   }
}

/* ============================== deferred-renorm batch (what the CUDA path computes) ============ */

/* x_i for one triple: 2*residual or its L1 sign (0 -> -1), plus the TransH/TransR side products. */
static void residual_sign(int model, int distance, int D, const double* ent, const double* rel, const double* w,
                          int h, int t, int r, double* x, double* hs_out, double* ts_out) {
   const double* eh = ent + (size_t)h * D;
   const double* et = ent + (size_t)t * D;
   const double* er = rel + (size_t)r * D;
   if (model == 0) {
      for (int i = 0; i < D; i++) x[i] = 2.0 * (et[i] - eh[i] - er[i]);
   } else if (model == 1) {
      const double* wr = w + (size_t)r * D;
      double hs = 0, ts = 0;
      for (int i = 0; i < D; i++) { hs += wr[i] * eh[i]; ts += wr[i] * et[i]; }
      for (int i = 0; i < D; i++) x[i] = 2 * (et[i] - ts * wr[i] - (eh[i] - hs * wr[i]) - er[i]);
      *hs_out = hs; *ts_out = ts;
   } else {
      const double* M = w + (size_t)r * D * D;
      for (int i = 0; i < D; i++) {
         double hs = 0, ts = 0;
         for (int j = 0; j < D; j++) { hs += M[(size_t)j * D + i] * eh[j]; ts += M[(size_t)j * D + i] * et[j]; }
         x[i] = 2.0 * (ts - hs - er[i]);
      }
   }
   if (distance == 0 || model == 1) {
      for (int i = 0; i < D; i++) x[i] = (x[i] > 0) ? 1 : -1;
   }
}

/* Accumulate the update of one triple into the delta tables (no normalisation here). */
static void accumulate(int model, int distance, int D, double lr, const double* ent, const double* rel, const double* w,
                       double* dE, double* dR, double* dW, int h, int t, int r, int corrupted, double* x) {
   double beta = corrupted ? 1.0 : -1.0;
   double hs = 0, ts = 0;
   residual_sign(model, distance, D, ent, rel, w, h, t, r, x, &hs, &ts);
   const double* eh = ent + (size_t)h * D;
   const double* et = ent + (size_t)t * D;
   double* dh = dE + (size_t)h * D;
   double* dt = dE + (size_t)t * D;
   double* dr = dR + (size_t)r * D;
   if (model == 0) {
      for (int i = 0; i < D; i++) {
         dr[i] -= beta * lr * x[i];
         dh[i] -= beta * lr * x[i];
         dt[i] += beta * lr * x[i];
      }
   } else if (model == 1) {
      const double* wr = w + (size_t)r * D;
      double* dw = dW + (size_t)r * D;
      double sx = 0;
      for (int i = 0; i < D; i++) sx += x[i] * wr[i];
      for (int i = 0; i < D; i++) {
         dr[i] -= beta * lr * x[i];
         dh[i] -= beta * lr * x[i];
         dt[i] += beta * lr * x[i];
         dw[i] += beta * lr * (x[i] * (hs - ts) + sx * (eh[i] - et[i]));
      }
   } else {
      const double* M = w + (size_t)r * D * D;
      double* dM = dW + (size_t)r * D * D;
      for (int i = 0; i < D; i++) {
         for (int j = 0; j < D; j++) {
            dM[(size_t)j * D + i] -= beta * lr * x[i] * (eh[j] - et[j]);
            dh[j] -= beta * lr * x[i] * M[(size_t)j * D + i];
            dt[j] += beta * lr * x[i] * M[(size_t)j * D + i];
         }
         dr[i] -= beta * lr * x[i];
      }
   }
}

/* The soft-orthogonality step of common/utils.cpp:79-111 applied to an entity row against a
 * FIXED, already unit-length hyperplane normal: `a` is updated in place, the perturbation the
 * reference would have applied to b is returned in db (b_final - b_initial) so the caller can
 * fold it into the NEXT batch's delta.  Returns the number of corrective iterations. */
static int soft_orth_entity(double* a, const double* b0, double* db, int n, double rate, double* b) {
   memcpy(b, b0, sizeof(double) * (size_t)n);
   double sum = 0;
   int iters = 0;
   while (1) {
      for (int i = 0; i < n; i++) sum += sqr(b[i]);
      sum = sqrt(sum);
      for (int i = 0; i < n; i++) b[i] /= sum;
      double x = 0;
      for (int i = 0; i < n; i++) x += b[i] * a[i];
      if (x > 0.1) {
         for (int i = 0; i < n; i++) {
            a[i] -= rate * b[i];
            b[i] -= rate * a[i];
         }
         iters++;
      } else {
         break;
      }
   }
   if (iters) {
      orc_norm(b, n, 0);
      for (int i = 0; i < n; i++) db[i] += b[i] - b0[i];
   }
   return iters;
}

/* transRNorm (transr/trainer.cpp:35-64) on an entity row against a FIXED matrix: the published M_r is
 * read-only during the phase, so every sweep reads M0; `a` is updated in place exactly as the
 * reference does within a sweep (column i of M is perturbed first, then a uses the perturbed
 * column), and the perturbations of all sweeps are accumulated into dM for the next batch. */
static int transr_norm_entity(double* a, const double* M0, double* dM, int D, double lr, double* unused) {
   (void)unused;
   int iters = 0;
   while (1) {
      double x = 0;
      for (int i = 0; i < D; i++) {
         double tmp = 0;
         for (int j = 0; j < D; j++) tmp += M0[(size_t)j * D + i] * a[j];
         x += sqr(tmp);
      }
      if (x <= 1 || iters >= 64) break;
      for (int i = 0; i < D; i++) {
         double tmp = 0;
         for (int j = 0; j < D; j++) tmp += M0[(size_t)j * D + i] * a[j];
         tmp *= 2;
         for (int j = 0; j < D; j++) {
            double delta = -(lr * tmp * a[j]);
            dM[(size_t)j * D + i] += delta;
            a[j] -= lr * tmp * (M0[(size_t)j * D + i] + delta);
         }
      }
      iters++;
   }
   return iters;
}

double orc_train_batch_dfr(int model, int distance, int D, int nE, int nR, double lr, double margin,
                           double* ent, double* rel, double* w, double* carry,
                           long n, const int* pairs, long* n_active) {
   size_t we = w_elems(model, D, nR);
   size_t wrow = (model == 2) ? (size_t)D * D : (size_t)D;
   double* dE = (double*)calloc((size_t)nE * D, sizeof(double));
   double* dR = (double*)calloc((size_t)nR * D, sizeof(double));
   double* dW = (double*)calloc(we ? we : 1, sizeof(double));
   unsigned char* tE = (unsigned char*)calloc((size_t)nE, 1);
   unsigned char* tR = (unsigned char*)calloc((size_t)nR, 1);
   /* carry = perturbations of w_r / M_r produced by the PREVIOUS batch's entity-side constraint
    * steps; they enter this batch's delta (and mark the relation touched). */
   if (we && carry) {
      for (int r = 0; r < nR; r++) {
         for (size_t q = 0; q < wrow; q++) {
            double c = carry[(size_t)r * wrow + q];
            if (c != 0) { tR[r] = 1; dW[(size_t)r * wrow + q] = c; }
            carry[(size_t)r * wrow + q] = 0;
         }
      }
   }
   int* rmin = (int*)malloc(sizeof(int) * (size_t)nE);
   int* rmax = (int*)malloc(sizeof(int) * (size_t)nE);
   for (int e = 0; e < nE; e++) { rmin[e] = 0x7fffffff; rmax[e] = -1; }
   unsigned char* aR = (unsigned char*)calloc((size_t)nR, 1); /* relation hit by an active sample */
   double* x = (double*)malloc(sizeof(double) * (size_t)D);
   double total = 0;
   long active = 0;
   for (long k = 0; k < n; k++) {
      const int* p = pairs + 6 * k;
      double ep = orc_energy(model, distance, D, ent, rel, w, p[0], p[1], p[2]);
      double en = orc_energy(model, distance, D, ent, rel, w, p[3], p[4], p[5]);
      if (ep + margin > en) {
         total += margin + ep - en;
         active++;
         accumulate(model, distance, D, lr, ent, rel, w, dE, dR, dW, p[0], p[1], p[2], 0, x);
         accumulate(model, distance, D, lr, ent, rel, w, dE, dR, dW, p[3], p[4], p[5], 1, x);
         int ents[4] = {p[0], p[1], p[3], p[4]};
         for (int q = 0; q < 4; q++) {
            tE[ents[q]] = 1;
            if (p[2] < rmin[ents[q]]) rmin[ents[q]] = p[2];
            if (p[2] > rmax[ents[q]]) rmax[ents[q]] = p[2];
         }
         tR[p[2]] = 1;
         aR[p[2]] = 1;
      }
   }
   /* phase 2a: relation-side rows */
   for (int r = 0; r < nR; r++) {
      if (!tR[r]) continue;
      double* rr = rel + (size_t)r * D;
      for (int i = 0; i < D; i++) rr[i] += dR[(size_t)r * D + i];
      if (model == 0) {
         orc_norm(rr, D, 1);
      } else if (model == 1) {
         double* wr = w + (size_t)r * D;
         for (int i = 0; i < D; i++) wr[i] += dW[(size_t)r * D + i];
         orc_norm(rr, D, 1);
         orc_norm(wr, D, 0);
         orc_norm2(rr, wr, D, lr);
      } else {
         double* M = w + (size_t)r * D * D;
         for (size_t q = 0; q < (size_t)D * D; q++) M[q] += dW[(size_t)r * D * D + q];
         orc_norm(rr, D, 0);
         for (int j = 0; j < D; j++) orc_norm(M + (size_t)j * D, D, 0);
      }
   }
   /* phase 2b: entity rows, against the relation-side rows just published (read-only here) */
   double* scratch = (double*)malloc(sizeof(double) * (model == 2 ? (size_t)D * D : (size_t)D));
   double* sink = (double*)calloc(wrow, sizeof(double));
   for (int e = 0; e < nE; e++) {
      int quirk = (model == 2 && e < nR && tR[e]); /* transr/trainer.cpp:187 */
      if (!tE[e] && !quirk) continue;
      double* row = ent + (size_t)e * D;
      if (tE[e]) {
         for (int i = 0; i < D; i++) row[i] += dE[(size_t)e * D + i];
         orc_norm(row, D, model == 2 ? 0 : 1);
      }
      int rels[3];
      int cnt = 0;
      if (model != 0 && tE[e]) {
         rels[cnt++] = rmin[e];
         if (rmax[e] != rmin[e]) rels[cnt++] = rmax[e];
      }
      if (quirk) {
         int dup = 0;
         for (int q = 0; q < cnt; q++) dup |= (rels[q] == e);
         if (!dup) rels[cnt++] = e;
      }
      for (int q = 0; q < cnt; q++) {
         double* dst = carry ? carry + (size_t)rels[q] * wrow : sink;
         if (model == 1) soft_orth_entity(row, w + (size_t)rels[q] * D, dst, D, lr, scratch);
         else transr_norm_entity(row, w + (size_t)rels[q] * D * D, dst, D, lr, scratch);
      }
   }
   free(sink); free(scratch); free(x); free(aR); free(rmin); free(rmax); free(tE); free(tR); free(dE); free(dR); free(dW);
   if (n_active) *n_active = active;
   return total;
}

/* Study switch (tools/stat_parity_cpu.py): drop the perturbation of w_r / M_r that the entity-side constraint steps
 * hand to the next batch, to measure what that carry is worth. */
static int g_dfr_no_carry = 0;
void orc_set_dfr_no_carry(int on) { g_dfr_no_carry = on; }

void orc_train_epochs_dfr(const orc_sampler* s, int model, int distance, int D, int nE, int nR,
                          double lr, double margin, int batches, int first_epoch, int epochs, uint64_t seed,
                          double* ent, double* rel, double* w, double* loss_out) {
   long batchsize = s->n / batches; /* common/trainer.cpp:70 */
   int* pairs = (int*)malloc(sizeof(int) * 6 * (size_t)(batchsize > 0 ? batchsize : 1));
   size_t we = w_elems(model, D, nR);
   double* carry = g_dfr_no_carry ? NULL : (double*)calloc(we ? we : 1, sizeof(double));
   for (int e = 0; e < epochs; e++) {
      double loss = 0;
      for (int b = 0; b < batches; b++) {
         uint32_t gb = (uint32_t)(first_epoch + e) * (uint32_t)batches + (uint32_t)b;
         orc_sample_batch(s, seed, gb, batchsize, pairs);
         loss += orc_train_batch_dfr(model, distance, D, nE, nR, lr, margin, ent, rel, w, carry, batchsize, pairs, NULL);
      }
      if (loss_out) loss_out[e] = loss;
   }
   free(carry);
   free(pairs);
}

/* epochs x batches of orc_sample_batch + orc_train_batch_ref: the REFERENCE's sequential batch semantics
 * (bitwise-pinned above) driven by the uniform counter sampler instead of the reference's randMax
 * (common/utils.cpp:113-120, whose int-overflowing product of two rand() values is heavily non-uniform).
 * Separates "what the sampler changes" from "what the deferred renormalisation changes" in the
 * trained-model parity study (tools/stat_parity.py). */
void orc_train_epochs_ref(const orc_sampler* s, int model, int distance, int D, int nE, int nR,
                          double lr, double margin, int batches, int first_epoch, int epochs, uint64_t seed,
                          double* ent, double* rel, double* w, double* loss_out) {
   long batchsize = s->n / batches;
   int* pairs = (int*)malloc(sizeof(int) * 6 * (size_t)(batchsize > 0 ? batchsize : 1));
   size_t we = w_elems(model, D, nR);
   double* entN = (double*)malloc(sizeof(double) * (size_t)nE * D);
   double* relN = (double*)malloc(sizeof(double) * (size_t)nR * D);
   double* wN = (double*)malloc(sizeof(double) * (we ? we : 1));
   for (int e = 0; e < epochs; e++) {
      double loss = 0;
      for (int b = 0; b < batches; b++) {
         uint32_t gb = (uint32_t)(first_epoch + e) * (uint32_t)batches + (uint32_t)b;
         orc_sample_batch(s, seed, gb, batchsize, pairs);
         loss += orc_train_batch_ref(model, distance, D, nE, nR, lr, margin, ent, rel, w, batchsize, pairs, entN, relN, wN, NULL);
         /* postbatch: cur = next (transe/trainer.cpp:48-51) */
         memcpy(ent, entN, sizeof(double) * (size_t)nE * D);
         memcpy(rel, relN, sizeof(double) * (size_t)nR * D);
         if (we) memcpy(w, wN, sizeof(double) * we);
      }
      if (loss_out) loss_out[e] = loss;
   }
   free(entN); free(relN); free(wN); free(pairs);
}
