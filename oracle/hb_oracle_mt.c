/* hb_oracle_mt.c -- pthread work-queue driver over the oracle's per-feature
 * functions (TEST INFRASTRUCTURE: bench.py CPU legs only; see hb_oracle.h). */
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>

#include "hb_oracle.h"

typedef struct {
  int kind;        /* 0 forward chunk, 1 backward part */
  int feat;
  int64_t a, b;    /* forward: bag range [a,b) ; backward: part a of b */
} task_t;

typedef struct {
  const hbo_mt_feature* feats;
  task_t* tasks;
  int ntasks;
  atomic_int next;
  atomic_int err;
  float lr;
} job_t;

static int run_task(const job_t* J, const task_t* t) {
  const hbo_mt_feature* f = &J->feats[t->feat];
  if (t->kind == 0) {
    const int64_t n = t->b - t->a;
    int64_t* off = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n + 1));
    if (!off) return 2;
    for (int64_t i = 0; i <= n; ++i) off[i] = i;
    int rc = hbo_embedding_lookup_sparse(f->table, f->rows, f->dim, f->ids + t->a, off, n, HBO_MEAN,
                                         f->out + t->a * f->out_stride, f->out_stride);
    free(off);
    return rc;
  }
  /* backward part: rows with row % parts == part */
  const int64_t parts = t->b, part = t->a;
  int64_t cnt = 0;
  for (int64_t i = 0; i < f->nbags; ++i) cnt += (f->ids[i] % parts == part);
  if (cnt == 0) return 0;
  int64_t* sel = (int64_t*)malloc(sizeof(int64_t) * (size_t)cnt);
  float* g = (float*)malloc(sizeof(float) * (size_t)cnt * (size_t)f->dim);
  if (!sel || !g) { free(sel); free(g); return 2; }
  int64_t k = 0;
  for (int64_t i = 0; i < f->nbags; ++i)
    if (f->ids[i] % parts == part) {
      sel[k] = f->ids[i];
      /* mean combiner, one id per bag: row gradient == bag gradient */
      memcpy(g + k * f->dim, f->grad + i * f->grad_stride, sizeof(float) * (size_t)f->dim);
      ++k;
    }
  int rc = hbo_sparse_apply_adagrad(f->table, f->accum, f->rows, f->dim, sel, g, cnt, J->lr);
  free(sel);
  free(g);
  return rc;
}

static void* worker(void* arg) {
  job_t* J = (job_t*)arg;
  for (;;) {
    const int i = atomic_fetch_add(&J->next, 1);
    if (i >= J->ntasks) break;
    const int rc = run_task(J, &J->tasks[i]);
    if (rc) atomic_store(&J->err, rc);
  }
  return NULL;
}

static int run_phase(job_t* J, int nthreads) {
  if (J->ntasks == 0) return 0;
  if (nthreads > J->ntasks) nthreads = J->ntasks;
  if (nthreads < 1) nthreads = 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
  if (!th) return 2;
  atomic_store(&J->next, 0);
  atomic_store(&J->err, 0);
  for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], NULL, worker, J);
  for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
  free(th);
  return atomic_load(&J->err);
}

int hbo_mt_step(int nfeat, const hbo_mt_feature* feats, int64_t fwd_chunk, float lr, int nthreads) {
  if (nfeat < 1 || !feats || fwd_chunk < 1) return 1;
  /* forward tasks */
  int cap = 0;
  for (int k = 0; k < nfeat; ++k) cap += (int)((feats[k].nbags + fwd_chunk - 1) / fwd_chunk) + feats[k].parts + 1;
  task_t* tasks = (task_t*)malloc(sizeof(task_t) * (size_t)(cap > 0 ? cap : 1));
  if (!tasks) return 2;
  job_t J;
  J.feats = feats;
  J.tasks = tasks;
  J.lr = lr;
  int n = 0;
  for (int k = 0; k < nfeat; ++k)
    for (int64_t a = 0; a < feats[k].nbags; a += fwd_chunk) {
      tasks[n].kind = 0; tasks[n].feat = k; tasks[n].a = a;
      tasks[n].b = a + fwd_chunk < feats[k].nbags ? a + fwd_chunk : feats[k].nbags;
      ++n;
    }
  J.ntasks = n;
  int rc = run_phase(&J, nthreads);
  if (rc) { free(tasks); return rc; }
  /* backward tasks (the forward of ALL features is complete: tables are read by it) */
  n = 0;
  for (int k = 0; k < nfeat; ++k) {
    if (!feats[k].accum) continue;
    const int parts = feats[k].parts > 0 ? feats[k].parts : 1;
    for (int p = 0; p < parts; ++p) { tasks[n].kind = 1; tasks[n].feat = k; tasks[n].a = p; tasks[n].b = parts; ++n; }
  }
  J.ntasks = n;
  rc = run_phase(&J, nthreads);
  free(tasks);
  return rc;
}
