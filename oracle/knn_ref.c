/*
 * knn_ref.c — plain C restatement of the exhaustive inner-product search the reference reaches
 * through faiss (index.search / index.search_and_reconstruct, clip_retrieval/clip_back.py:362;
 * clip_retrieval/clip_filter.py:55).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): tests and
 * bench.py's CPU baseline call it, the product never does.
 *
 * Algorithm restated (FAISS IndexScalarQuantizer QT_fp16 / IndexFlatIP semantics, faiss-cpu>=1.7.2,<2,
 * reference requirements.txt:8 — not vendored, not installable here): for every query, scan all
 * codes, decode fp16 -> fp32, accumulate the inner product in fp32 (8 SIMD partial sums, as the
 * AVX2 scanner does), keep the k best in a min-heap; output sorted by score descending, ties by
 * ascending id (FAISS leaves tie order unspecified); unfilled slots id -1 / score -FLT_MAX.
 * Rows are split across pthreads (no libgomp in this image); per-thread heaps are merged at the end.
 *
 * Build: gcc -O3 -mavx2 -mf16c -mfma -pthread -shared -fPIC oracle/knn_ref.c -o oracle/_build/libknn_ref.so
 */
#include <float.h>
#include <immintrin.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

typedef struct { float s; int64_t id; } cand_t;

/* a "worse" than b: lower score, or equal score and higher id */
static inline int worse(cand_t a, cand_t b) { return a.s < b.s || (a.s == b.s && a.id > b.id); }

static void heap_push(cand_t* h, int* n, int k, cand_t c) {
  if (*n < k) {
    int i = (*n)++;
    h[i] = c;
    while (i > 0) {
      int p = (i - 1) / 2;
      if (worse(h[i], h[p])) { cand_t t = h[i]; h[i] = h[p]; h[p] = t; i = p; } else break;
    }
  } else if (worse(h[0], c)) {
    h[0] = c;
    int i = 0;
    for (;;) {
      int l = 2 * i + 1, r = l + 1, m = i;
      if (l < k && worse(h[l], h[m])) m = l;
      if (r < k && worse(h[r], h[m])) m = r;
      if (m == i) break;
      cand_t t = h[i]; h[i] = h[m]; h[m] = t; i = m;
    }
  }
}

static int cmp_best_first(const void* a, const void* b) {
  const cand_t* x = (const cand_t*)a; const cand_t* y = (const cand_t*)b;
  if (x->s > y->s) return -1;
  if (x->s < y->s) return 1;
  return (x->id > y->id) - (x->id < y->id);
}

static inline float dot_f16(const uint16_t* x, const float* q, int d) {
  __m256 acc = _mm256_setzero_ps();
  int j = 0;
  for (; j + 8 <= d; j += 8) {
    __m256 xv = _mm256_cvtph_ps(_mm_loadu_si128((const __m128i*)(x + j)));
    acc = _mm256_fmadd_ps(xv, _mm256_loadu_ps(q + j), acc);
  }
  float t[8];
  _mm256_storeu_ps(t, acc);
  float s = ((t[0] + t[4]) + (t[1] + t[5])) + ((t[2] + t[6]) + (t[3] + t[7]));
  for (; j < d; j++) s += _cvtsh_ss(x[j]) * q[j];
  return s;
}

typedef struct {
  const uint16_t* X; int64_t lo, hi; int d; const float* Q; int nq, k; cand_t* heaps; int* counts;
} job_t;

static void* scan_rows(void* arg) {
  job_t* j = (job_t*)arg;
  /* block rows so a block of codes is reused across all queries while hot in cache */
  const int64_t BLK = 256;
  for (int64_t b = j->lo; b < j->hi; b += BLK) {
    int64_t e = b + BLK < j->hi ? b + BLK : j->hi;
    for (int q = 0; q < j->nq; q++) {
      cand_t* h = j->heaps + (size_t)q * j->k;
      int* cnt = j->counts + q;
      const float* qv = j->Q + (size_t)q * j->d;
      for (int64_t r = b; r < e; r++) {
        cand_t c; c.s = dot_f16(j->X + (size_t)r * j->d, qv, j->d); c.id = r;
        if (c.s == c.s) heap_push(h, cnt, j->k, c);
      }
    }
  }
  return NULL;
}

/* X: [n, d] fp16 bits; Q: [nq, d] fp32; D: [nq, k]; I: [nq, k].  Returns threads used. */
int knn_flat_ip_f16(const uint16_t* X, int64_t n, int d, const float* Q, int nq, int k, float* D, int64_t* I,
                    int64_t id_base, int nthreads) {
  int T = nthreads > 0 ? nthreads : (int)sysconf(_SC_NPROCESSORS_ONLN);
  if (T < 1) T = 1;
  if (T > 256) T = 256;
  cand_t* heaps = (cand_t*)malloc((size_t)T * nq * k * sizeof(cand_t));
  int* counts = (int*)calloc((size_t)T * nq, sizeof(int));
  job_t* jobs = (job_t*)malloc((size_t)T * sizeof(job_t));
  pthread_t* th = (pthread_t*)malloc((size_t)T * sizeof(pthread_t));
  for (int t = 0; t < T; t++) {
    job_t j = {X, n * t / T, n * (t + 1) / T, d, Q, nq, k, heaps + (size_t)t * nq * k, counts + (size_t)t * nq};
    jobs[t] = j;
    pthread_create(&th[t], NULL, scan_rows, &jobs[t]);
  }
  for (int t = 0; t < T; t++) pthread_join(th[t], NULL);
  free(jobs); free(th);
  cand_t* all = (cand_t*)malloc((size_t)T * k * sizeof(cand_t));
  for (int q = 0; q < nq; q++) {
    int m = 0;
    for (int t = 0; t < T; t++) {
      int c = counts[(size_t)t * nq + q];
      memcpy(all + m, heaps + ((size_t)t * nq + q) * k, (size_t)c * sizeof(cand_t));
      m += c;
    }
    qsort(all, m, sizeof(cand_t), cmp_best_first);
    for (int j = 0; j < k; j++) {
      if (j < m) { D[(size_t)q * k + j] = all[j].s; I[(size_t)q * k + j] = all[j].id + id_base; }
      else { D[(size_t)q * k + j] = -FLT_MAX; I[(size_t)q * k + j] = -1; }
    }
  }
  free(all); free(heaps); free(counts);
  return T;
}

/* ---------------------------------------------------------------------------------------------------
 * IVF-Flat, inner product (FAISS IndexIVFFlat / "IVF{nlist},SQfp16" with METRIC_INNER_PRODUCT, restated):
 * per query, (1) coarse quantiser = exhaustive inner product with the nlist centroids, keep the nprobe
 * largest (ties to the lower list id); (2) scan the rows of those lists (stored list after list), exact
 * fp32-accumulated inner products, k-best heap; (3) sort: score descending, ties by ascending id.
 * FAISS parallelises IVF search over queries (OpenMP); so do the pthreads here.
 *   X: [n, d] fp16 bits in LIST ORDER; offsets: [nlist + 1]; ids: [n] id of the row at each slot;
 *   C: [nlist, d] fp16 bits (centroids as the index stores them); Q: [nq, d] fp32.
 * Returns threads used. */
typedef struct {
  const uint16_t* X; const int64_t* offsets; const int64_t* ids; const uint16_t* C; int nlist, d;
  const float* Q; int q_lo, q_hi, k, nprobe; float* D; int64_t* I; int64_t* probes_out;
} ivf_job_t;

static void* ivf_queries(void* arg) {
  ivf_job_t* j = (ivf_job_t*)arg;
  const int k = j->k, np = j->nprobe;
  cand_t* ph = (cand_t*)malloc((size_t)np * sizeof(cand_t));
  cand_t* h = (cand_t*)malloc((size_t)k * sizeof(cand_t));
  for (int q = j->q_lo; q < j->q_hi; q++) {
    const float* qv = j->Q + (size_t)q * j->d;
    int pc = 0;
    for (int l = 0; l < j->nlist; l++) {
      cand_t c; c.s = dot_f16(j->C + (size_t)l * j->d, qv, j->d); c.id = l;
      if (c.s == c.s) heap_push(ph, &pc, np, c);
    }
    qsort(ph, pc, sizeof(cand_t), cmp_best_first);
    int hc = 0;
    for (int p = 0; p < pc; p++) {
      const int64_t l = ph[p].id;
      if (j->probes_out) j->probes_out[(size_t)q * np + p] = l;
      for (int64_t r = j->offsets[l]; r < j->offsets[l + 1]; r++) {
        cand_t c; c.s = dot_f16(j->X + (size_t)r * j->d, qv, j->d); c.id = j->ids[r];
        if (c.s == c.s) heap_push(h, &hc, k, c);
      }
    }
    if (j->probes_out) for (int p = pc; p < np; p++) j->probes_out[(size_t)q * np + p] = -1;
    qsort(h, hc, sizeof(cand_t), cmp_best_first);
    for (int i = 0; i < k; i++) {
      if (i < hc) { j->D[(size_t)q * k + i] = h[i].s; j->I[(size_t)q * k + i] = h[i].id; }
      else { j->D[(size_t)q * k + i] = -FLT_MAX; j->I[(size_t)q * k + i] = -1; }
    }
  }
  free(ph); free(h);
  return NULL;
}

int knn_ivf_ip_f16(const uint16_t* X, const int64_t* offsets, const int64_t* ids, const uint16_t* C, int nlist, int d,
                   const float* Q, int nq, int k, int nprobe, float* D, int64_t* I, int64_t* probes_out, int nthreads) {
  int T = nthreads > 0 ? nthreads : (int)sysconf(_SC_NPROCESSORS_ONLN);
  if (T < 1) T = 1;
  if (T > 256) T = 256;
  if (T > nq) T = nq > 0 ? nq : 1;
  if (nprobe > nlist) nprobe = nlist;
  ivf_job_t* jobs = (ivf_job_t*)malloc((size_t)T * sizeof(ivf_job_t));
  pthread_t* th = (pthread_t*)malloc((size_t)T * sizeof(pthread_t));
  for (int t = 0; t < T; t++) {
    ivf_job_t j = {X, offsets, ids, C, nlist, d, Q, (int)((int64_t)nq * t / T), (int)((int64_t)nq * (t + 1) / T), k, nprobe,
                   D, I, probes_out};
    jobs[t] = j;
    pthread_create(&th[t], NULL, ivf_queries, &jobs[t]);
  }
  for (int t = 0; t < T; t++) pthread_join(th[t], NULL);
  free(jobs); free(th);
  return T;
}
