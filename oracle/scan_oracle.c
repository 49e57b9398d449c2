/* CPU restatement (plain C + OpenMP) of the reference's fused ADC scan + top-k.
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY -- see oracle/__init__.py.
 *
 * Follows torchpq/kernels/cuda/ivfpq_topk.cu:822-971 (which cells are scanned: first nProbe
 * entries, cell 0 always, equal-start skip :864-866, range [start, start+size), isEmpty mask
 * :878-884), consume_data :662-679 (fp32 sum, m ascending from 0.f), write-out :966-970 and
 * IVFPQTopkCuda.py:118-120,142 ((-inf,-1) padding).  Tie rule: (score desc, address asc).
 * One query per OpenMP task; layouts are the reference's own (storage [M/4, cap, 4], LUT [M, nq, 256]).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { float s; int64_t a; } cand_t;

/* "worse" ordering for a min-heap whose root is the worst kept candidate */
static inline int worse(const cand_t* x, const cand_t* y) {
  return (x->s < y->s) || (x->s == y->s && x->a > y->a);
}
static void sift_down(cand_t* h, int n, int i) {
  for (;;) {
    int l = 2 * i + 1, r = l + 1, m = i;
    if (l < n && worse(&h[l], &h[m])) m = l;
    if (r < n && worse(&h[r], &h[m])) m = r;
    if (m == i) return;
    cand_t t = h[i]; h[i] = h[m]; h[m] = t; i = m;
  }
}
static void sift_up(cand_t* h, int i) {
  while (i > 0) {
    int p = (i - 1) / 2;
    if (!worse(&h[i], &h[p])) return;
    cand_t t = h[i]; h[i] = h[p]; h[p] = t; i = p;
  }
}
static int cmp_best_first(const void* a, const void* b) {
  const cand_t* x = (const cand_t*)a; const cand_t* y = (const cand_t*)b;
  if (worse(y, x)) return -1;
  if (worse(x, y)) return 1;
  return 0;
}

int oracle_ivfpq_topk(const uint8_t* storage, const float* lut, const uint8_t* is_empty,
                      const int64_t* cell_start_g, const int64_t* cell_size_g, const int64_t* n_probe_list,
                      int64_t cap, int M, int nq, int n_probe, int k,
                      float* vals, int64_t* addr, int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  int used = 1;
#pragma omp parallel
  {
#ifdef _OPENMP
#pragma omp single
    used = omp_get_num_threads();
#endif
    float* t = (float*)malloc((size_t)M * 256 * sizeof(float));
    cand_t* heap = (cand_t*)malloc((size_t)k * sizeof(cand_t));
#pragma omp for schedule(dynamic, 1)
    for (int q = 0; q < nq; ++q) {
      for (int m = 0; m < M; ++m) memcpy(t + (size_t)m * 256, lut + ((size_t)m * nq + q) * 256, 256 * sizeof(float));
      int P = (int)n_probe_list[q];              /* (int) truncation, ivfpq_topk.cu:837 */
      if (P > n_probe) P = n_probe;
      if (P < 1) P = 1;
      int hn = 0;
      int64_t prev_start = 0;
      for (int j = 0; j < P; ++j) {
        const int64_t s = cell_start_g[(size_t)q * n_probe + j];
        const int64_t n = cell_size_g[(size_t)q * n_probe + j];
        if (j > 0 && s == prev_start) continue;
        prev_start = s;
        for (int64_t a = s; a < s + n; ++a) {
          if (is_empty[a]) continue;
          float acc = 0.f;
          for (int m = 0; m < M; ++m) {
            const uint8_t c = storage[((size_t)(m >> 2) * cap + a) * 4 + (m & 3)];
            acc += t[(size_t)m * 256 + c];
          }
          cand_t c = {acc, a};
          if (hn < k) { heap[hn] = c; sift_up(heap, hn); ++hn; }
          else if (worse(&heap[0], &c)) { heap[0] = c; sift_down(heap, hn, 0); }
        }
      }
      qsort(heap, hn, sizeof(cand_t), cmp_best_first);
      for (int i = 0; i < k; ++i) {
        vals[(size_t)q * k + i] = i < hn ? heap[i].s : -INFINITY;
        addr[(size_t)q * k + i] = i < hn ? heap[i].a : -1;
      }
    }
    free(t); free(heap);
  }
  return used;
}
