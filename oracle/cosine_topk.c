/*
 * C restatement of the reference cosine + a scalar brute-force top-k.  TEST INFRASTRUCTURE
 * ONLY (see oracle/__init__.py): only tests/, __graft_entry__.smoke() and bench.py's CPU legs
 * may load this.
 *
 * orc_cosine() follows server/services/correlation/strategies/similarity.py:84-98:
 *   len mismatch is the caller's business (one length here); n == 0 -> 0.0 (:87-88);
 *   dot / (|a| |b|) in double (:90-97); zero norm -> 0.0 (:94-95); clamp to [0,1] (:98).
 * orc_topk() is the flat scan a cosine index performs below Weaviate's flatSearchCutoff
 * (call sites: server/routes/knowledge_base/weaviate_client.py:252-259), ordered by
 * (score desc, id asc).  Pinned by tests/test_oracle_c.py against the golden vectors the real
 * reference function produced (tests/golden/cosine_ref.json).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

double orc_cosine(const double* a, const double* b, int n, int clamp) {
  if (n <= 0) return 0.0;
  double ab = 0.0, aa = 0.0, bb = 0.0;
  for (int i = 0; i < n; ++i) { ab += a[i] * b[i]; aa += a[i] * a[i]; bb += b[i] * b[i]; }
  const double na = sqrt(aa), nb = sqrt(bb);
  if (na == 0.0 || nb == 0.0) return 0.0;
  double c = ab / (na * nb);
  if (clamp) { if (c < 0.0) c = 0.0; if (c > 1.0) c = 1.0; }
  return c;
}

typedef struct { double s; int64_t id; } cand_t;

static int better(double s, int64_t id, const cand_t* c) { return s > c->s || (s == c->s && id < c->id); }

/* Q [nq,d], C [n,d] fp32 row-major; ids [n] or NULL (row numbers); out_ids / out_scores [nq,k],
 * padded with -1 / -INFINITY.  Insertion into a sorted k-array: O(n k) worst case, fine for tests. */
void orc_topk(const float* Q, const float* C, const int64_t* ids, int nq, int64_t n, int d, int k,
              int64_t* out_ids, float* out_scores) {
  cand_t* top = (cand_t*)malloc(sizeof(cand_t) * (size_t)(k > 0 ? k : 1));
  for (int qi = 0; qi < nq; ++qi) {
    const float* q = Q + (size_t)qi * d;
    double qq = 0.0;
    for (int i = 0; i < d; ++i) qq += (double)q[i] * q[i];
    int cnt = 0;
    for (int64_t r = 0; r < n; ++r) {
      const float* c = C + (size_t)r * d;
      double dot = 0.0, cc = 0.0;
      for (int i = 0; i < d; ++i) { dot += (double)q[i] * c[i]; cc += (double)c[i] * c[i]; }
      const double den = sqrt(qq) * sqrt(cc);
      const double s = den > 0.0 ? dot / den : 0.0;
      const int64_t id = ids ? ids[r] : r;
      if (cnt < k || better(s, id, &top[cnt - 1])) {
        int pos = cnt < k ? cnt : k - 1;
        while (pos > 0 && better(s, id, &top[pos - 1])) { top[pos] = top[pos - 1]; --pos; }
        top[pos].s = s; top[pos].id = id;
        if (cnt < k) ++cnt;
      }
    }
    for (int t = 0; t < k; ++t) {
      out_ids[(size_t)qi * k + t] = t < cnt ? top[t].id : -1;
      out_scores[(size_t)qi * k + t] = t < cnt ? (float)top[t].s : -INFINITY;
    }
  }
  free(top);
}
