/* CPU oracle in plain C -- TEST INFRASTRUCTURE, see oracle/__init__.py.
 *
 * A second, independent restatement of the hot path (the first is numpy in
 * oracle/flat_ip.py / oracle/maxsim.py); tests cross-check the two, and
 * bench.py may time this one as the "port" CPU baseline.  "parity unpinned"
 * for the FAISS arithmetic (faiss-cpu is an absent, unpinned dependency);
 * Stage 2 is pinned by tests/golden/stage2_reference.json.
 *
 *   oracle_flat_ip_search  <- faiss.IndexFlatIP.search as called at
 *                             /root/reference/src/stage1_retriever.py:380
 *   oracle_normalize_rows  <- src/stage1_retriever.py:285-288
 *   oracle_maxsim          <- src/stage2_rescorer.py:167-201
 *
 * Build: make -C oracle   (gcc -O3 -fopenmp -shared -fPIC)
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { float s; int64_t id; } ent_t;

/* "a ranks before b": score desc, id asc */
static inline int before(float sa, int64_t ia, float sb, int64_t ib) {
    return sa > sb || (sa == sb && ia < ib);
}

/* min-heap on rank (root = worst kept entry) */
static void sift_down(ent_t* h, int n, int i) {
    for (;;) {
        int l = 2 * i + 1, r = l + 1, w = i;
        if (l < n && before(h[w].s, h[w].id, h[l].s, h[l].id)) w = l;
        if (r < n && before(h[w].s, h[w].id, h[r].s, h[r].id)) w = r;
        if (w == i) return;
        ent_t t = h[i]; h[i] = h[w]; h[w] = t; i = w;
    }
}
static void heap_push(ent_t* h, int* n, int k, float s, int64_t id) {
    if (*n < k) {
        int i = (*n)++;
        h[i].s = s; h[i].id = id;
        while (i > 0) {
            int p = (i - 1) / 2;
            if (before(h[p].s, h[p].id, h[i].s, h[i].id)) {
                ent_t t = h[i]; h[i] = h[p]; h[p] = t; i = p;
            } else break;
        }
    } else if (before(s, id, h[0].s, h[0].id)) {
        h[0].s = s; h[0].id = id;
        sift_down(h, k, 0);
    }
}
static int cmp_rank(const void* a, const void* b) {
    const ent_t* x = (const ent_t*)a; const ent_t* y = (const ent_t*)b;
    if (before(x->s, x->id, y->s, y->id)) return -1;
    if (before(y->s, y->id, x->s, x->id)) return 1;
    return 0;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* X[n,d], Q[B,d] fp32 row-major. D[B,k] desc, I[B,k]; (-FLT_MAX,-1) padding. */
int oracle_flat_ip_search(const float* X, int64_t n, int d, const float* Q, int B, int k,
                          float* D, int64_t* I) {
    int T = oracle_num_threads();
    ent_t* heaps = (ent_t*)malloc((size_t)T * B * k * sizeof(ent_t));
    int* cnt = (int*)calloc((size_t)T * B, sizeof(int));
    if (!heaps || !cnt) return -1;
#pragma omp parallel
    {
#ifdef _OPENMP
        int t = omp_get_thread_num();
#else
        int t = 0;
#endif
        ent_t* hp = heaps + (size_t)t * B * k;
        int* cp = cnt + (size_t)t * B;
#pragma omp for schedule(static)
        for (int64_t j = 0; j < n; ++j) {
            const float* x = X + j * d;
            for (int b = 0; b < B; ++b) {
                const float* q = Q + (size_t)b * d;
                float acc = 0.f;
                for (int t2 = 0; t2 < d; ++t2) acc += q[t2] * x[t2];
                heap_push(hp + (size_t)b * k, cp + b, k, acc, j);
            }
        }
    }
    ent_t* all = (ent_t*)malloc((size_t)T * k * sizeof(ent_t));
    for (int b = 0; b < B; ++b) {
        int m = 0;
        for (int t = 0; t < T; ++t) {
            int c = cnt[(size_t)t * B + b];
            memcpy(all + m, heaps + ((size_t)t * B + b) * k, c * sizeof(ent_t));
            m += c;
        }
        qsort(all, m, sizeof(ent_t), cmp_rank);
        for (int r = 0; r < k; ++r) {
            if (r < m) { D[(size_t)b * k + r] = all[r].s; I[(size_t)b * k + r] = all[r].id; }
            else { D[(size_t)b * k + r] = -FLT_MAX; I[(size_t)b * k + r] = -1; }
        }
    }
    free(all); free(heaps); free(cnt);
    return 0;
}

/* in place: x / (|x| + 1e-8), norm accumulated in fp32 like numpy's sqrt(sum(x*x)) */
void oracle_normalize_rows(float* X, int64_t n, int d) {
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < n; ++j) {
        float* x = X + j * d;
        double ss = 0.0;
        for (int t = 0; t < d; ++t) ss += (double)x[t] * x[t];
        float inv = 1.0f / ((float)sqrt(ss) + 1e-8f);
        for (int t = 0; t < d; ++t) x[t] = x[t] * inv;
    }
}

/* q[Lq,H], d[Ld,H] raw hidden states; mode 0 = maxsim (mean of row max),
 * 1 = colbert (softmax-weighted sum).  F.normalize eps = 1e-12. */
float oracle_maxsim(const float* q, int Lq, const float* dtok, int Ld, int H, int mode) {
    float* dn = (float*)malloc((size_t)Ld * sizeof(float));
    float* m = (float*)malloc((size_t)Lq * sizeof(float));
    for (int j = 0; j < Ld; ++j) {
        float ss = 0.f;
        for (int t = 0; t < H; ++t) ss += dtok[(size_t)j * H + t] * dtok[(size_t)j * H + t];
        dn[j] = fmaxf(sqrtf(ss), 1e-12f);
    }
    for (int i = 0; i < Lq; ++i) {
        float ss = 0.f;
        for (int t = 0; t < H; ++t) ss += q[(size_t)i * H + t] * q[(size_t)i * H + t];
        float qn = fmaxf(sqrtf(ss), 1e-12f);
        float best = -FLT_MAX;
        for (int j = 0; j < Ld; ++j) {
            float acc = 0.f;
            for (int t = 0; t < H; ++t)
                acc += (q[(size_t)i * H + t] / qn) * (dtok[(size_t)j * H + t] / dn[j]);
            if (acc > best) best = acc;
        }
        m[i] = best;
    }
    float out;
    if (mode == 0) {
        float s = 0.f;
        for (int i = 0; i < Lq; ++i) s += m[i];
        out = s / (float)Lq;
    } else {
        float mx = -FLT_MAX, z = 0.f, s = 0.f;
        for (int i = 0; i < Lq; ++i) if (m[i] > mx) mx = m[i];
        for (int i = 0; i < Lq; ++i) z += expf(m[i] - mx);
        for (int i = 0; i < Lq; ++i) s += m[i] * (expf(m[i] - mx) / z);
        out = s;
    }
    free(dn); free(m);
    return out;
}

/* per-candidate loop of rescore_candidates (src/stage2_rescorer.py:268-273):
 * tok = concatenated [sum L, H] rows, off[n+1] row offsets. */
void oracle_maxsim_batch(const float* q, int Lq, const float* tok, const int64_t* off, int n,
                         int H, int mode, float* out) {
#pragma omp parallel for schedule(dynamic, 8)
    for (int c = 0; c < n; ++c)
        out[c] = oracle_maxsim(q, Lq, tok + off[c] * H, (int)(off[c + 1] - off[c]), H, mode);
}
