/*
 * hnsw_oracle.c -- see hnsw_oracle.h.  TEST INFRASTRUCTURE ONLY.
 */
#include "hnsw_oracle.h"
#include "hnsw_build_internal.h"

#include <stdlib.h>
#include <string.h>

static inline uint32_t g_nbrs(const orc_graph *g, uint32_t level) {
    return level == 0 ? g->level0_neighbors_count : g->neighbors_count;
}
static inline uint32_t g_id(const orc_graph *g, uint32_t row) { return row < g->n ? row : ORC_ROOT_ID; }
static inline const uint8_t *g_code(const orc_graph *g, uint32_t row) {
    return (const uint8_t *)g->codes + (size_t)row * orc_code_bytes(g->storage_type, g->dim);
}
static inline uint64_t mk_key(const orc_graph *g, float score, uint32_t id) {
    return ((uint64_t)orc_order_key(g->metric, score) << 32) | (uint64_t)(~id);
}

/* PerformantFixedSet (src/models/fixedset.rs:13-28); len = number of u64 buckets */
static inline void fs_insert(uint64_t *b, uint32_t len, uint32_t v) {
    uint32_t mask = len - 1u;
    b[(v >> 6) & mask] |= 1ull << (v & 0x3f);
}
static inline int fs_member(const uint64_t *b, uint32_t len, uint32_t v) {
    uint32_t mask = len - 1u;
    return (b[(v >> 6) & mask] >> (v & 0x3f)) & 1ull;
}

typedef struct { uint64_t key; uint32_t node; float score; } hitem;

static void heap_push(hitem *h, size_t *n, hitem it) {
    size_t i = (*n)++;
    h[i] = it;
    while (i > 0) {
        size_t p = (i - 1) / 2;
        if (h[p].key >= h[i].key) break;
        hitem t = h[p]; h[p] = h[i]; h[i] = t;
        i = p;
    }
}
static hitem heap_pop(hitem *h, size_t *n) {
    hitem top = h[0];
    h[0] = h[--(*n)];
    size_t i = 0;
    for (;;) {
        size_t l = 2 * i + 1, r = l + 1, m = i;
        if (l < *n && h[l].key > h[m].key) m = l;
        if (r < *n && h[r].key > h[m].key) m = r;
        if (m == i) break;
        hitem t = h[m]; h[m] = h[i]; h[i] = t;
        i = m;
    }
    return top;
}
static int cmp_hitem_desc(const void *a, const void *b) {
    uint64_t x = ((const hitem *)a)->key, y = ((const hitem *)b)->key;
    return (x < y) - (x > y);
}

/* src/vector_store.rs:1112-1204 */
int orc_traverse(const orc_graph *g, uint32_t level, uint32_t entry, const void *qcode, float qmag,
                 const orc_traverse_params *p, uint64_t *fs, uint32_t *out_nodes, float *out_scores, uint32_t *out_n,
                 uint64_t *evals, uint64_t *pops) {
    const uint32_t nb = g_nbrs(g, level);
    const uint32_t *adj = g->adj[level];
    const uint32_t *rows = g->node_row[level];
    const uint32_t take = p->shortlist_size < nb ? p->shortlist_size : nb;
    size_t hcap = (size_t)p->ef * take + 2, hn = 0, rn = 0;
    hitem *heap = (hitem *)malloc(sizeof(hitem) * hcap);
    hitem *res = (hitem *)malloc(sizeof(hitem) * ((size_t)p->ef + 1));
    int rc = ORC_OK;
    float d;
    uint32_t erow = rows[entry];
    rc = orc_distance(g->metric, g->storage_type, g->dim, qcode, qmag, g_code(g, erow), g->mags[erow], &d);
    if (evals) (*evals)++;
    if (rc != ORC_OK) goto done;
    fs_insert(fs, nb, g_id(g, erow));
    heap_push(heap, &hn, (hitem){mk_key(g, d, g_id(g, erow)), entry, d});
    uint32_t visited = 0;
    while (hn > 0) {
        hitem cur = heap_pop(heap, &hn);
        if (visited >= p->ef) break;
        visited++;
        if (pops) (*pops)++;
        res[rn++] = cur;
        const uint32_t *slots = adj + (size_t)cur.node * nb;
        for (uint32_t s = 0; s < take; ++s) {
            uint32_t nbl = slots[s];
            if (nbl == ORC_EMPTY) continue;
            uint32_t nrow = rows[nbl], nid = g_id(g, nrow);
            if (fs_member(fs, nb, nid)) continue;
            rc = orc_distance(g->metric, g->storage_type, g->dim, qcode, qmag, g_code(g, nrow), g->mags[nrow], &d);
            if (evals) (*evals)++;
            if (rc != ORC_OK) goto done;
            fs_insert(fs, nb, nid);
            heap_push(heap, &hn, (hitem){mk_key(g, d, nid), nbl, d});
        }
    }
    qsort(res, rn, sizeof(hitem), cmp_hitem_desc);
    if (rn > p->final_len) rn = p->final_len;
    for (size_t i = 0; i < rn; ++i) { out_nodes[i] = res[i].node; out_scores[i] = res[i].score; }
    *out_n = (uint32_t)rn;
done:
    free(heap);
    free(res);
    return rc;
}

/* src/vector_store.rs:256-402 (no metadata filter) */
int orc_ann_search(const orc_graph *g, const void *qcode, float qmag, uint32_t ef_search, uint32_t shortlist_size,
                   uint32_t *out_rows, float *out_scores, size_t cap, size_t *out_n, uint64_t *evals, uint64_t *pops) {
    uint32_t entry = g->entry;
    size_t total = 0;
    uint32_t nodes[128];
    float scores[128];
    orc_traverse_params p = {ef_search, shortlist_size, 100, ORC_QUERY_ID};
    uint32_t maxnb = g->level0_neighbors_count > g->neighbors_count ? g->level0_neighbors_count : g->neighbors_count;
    uint64_t *fs = (uint64_t *)malloc(sizeof(uint64_t) * maxnb);
    int rc = ORC_OK;
    for (int level = (int)g->num_levels; level >= 0; --level) {
        uint32_t nb = g_nbrs(g, (uint32_t)level), zn = 0;
        memset(fs, 0, sizeof(uint64_t) * nb);
        fs_insert(fs, nb, p.self_id);
        rc = orc_traverse(g, (uint32_t)level, entry, qcode, qmag, &p, fs, nodes, scores, &zn, evals, pops);
        if (rc != ORC_OK) break;
        if (zn == 0) { /* vector_store.rs:329-380: fall back to the entry node itself */
            uint32_t erow = g->node_row[level][entry];
            rc = orc_distance(g->metric, g->storage_type, g->dim, qcode, qmag, g_code(g, erow), g->mags[erow], &scores[0]);
            if (rc != ORC_OK) break;
            nodes[0] = entry;
            zn = 1;
        }
        for (uint32_t i = 0; i < zn && total < cap; ++i) {
            out_rows[total] = g->node_row[level][nodes[i]];
            out_scores[total] = scores[i];
            total++;
        }
        if (level > 0) entry = g->child[level][nodes[0]];
    }
    free(fs);
    *out_n = total;
    return rc;
}

/* src/models/common.rs:381-412 */
size_t orc_dedup_filter(const orc_graph *g, uint32_t *rows, float *scores, size_t n, size_t k) {
    hitem *tmp = (hitem *)malloc(sizeof(hitem) * (n ? n : 1));
    size_t m = 0;
    for (size_t i = 0; i < n; ++i) {
        int seen = 0;
        for (size_t j = 0; j < i; ++j)
            if (rows[j] == rows[i]) { seen = 1; break; }
        if (seen) continue;
        if (g_id(g, rows[i]) == ORC_ROOT_ID) continue;
        tmp[m++] = (hitem){mk_key(g, scores[i], rows[i]), rows[i], scores[i]};
    }
    qsort(tmp, m, sizeof(hitem), cmp_hitem_desc);
    if (m > 5 * k) m = 5 * k;
    for (size_t i = 0; i < m; ++i) { rows[i] = tmp[i].node; scores[i] = tmp[i].score; }
    free(tmp);
    return m;
}

/* search_internal (src/indexes/hnsw/mod.rs:390-440) for a batch; one query per thread */
int orc_hnsw_search_batch(const orc_graph *g, const float *raw, const float *queries, size_t nq, float lo, float hi,
                          uint32_t ef_search, uint32_t shortlist_size, size_t k, int threads, uint32_t *out_ids,
                          float *out_scores, uint32_t *out_counts, uint8_t *err, uint64_t *evals, uint64_t *pops) {
    if (threads < 1) threads = 1;
    const size_t cb = orc_code_bytes(g->storage_type, g->dim);
    const size_t cap = ((size_t)g->num_levels + 1) * 100;
    uint64_t ev_total = 0, pop_total = 0;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1) reduction(+ : ev_total, pop_total)
    for (long long qi = 0; qi < (long long)nq; ++qi) {
        const float *q = queries + (size_t)qi * g->dim;
        uint8_t *qcode = (uint8_t *)malloc(cb ? cb : 1);
        float qmag;
        orc_quantize(g->storage_type, lo, hi, q, g->dim, qcode, &qmag);
        uint32_t *rows = (uint32_t *)malloc(sizeof(uint32_t) * cap);
        float *sc = (float *)malloc(sizeof(float) * cap);
        size_t n = 0;
        uint64_t ev = 0, pp = 0;
        int rc = orc_ann_search(g, qcode, qmag, ef_search, shortlist_size, rows, sc, cap, &n, &ev, &pp);
        ev_total += ev;
        pop_total += pp;
        for (size_t j = 0; j < k; ++j) { out_ids[(size_t)qi * k + j] = 0xFFFFFFFFu; out_scores[(size_t)qi * k + j] = 0.0f; }
        if (rc != ORC_OK) {
            if (err) err[qi] = (uint8_t)(rc == ORC_CALCULATION_ERROR ? 1 : 2);
            if (out_counts) out_counts[qi] = 0;
        } else {
            if (err) err[qi] = 0;
            size_t m = orc_dedup_filter(g, rows, sc, n, k);
            orc_rerank_f32(raw, g->dim, q, rows, m, k, out_ids + (size_t)qi * k, out_scores + (size_t)qi * k);
            if (out_counts) out_counts[qi] = (uint32_t)(m < k ? m : k);
        }
        free(qcode);
        free(rows);
        free(sc);
    }
    if (evals) *evals = ev_total;
    if (pops) *pops = pop_total;
    return ORC_OK;
}

/* ------------------------------------------------------------------ builder */
struct orc_built {
    orc_graph g;
    uint32_t nlevels1;
    blevel *lv;
    uint32_t *cnt_arr;
    const uint32_t **node_row_arr, **adj_arr, **child_arr;
    uint32_t min_key, max_key; /* MetricResult::min / ::max as order keys (src/models/types.rs:435-457) */
};

static void refresh_view(orc_built *b) {
    for (uint32_t L = 0; L < b->nlevels1; ++L) {
        b->cnt_arr[L] = b->lv[L].cnt;
        b->node_row_arr[L] = b->lv[L].node_row;
        b->adj_arr[L] = b->lv[L].adj;
        b->child_arr[L] = b->lv[L].child;
    }
}
static int add_neighbor(orc_built *b, blevel *l, uint32_t node, uint32_t nbr, uint32_t dkey) {
    return bl_add_neighbor(b->min_key, b->max_key, l, node, nbr, dkey);
}

/* create_node_edges (src/vector_store.rs:976-1070), Base nodes only */
static void create_node_edges(orc_built *b, blevel *l, uint32_t node, const uint32_t *znodes, const float *zscores, uint32_t zn) {
    uint32_t successful = 0;
    for (uint32_t i = 0; i < zn; ++i) {
        if (successful >= l->nb) break;
        uint32_t nbr = znodes[i];
        uint32_t dkey = orc_order_key(b->g.metric, zscores[i]);
        int idx = add_neighbor(b, l, node, nbr, dkey);
        if (idx >= 0) {
            int j = add_neighbor(b, l, nbr, node, dkey);
            if (j >= 0) successful++;
            else if (l->adj[(size_t)node * l->nb + idx] == nbr) l->adj[(size_t)node * l->nb + idx] = ORC_EMPTY; /* remove_neighbor_by_index_and_id */
        }
    }
}

static float rng_unit(uint64_t seed, uint64_t i) { /* splitmix64 -> f32 in [0,1), stands in for rand::random::<f32>() */
    uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}

orc_built *orc_hnsw_build(int metric, int st, size_t dim, const void *codes, const float *mags, uint32_t n,
                          uint32_t num_levels, uint32_t nbrs, uint32_t nbrs0, uint32_t ef_construction,
                          uint32_t shortlist, uint64_t seed) {
    orc_built *b = (orc_built *)calloc(1, sizeof(orc_built));
    b->nlevels1 = num_levels + 1;
    b->lv = (blevel *)calloc(b->nlevels1, sizeof(blevel));
    b->cnt_arr = (uint32_t *)calloc(b->nlevels1, sizeof(uint32_t));
    b->node_row_arr = (const uint32_t **)calloc(b->nlevels1, sizeof(uint32_t *));
    b->adj_arr = (const uint32_t **)calloc(b->nlevels1, sizeof(uint32_t *));
    b->child_arr = (const uint32_t **)calloc(b->nlevels1, sizeof(uint32_t *));
    orc_graph *g = &b->g;
    g->num_levels = num_levels; g->neighbors_count = nbrs; g->level0_neighbors_count = nbrs0; g->n = n;
    g->metric = metric; g->storage_type = st; g->dim = dim; g->codes = codes; g->mags = mags;
    g->cnt = b->cnt_arr; g->node_row = b->node_row_arr; g->adj = b->adj_arr; g->child = b->child_arr;
    bl_min_max_keys(metric, &b->min_key, &b->max_key);
    /* root nodes on every level (vector_store.rs:44-140): row n, child links downwards */
    for (uint32_t L = 0; L <= num_levels; ++L) {
        blevel *l = &b->lv[L];
        l->nb = L == 0 ? nbrs0 : nbrs;
        if (L == 0) { lv_reserve(l, n + 1); l->cnt = n + 1; for (uint32_t i = 0; i <= n; ++i) lv_init_node(b->min_key, l, i, i); }
        else { lv_reserve(l, 64); l->cnt = 1; lv_init_node(b->min_key, l, 0, n); l->child[0] = (L == 1) ? n : 0; }
    }
    g->entry = num_levels == 0 ? n : 0;
    /* level 0 nodes exist as slots for every row but are only linked once inserted */
    uint32_t *present0 = (uint32_t *)calloc((size_t)n + 1, sizeof(uint32_t));
    present0[n] = 1;
    uint32_t maxnb = nbrs0 > nbrs ? nbrs0 : nbrs;
    uint64_t *fs = (uint64_t *)malloc(sizeof(uint64_t) * maxnb);
    uint32_t *znodes = (uint32_t *)malloc(sizeof(uint32_t) * 64 * (num_levels + 1));
    float *zscores = (float *)malloc(sizeof(float) * 64 * (num_levels + 1));
    uint32_t *zn = (uint32_t *)malloc(sizeof(uint32_t) * (num_levels + 1));
    uint32_t *node_at = (uint32_t *)malloc(sizeof(uint32_t) * (num_levels + 1));
    const size_t cb = orc_code_bytes(st, dim);
    for (uint32_t r = 0; r < n; ++r) {
        /* get_max_insert_level (common.rs:373-379) with probs 1 - 4^-n (api_service.rs:109,132) */
        double x = (double)rng_unit(seed, r);
        uint32_t max_level = 0;
        for (int lv = (int)num_levels; lv >= 0; --lv) {
            double p4 = 1.0;
            for (int e = 0; e < lv; ++e) p4 *= 4.0;
            if (x >= 1.0 - 1.0 / p4) { max_level = (uint32_t)lv; break; }
        }
        const uint8_t *code = (const uint8_t *)codes + (size_t)r * cb;
        uint32_t entry = g->entry, parent = ORC_EMPTY;
        int failed = 0;
        orc_traverse_params p = {ef_construction, shortlist, 64, r};
        for (int level = (int)num_levels; level >= 0; --level) {
            blevel *l = &b->lv[level];
            refresh_view(b);
            memset(fs, 0, sizeof(uint64_t) * l->nb);
            fs_insert(fs, l->nb, r);
            uint32_t *zl = znodes + 64 * level;
            float *sl = zscores + 64 * level;
            uint32_t cnt = 0;
            int rc = orc_traverse(g, (uint32_t)level, entry, code, mags[r], &p, fs, zl, sl, &cnt, NULL, NULL);
            if (rc == ORC_OK && cnt == 0) {
                uint32_t erow = l->node_row[entry];
                rc = orc_distance(metric, st, dim, code, mags[r], (const uint8_t *)codes + (size_t)erow * cb, mags[erow], &sl[0]);
                zl[0] = entry;
                cnt = 1;
            }
            if (rc != ORC_OK) { failed = 1; break; } /* the reference returns Err: embedding not indexed */
            zn[level] = cnt;
            uint32_t next_entry = level > 0 ? l->child[zl[0]] : 0;
            if ((uint32_t)level <= max_level) {
                uint32_t idx;
                if (level == 0) { idx = r; present0[r] = 1; }
                else { lv_reserve(l, l->cnt + 1); idx = l->cnt++; lv_init_node(b->min_key, l, idx, r); }
                if (parent != ORC_EMPTY) b->lv[level + 1].child[parent] = idx;
                node_at[level] = idx;
                parent = idx;
            }
            entry = next_entry;
        }
        if (failed) continue; /* nodes created above stay unlinked, like a failed reference insert */
        uint32_t top = max_level < num_levels ? max_level : num_levels;
        for (uint32_t level = 0; level <= top; ++level)
            create_node_edges(b, &b->lv[level], node_at[level], znodes + 64 * level, zscores + 64 * level, zn[level]);
    }
    refresh_view(b);
    free(present0); free(fs); free(znodes); free(zscores); free(zn); free(node_at);
    return b;
}

const orc_graph *orc_built_graph(const orc_built *b) { return &b->g; }

void orc_built_free(orc_built *b) {
    if (!b) return;
    for (uint32_t L = 0; L < b->nlevels1; ++L) lv_free(&b->lv[L]);
    free(b->lv); free(b->cnt_arr); free(b->node_row_arr); free(b->adj_arr); free(b->child_arr);
    free(b);
}
