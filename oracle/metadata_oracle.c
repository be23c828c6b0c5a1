/*
 * metadata_oracle.c -- see metadata_oracle.h.  TEST INFRASTRUCTURE ONLY.
 */
#include "metadata_oracle.h"
#include "hnsw_build_internal.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* Metadata::from(MetadataDimensions), src/models/types.rs:111-125: sequential sum of (d as f32)^2, clamped, sqrt */
float orc_metadata_mag(const int32_t *dims, size_t m) {
    float total = 0.0f;
    for (size_t i = 0; i < m; ++i) { float x = (float)dims[i]; total += x * x; }
    if (total > FLT_MAX || total != total) total = FLT_MAX; /* total.min(f32::MAX): inf (and NaN) -> MAX */
    return sqrtf(total);
}
/* Metadata::from(&QueryFilterDimensions), src/models/types.rs:127-146 */
float orc_query_filter_mag(const int8_t *dims, size_t m) {
    float total = 0.0f;
    for (size_t i = 0; i < m; ++i) { float x = (float)dims[i]; total += x * x; }
    return sqrtf(total);
}

/* VectorData::replica_node_kind, src/models/types.rs:223-243 */
int orc_replica_kind(const orc_vector_data *v) {
    if (!v->md_bits) return ORC_KIND_BASE;
    if (v->md_mag == 0.0f) return ORC_KIND_BASE;
    if (v->has_id && v->id >= 0xFFFFFFFFu - 257u && v->id <= 0xFFFFFFFFu - 2u) return ORC_KIND_PSEUDO;
    return ORC_KIND_METADATA;
}

/* cosine_similarity_mdims, src/distance/cosine.rs:243-259: dot_product_f32 over the i32 -> f32 dims */
static int mdims_cosine(size_t m, const orc_vector_data *x, const orc_vector_data *y, float *out) {
    float *a = (float *)malloc(sizeof(float) * (m ? m : 1)), *b = (float *)malloc(sizeof(float) * (m ? m : 1));
    for (size_t i = 0; i < m; ++i) { a[i] = (float)x->md_bits[i]; b[i] = (float)y->md_bits[i]; }
    const float dp = orc_dot_f32_simd(a, b, m);
    free(a);
    free(b);
    const float den = x->md_mag * y->md_mag; /* cosine_similarity_from_dot_product, cosine.rs:223-235 */
    if (den == 0.0f) return ORC_CALCULATION_ERROR;
    *out = dp / den;
    return ORC_OK;
}

int orc_distance_md(int metric, int st, size_t dim, size_t m, const orc_vector_data *x, const orc_vector_data *y, float *out) {
    if (metric != ORC_METRIC_COSINE) return orc_distance(metric, st, dim, x->code, x->mag, y->code, y->mag, out);
    const int xk = orc_replica_kind(x), yk = orc_replica_kind(y);
    /* match (y_kind, x_kind), cosine.rs:45-100 */
    if (yk == ORC_KIND_PSEUDO && xk == ORC_KIND_PSEUDO) return mdims_cosine(m, x, y, out);
    if (yk == ORC_KIND_PSEUDO && xk == ORC_KIND_METADATA) {
        *out = memcmp(x->md_bits, y->md_bits, sizeof(int32_t) * m) == 0 ? 1.0f : -1.0f;
        return ORC_OK;
    }
    if (yk == ORC_KIND_BASE && xk == ORC_KIND_BASE) return orc_distance(metric, st, dim, x->code, x->mag, y->code, y->mag, out);
    if (yk == ORC_KIND_METADATA && xk == ORC_KIND_METADATA) {
        float mc;
        int rc = mdims_cosine(m, x, y, &mc);
        if (rc != ORC_OK) return rc;
        if (mc > 0.99f) return orc_distance(metric, st, dim, x->code, x->mag, y->code, y->mag, out);
        *out = -1.0f;
        return ORC_OK;
    }
    if (yk == ORC_KIND_BASE && xk == ORC_KIND_METADATA) { *out = 0.0f; return ORC_OK; }
    return ORC_UNREACHABLE; /* (Pseudo,Base) (Base,Pseudo) (Metadata,Pseudo) (Metadata,Base) */
}

/* ------------------------------------------------------------------ traversal with metadata */
static inline uint32_t mg_nbrs(const orc_md_graph *mg, uint32_t level) {
    return level == 0 ? mg->g.level0_neighbors_count : mg->g.neighbors_count;
}
static inline uint64_t mkey(const orc_md_graph *mg, float score, uint32_t id) {
    return ((uint64_t)orc_order_key(mg->g.metric, score) << 32) | (uint64_t)(~id);
}
static inline void fs_insert(uint64_t *b, uint32_t len, uint32_t v) { b[(v >> 6) & (len - 1u)] |= 1ull << (v & 0x3f); }
static inline int fs_member(const uint64_t *b, uint32_t len, uint32_t v) { return (b[(v >> 6) & (len - 1u)] >> (v & 0x3f)) & 1ull; }

static void node_data(const orc_md_graph *mg, uint32_t level, uint32_t node, orc_vector_data *v) {
    const uint32_t row = mg->g.node_row[level][node];
    const uint32_t md = mg->node_md[level][node];
    v->code = (const uint8_t *)mg->g.codes + (size_t)row * orc_code_bytes(mg->g.storage_type, mg->g.dim);
    v->mag = mg->g.mags[row];
    v->has_id = 1;
    v->id = mg->node_id[level][node];
    v->md_bits = md == ORC_EMPTY ? NULL : mg->md_bits + (size_t)md * mg->md_dims;
    v->md_mag = md == ORC_EMPTY ? 0.0f : mg->md_mags[md];
}

typedef struct { uint64_t key; uint32_t node; float score; } mitem;
static int cmp_desc(const void *a, const void *b) {
    uint64_t x = ((const mitem *)a)->key, y = ((const mitem *)b)->key;
    return (x < y) - (x > y);
}
static void push(mitem *h, size_t *n, mitem it) {
    size_t i = (*n)++;
    h[i] = it;
    while (i > 0) { size_t p = (i - 1) / 2; if (h[p].key >= h[i].key) break; mitem t = h[p]; h[p] = h[i]; h[i] = t; i = p; }
}
static mitem pop(mitem *h, size_t *n) {
    mitem top = h[0];
    h[0] = h[--(*n)];
    for (size_t i = 0;;) {
        size_t l = 2 * i + 1, r = l + 1, m = i;
        if (l < *n && h[l].key > h[m].key) m = l;
        if (r < *n && h[r].key > h[m].key) m = r;
        if (m == i) break;
        mitem t = h[m]; h[m] = h[i]; h[i] = t; i = m;
    }
    return top;
}

/* traverse_find_nearest (vector_store.rs:1112-1204) with fvec metadata `x`; the fixed set is the caller's (shared between
 * the traversals of one level, vector_store.rs:266-271, 277-291) */
static int traverse_md(const orc_md_graph *mg, uint32_t level, uint32_t entry, const orc_vector_data *x, uint32_t ef,
                       uint32_t shortlist, uint32_t final_len, uint64_t *fs, mitem *out, uint32_t *out_n, uint64_t *evals,
                       uint64_t *pops) {
    const uint32_t nb = mg_nbrs(mg, level);
    const uint32_t *adj = mg->g.adj[level];
    const uint32_t take = shortlist < nb ? shortlist : nb;
    size_t hn = 0, rn = 0;
    mitem *heap = (mitem *)malloc(sizeof(mitem) * ((size_t)ef * take + 2));
    mitem *res = (mitem *)malloc(sizeof(mitem) * ((size_t)ef + 1));
    orc_vector_data y;
    float d;
    node_data(mg, level, entry, &y);
    int rc = orc_distance_md(mg->g.metric, mg->g.storage_type, mg->g.dim, mg->md_dims, x, &y, &d);
    if (evals) (*evals)++;
    if (rc != ORC_OK) goto done;
    fs_insert(fs, nb, y.id);
    push(heap, &hn, (mitem){mkey(mg, d, y.id), entry, d});
    uint32_t visited = 0;
    while (hn > 0) {
        mitem cur = pop(heap, &hn);
        if (visited >= ef) break;
        visited++;
        if (pops) (*pops)++;
        res[rn++] = cur;
        for (uint32_t s = 0; s < take; ++s) {
            const uint32_t nbl = adj[(size_t)cur.node * nb + s];
            if (nbl == ORC_EMPTY) continue;
            node_data(mg, level, nbl, &y);
            if (fs_member(fs, nb, y.id)) continue;
            rc = orc_distance_md(mg->g.metric, mg->g.storage_type, mg->g.dim, mg->md_dims, x, &y, &d);
            if (evals) (*evals)++;
            if (rc != ORC_OK) goto done;
            fs_insert(fs, nb, y.id);
            push(heap, &hn, (mitem){mkey(mg, d, y.id), nbl, d});
        }
    }
    qsort(res, rn, sizeof(mitem), cmp_desc);
    if (rn > final_len) rn = final_len; /* 100 for search, 64 while indexing (vector_store.rs:1194) */
    memcpy(out, res, sizeof(mitem) * rn);
    *out_n = (uint32_t)rn;
done:
    free(heap);
    free(res);
    return rc;
}

int orc_ann_search_md(const orc_md_graph *mg, const void *qcode, float qmag, const int8_t *filters, size_t n_filters,
                      int has_filter, uint32_t ef_search, uint32_t shortlist_size, uint32_t *out_ids, uint32_t *out_rows,
                      float *out_scores, size_t cap, size_t *out_n, uint64_t *evals, uint64_t *pops) {
    const size_t M = mg->md_dims;
    uint32_t entry = has_filter ? mg->pseudo_entry : mg->g.entry; /* hnsw/mod.rs:416-420 */
    uint32_t maxnb = mg->g.level0_neighbors_count > mg->g.neighbors_count ? mg->g.level0_neighbors_count : mg->g.neighbors_count;
    uint64_t *fs = (uint64_t *)malloc(sizeof(uint64_t) * maxnb);
    int32_t *fbits = (int32_t *)malloc(sizeof(int32_t) * (M ? M : 1));
    mitem *z = (mitem *)malloc(sizeof(mitem) * (100 * (n_filters ? n_filters : 1) + 1));
    mitem tmp[100];
    size_t total = 0;
    int rc = ORC_OK;
    orc_vector_data x = {qcode, qmag, 0, 0, NULL, 0.0f};
    for (int level = (int)mg->g.num_levels; level >= 0 && rc == ORC_OK; --level) {
        const uint32_t nb = mg_nbrs(mg, (uint32_t)level);
        size_t zn = 0;
        memset(fs, 0, sizeof(uint64_t) * nb);
        fs_insert(fs, nb, ORC_QUERY_ID);
        if (has_filter) {
            for (size_t f = 0; f < n_filters && rc == ORC_OK; ++f) {
                for (size_t i = 0; i < M; ++i) fbits[i] = filters[f * M + i];
                x.md_bits = fbits;
                x.md_mag = orc_query_filter_mag(filters + f * M, M);
                uint32_t tn = 0;
                rc = traverse_md(mg, (uint32_t)level, entry, &x, ef_search, shortlist_size, 100, fs, tmp, &tn, evals, pops);
                if (rc != ORC_OK) break;
                for (uint32_t i = 0; i < tn; ++i) {
                    if (mg->g.metric == ORC_METRIC_COSINE && tmp[i].score == -1.0f) continue; /* vector_store.rs:294-303 */
                    z[zn++] = tmp[i];
                }
            }
            if (rc != ORC_OK) break;
            qsort(z, zn, sizeof(mitem), cmp_desc);
            if (zn > 100) zn = 100;
        } else {
            x.md_bits = NULL;
            uint32_t tn = 0;
            rc = traverse_md(mg, (uint32_t)level, entry, &x, ef_search, shortlist_size, 100, fs, z, &tn, evals, pops);
            if (rc != ORC_OK) break;
            zn = tn;
        }
        if (zn == 0) { /* vector_store.rs:329-380 */
            orc_vector_data y;
            node_data(mg, (uint32_t)level, entry, &y);
            float best = 0.0f;
            if (has_filter) {
                int have = 0;
                for (size_t f = 0; f < n_filters && rc == ORC_OK; ++f) {
                    for (size_t i = 0; i < M; ++i) fbits[i] = filters[f * M + i];
                    x.md_bits = fbits;
                    x.md_mag = orc_query_filter_mag(filters + f * M, M);
                    float d;
                    rc = orc_distance_md(mg->g.metric, mg->g.storage_type, mg->g.dim, M, &x, &y, &d);
                    if (rc != ORC_OK) break;
                    if (!have || orc_order_key(mg->g.metric, d) > orc_order_key(mg->g.metric, best)) { best = d; have = 1; }
                }
                if (rc == ORC_OK && !have) rc = ORC_UNREACHABLE; /* dists.into_iter().max().unwrap() on an empty vec panics */
            } else {
                x.md_bits = NULL;
                rc = orc_distance_md(mg->g.metric, mg->g.storage_type, mg->g.dim, M, &x, &y, &best);
            }
            if (rc != ORC_OK) break;
            z[0] = (mitem){mkey(mg, best, y.id), entry, best};
            zn = 1;
        }
        for (size_t i = 0; i < zn && total < cap; ++i) {
            orc_vector_data y;
            node_data(mg, (uint32_t)level, z[i].node, &y);
            out_ids[total] = y.id;
            out_rows[total] = orc_replica_kind(&y) == ORC_KIND_PSEUDO ? ORC_EMPTY : mg->g.node_row[level][z[i].node];
            out_scores[total] = z[i].score;
            total++;
        }
        if (level > 0) entry = mg->g.child[level][z[0].node];
    }
    free(fs);
    free(fbits);
    free(z);
    *out_n = total;
    return rc;
}

int orc_hnsw_search_batch_md(const orc_md_graph *mg, const float *raw, const float *queries, size_t nq, float lo, float hi,
                             const uint32_t *filter_offsets, const int8_t *filter_dims, const uint8_t *has_filter,
                             uint32_t ef_search, uint32_t shortlist_size, size_t k, int threads, uint32_t *out_ids,
                             float *out_scores, uint32_t *out_counts, uint8_t *err, uint64_t *evals, uint64_t *pops) {
    if (threads < 1) threads = 1;
    const size_t dim = mg->g.dim, cb = orc_code_bytes(mg->g.storage_type, dim);
    const size_t cap = ((size_t)mg->g.num_levels + 1) * 100;
    uint64_t ev_total = 0, pop_total = 0;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1) reduction(+ : ev_total, pop_total)
    for (long long qi = 0; qi < (long long)nq; ++qi) {
        const float *q = queries + (size_t)qi * dim;
        uint8_t *qcode = (uint8_t *)malloc(cb ? cb : 1);
        float qmag;
        orc_quantize(mg->g.storage_type, lo, hi, q, dim, qcode, &qmag);
        uint32_t *ids = (uint32_t *)malloc(sizeof(uint32_t) * cap), *rows = (uint32_t *)malloc(sizeof(uint32_t) * cap);
        float *sc = (float *)malloc(sizeof(float) * cap);
        size_t n = 0;
        uint64_t ev = 0, pp = 0;
        const int hf = has_filter ? has_filter[qi] : 0;
        const uint32_t f0 = filter_offsets ? filter_offsets[qi] : 0, f1 = filter_offsets ? filter_offsets[qi + 1] : 0;
        int rc = orc_ann_search_md(mg, qcode, qmag, filter_dims ? filter_dims + (size_t)f0 * mg->md_dims : NULL, f1 - f0, hf,
                                   ef_search, shortlist_size, ids, rows, sc, cap, &n, &ev, &pp);
        ev_total += ev;
        pop_total += pp;
        for (size_t j = 0; j < k; ++j) { out_ids[(size_t)qi * k + j] = 0xFFFFFFFFu; out_scores[(size_t)qi * k + j] = 0.0f; }
        if (out_counts) out_counts[qi] = 0;
        if (rc != ORC_OK) {
            if (err) err[qi] = (uint8_t)(rc == ORC_CALCULATION_ERROR ? 1 : rc == ORC_UNREACHABLE ? 4 : 2);
        } else {
            if (err) err[qi] = 0;
            /* remove_duplicates_and_filter (common.rs:381-412): first occurrence per replica id, no root, no pseudo nodes,
             * sort by the traversal similarity, keep 5k */
            mitem *c = (mitem *)malloc(sizeof(mitem) * (n ? n : 1));
            size_t m = 0;
            for (size_t i = 0; i < n; ++i) {
                int seen = 0;
                for (size_t j = 0; j < i; ++j)
                    if (ids[j] == ids[i]) { seen = 1; break; }
                if (seen || ids[i] == 0xFFFFFFFFu || rows[i] == ORC_EMPTY) continue;
                c[m++] = (mitem){mkey(mg, sc[i], ids[i]), (uint32_t)i, sc[i]};
            }
            qsort(c, m, sizeof(mitem), cmp_desc);
            if (m > 5 * k) m = 5 * k;
            /* finalize_ann_results (vector_store.rs:404-445): exact cosine on the base vector of each replica */
            const float mag_q = orc_mag_f32(q, dim);
            for (size_t i = 0; i < m; ++i) {
                const size_t src = c[i].node;
                const float cs = orc_rerank_cosine(q, mag_q, raw + (size_t)rows[src] * dim, dim);
                c[i].key = ((uint64_t)orc_order_key(ORC_METRIC_COSINE, cs) << 32) | (uint64_t)(~ids[src]);
                c[i].score = cs;
                c[i].node = ids[src];
            }
            qsort(c, m, sizeof(mitem), cmp_desc);
            const size_t mk = m < k ? m : k;
            for (size_t i = 0; i < mk; ++i) { out_ids[(size_t)qi * k + i] = c[i].node; out_scores[(size_t)qi * k + i] = c[i].score; }
            if (out_counts) out_counts[qi] = (uint32_t)mk;
            free(c);
        }
        free(qcode);
        free(ids);
        free(rows);
        free(sc);
    }
    if (evals) *evals = ev_total;
    if (pops) *pops = pop_total;
    return ORC_OK;
}

/* ------------------------------------------------------------------ builder with replica nodes */
#define ORC_PSEUDO_ROOT_ID (0xFFFFFFFFu - 257u) /* metadata/mod.rs:219-225 */

struct orc_md_built {
    orc_md_graph mg;
    uint32_t nlevels1;
    blevel *lv;
    uint32_t *cnt_arr;
    const uint32_t **node_row_arr, **adj_arr, **child_arr, **node_id_arr, **node_md_arr;
    uint32_t min_key, max_key;
};

static void md_refresh_view(orc_md_built *b) {
    for (uint32_t L = 0; L < b->nlevels1; ++L) {
        b->cnt_arr[L] = b->lv[L].cnt;
        b->node_row_arr[L] = b->lv[L].node_row;
        b->adj_arr[L] = b->lv[L].adj;
        b->child_arr[L] = b->lv[L].child;
        b->node_id_arr[L] = b->lv[L].node_id;
        b->node_md_arr[L] = b->lv[L].node_md;
    }
}

/* ProbNode::replica_node_kind (prob_node.rs:487-496): VectorData { id: Some(get_id()), metadata } */
static int node_kind(const orc_md_built *b, const blevel *l, uint32_t node) {
    orc_vector_data v;
    memset(&v, 0, sizeof v);
    const uint32_t md = l->node_md[node];
    v.has_id = 1;
    v.id = l->node_id[node];
    v.md_bits = md == ORC_EMPTY ? NULL : b->mg.md_bits + (size_t)md * b->mg.md_dims;
    v.md_mag = md == ORC_EMPTY ? 0.0f : b->mg.md_mags[md];
    return orc_replica_kind(&v);
}

/* create_node_edges (src/vector_store.rs:976-1070) with the replica rules of :1014-1040 */
static void md_create_node_edges(orc_md_built *b, blevel *l, uint32_t node, const mitem *z, uint32_t zn) {
    const int cosine = b->mg.g.metric == ORC_METRIC_COSINE; /* the rules match MetricResult::CosineSimilarity only */
    const int nk = node_kind(b, l, node);
    uint32_t successful = 0;
    for (uint32_t i = 0; i < zn; ++i) {
        if (successful >= l->nb) break;
        const uint32_t nbr = z[i].node;
        if (cosine) {
            const int bk = node_kind(b, l, nbr);
            if (bk == ORC_KIND_PSEUDO && nk == ORC_KIND_METADATA && z[i].score != 1.0f) continue;
            if (bk == ORC_KIND_METADATA && nk == ORC_KIND_METADATA && z[i].score == -1.0f) continue;
        }
        const uint32_t dkey = orc_order_key(b->mg.g.metric, z[i].score);
        const int idx = bl_add_neighbor(b->min_key, b->max_key, l, node, nbr, dkey);
        if (idx >= 0) {
            const int j = bl_add_neighbor(b->min_key, b->max_key, l, nbr, node, dkey);
            if (j >= 0) successful++;
            else if (l->adj[(size_t)node * l->nb + idx] == nbr) l->adj[(size_t)node * l->nb + idx] = ORC_EMPTY; /* remove_neighbor_by_index_and_id */
        }
    }
}

orc_md_built *orc_hnsw_build_md(int metric, int st, size_t dim, const void *codes, const float *mags, size_t md_dims,
                                const int32_t *md_bits, const float *md_mags, const orc_replica_list *rl, uint32_t num_levels,
                                uint32_t nbrs, uint32_t nbrs0, uint32_t ef_construction, uint32_t shortlist, uint8_t *failed_out) {
    orc_md_built *b = (orc_md_built *)calloc(1, sizeof(orc_md_built));
    const uint32_t L1 = num_levels + 1;
    b->nlevels1 = L1;
    b->lv = (blevel *)calloc(L1, sizeof(blevel));
    b->cnt_arr = (uint32_t *)calloc(L1, sizeof(uint32_t));
    b->node_row_arr = (const uint32_t **)calloc(L1, sizeof(uint32_t *));
    b->adj_arr = (const uint32_t **)calloc(L1, sizeof(uint32_t *));
    b->child_arr = (const uint32_t **)calloc(L1, sizeof(uint32_t *));
    b->node_id_arr = (const uint32_t **)calloc(L1, sizeof(uint32_t *));
    b->node_md_arr = (const uint32_t **)calloc(L1, sizeof(uint32_t *));
    orc_md_graph *mg = &b->mg;
    orc_graph *g = &mg->g;
    g->num_levels = num_levels; g->neighbors_count = nbrs; g->level0_neighbors_count = nbrs0; g->n = rl->main_root_row;
    g->metric = metric; g->storage_type = st; g->dim = dim; g->codes = codes; g->mags = mags;
    g->cnt = b->cnt_arr; g->node_row = b->node_row_arr; g->adj = b->adj_arr; g->child = b->child_arr;
    mg->md_dims = md_dims; mg->md_bits = md_bits; mg->md_mags = md_mags;
    mg->node_id = b->node_id_arr; mg->node_md = b->node_md_arr;
    bl_min_max_keys(metric, &b->min_key, &b->max_key);
    /* both roots exist on every level, linked downwards (vector_store.rs:96-140, 203-250) */
    for (uint32_t L = 0; L < L1; ++L) {
        blevel *l = &b->lv[L];
        l->nb = L == 0 ? nbrs0 : nbrs;
        lv_reserve(l, 64);
        l->cnt = 2;
        lv_init_node(b->min_key, l, 0, rl->main_root_row);
        l->node_id[0] = ORC_ROOT_ID; l->node_md[0] = rl->main_root_md;
        lv_init_node(b->min_key, l, 1, rl->pseudo_root_row);
        l->node_id[1] = ORC_PSEUDO_ROOT_ID; l->node_md[1] = rl->pseudo_root_md;
        if (L > 0) { l->child[0] = 0; l->child[1] = 1; }
    }
    g->entry = 0;
    mg->pseudo_entry = 1;
    const uint32_t maxnb = nbrs0 > nbrs ? nbrs0 : nbrs;
    uint64_t *fs = (uint64_t *)malloc(sizeof(uint64_t) * maxnb);
    mitem *z = (mitem *)malloc(sizeof(mitem) * 64 * L1);
    uint32_t *zn = (uint32_t *)malloc(sizeof(uint32_t) * L1);
    uint32_t *node_at = (uint32_t *)malloc(sizeof(uint32_t) * L1);
    const size_t cb = orc_code_bytes(st, dim);
    for (uint32_t t = 0; t < rl->n_nodes; ++t) {
        if (failed_out) failed_out[t] = 0;
        const uint32_t row = rl->row[t], md = rl->md_row[t];
        orc_vector_data x;
        x.code = (const uint8_t *)codes + (size_t)row * cb;
        x.mag = mags[row];
        x.has_id = 1;
        x.id = rl->base_id[t];
        x.md_bits = md == ORC_EMPTY ? NULL : md_bits + (size_t)md * md_dims;
        x.md_mag = md == ORC_EMPTY ? 0.0f : md_mags[md];
        /* IndexableEmbedding::node_kind / root_node_kind (vector_store.rs:461-483): metadata with mag != 0 -> pseudo root */
        const int under_pseudo = x.md_bits && x.md_mag != 0.0f;
        const uint32_t max_level = rl->max_level[t];
        uint32_t entry = under_pseudo ? mg->pseudo_entry : g->entry, parent = ORC_EMPTY;
        int failed = 0;
        for (int level = (int)num_levels; level >= 0; --level) {
            blevel *l = &b->lv[level];
            md_refresh_view(b);
            memset(fs, 0, sizeof(uint64_t) * l->nb);
            fs_insert(fs, l->nb, rl->node_id[t]); /* skipm.insert(new_node_id), :803-807 */
            mitem *zl = z + 64 * level;
            uint32_t cnt = 0;
            int rc = traverse_md(mg, (uint32_t)level, entry, &x, ef_construction, shortlist, 64, fs, zl, &cnt, NULL, NULL);
            if (rc == ORC_OK && cnt == 0) { /* :829-851: the entry itself, fvec_data without an id */
                orc_vector_data x0 = x, y;
                x0.has_id = 0;
                node_data(mg, (uint32_t)level, entry, &y);
                float d = 0.0f;
                rc = orc_distance_md(metric, st, dim, md_dims, &x0, &y, &d);
                zl[0] = (mitem){mkey(mg, d, y.id), entry, d};
                cnt = 1;
            }
            if (rc != ORC_OK) { failed = 1; break; } /* the reference returns Err (or panics on an unreachable arm): not indexed */
            zn[level] = cnt;
            const uint32_t next_entry = level > 0 ? l->child[zl[0].node] : 0;
            if ((uint32_t)level <= max_level) {
                lv_reserve(l, l->cnt + 1);
                const uint32_t idx = l->cnt++;
                lv_init_node(b->min_key, l, idx, row);
                l->node_id[idx] = rl->node_id[t];
                l->node_md[idx] = md;
                if (parent != ORC_EMPTY) b->lv[level + 1].child[parent] = idx;
                node_at[level] = idx;
                parent = idx;
            }
            entry = next_entry;
        }
        if (failed) { /* nodes created on the way down stay unlinked, like a failed reference insert */
            if (failed_out) failed_out[t] = 1;
            continue;
        }
        const uint32_t top = max_level < num_levels ? max_level : num_levels;
        for (uint32_t level = 0; level <= top; ++level)
            md_create_node_edges(b, &b->lv[level], node_at[level], z + 64 * level, zn[level]);
    }
    md_refresh_view(b);
    free(fs); free(z); free(zn); free(node_at);
    return b;
}

const orc_md_graph *orc_md_built_graph(const orc_md_built *b) { return &b->mg; }

void orc_md_built_free(orc_md_built *b) {
    if (!b) return;
    for (uint32_t L = 0; L < b->nlevels1; ++L) lv_free(&b->lv[L]);
    free(b->lv); free(b->cnt_arr); free(b->node_row_arr); free(b->adj_arr); free(b->child_arr);
    free(b->node_id_arr); free(b->node_md_arr);
    free(b);
}
