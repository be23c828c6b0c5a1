/*
 * hnsw_build_internal.h -- per-level node arrays and ProbNode::add_neighbor shared by the two oracle builders
 * (hnsw_oracle.c: collections without a metadata schema; metadata_oracle.c: replica nodes).  TEST INFRASTRUCTURE ONLY.
 */
#ifndef HNSW_BUILD_INTERNAL_H
#define HNSW_BUILD_INTERNAL_H
#include <stdlib.h>

#include "hnsw_oracle.h"

typedef struct {
    uint32_t cnt, cap, nb;
    uint32_t *node_row;
    uint32_t *adj;       /* [cap*nb] local indices */
    uint32_t *simkey;    /* [cap*nb] order key of the slot's similarity (MetricResult ordering) */
    uint32_t *child;
    uint32_t *lowest_idx;
    uint32_t *lowest_key;
    uint32_t *node_id;   /* replica builder only: ProbNode::get_id() */
    uint32_t *node_md;   /* replica builder only: metadata table row or ORC_EMPTY */
} blevel;

/* MetricResult::min / ::max as order keys (src/models/types.rs:435-457) */
static inline void bl_min_max_keys(int metric, uint32_t *min_key, uint32_t *max_key) {
    switch (metric) {
    case ORC_METRIC_COSINE: *min_key = orc_order_key(metric, -1.0f); *max_key = orc_order_key(metric, 2.0f); break;
    default: *min_key = orc_order_key(metric, -__builtin_inff()); *max_key = orc_order_key(metric, __builtin_inff()); break;
    }
}

static inline void lv_reserve(blevel *l, uint32_t want) {
    if (want <= l->cap) return;
    uint32_t nc = l->cap ? l->cap * 2 : 64;
    if (nc < want) nc = want;
    l->node_row = (uint32_t *)realloc(l->node_row, sizeof(uint32_t) * nc);
    l->adj = (uint32_t *)realloc(l->adj, sizeof(uint32_t) * (size_t)nc * l->nb);
    l->simkey = (uint32_t *)realloc(l->simkey, sizeof(uint32_t) * (size_t)nc * l->nb);
    l->child = (uint32_t *)realloc(l->child, sizeof(uint32_t) * nc);
    l->lowest_idx = (uint32_t *)realloc(l->lowest_idx, sizeof(uint32_t) * nc);
    l->lowest_key = (uint32_t *)realloc(l->lowest_key, sizeof(uint32_t) * nc);
    l->node_id = (uint32_t *)realloc(l->node_id, sizeof(uint32_t) * nc);
    l->node_md = (uint32_t *)realloc(l->node_md, sizeof(uint32_t) * nc);
    l->cap = nc;
}
static inline void lv_init_node(uint32_t min_key, blevel *l, uint32_t idx, uint32_t row) {
    l->node_row[idx] = row;
    for (uint32_t s = 0; s < l->nb; ++s) { l->adj[(size_t)idx * l->nb + s] = ORC_EMPTY; l->simkey[(size_t)idx * l->nb + s] = 0; }
    l->child[idx] = ORC_EMPTY;
    l->lowest_idx[idx] = 0;             /* ProbNode::new: lowest_index = (0, MetricResult::min) */
    l->lowest_key[idx] = min_key;
    l->node_id[idx] = ORC_EMPTY;
    l->node_md[idx] = ORC_EMPTY;
}
static inline void lv_free(blevel *l) {
    free(l->node_row); free(l->adj); free(l->simkey); free(l->child); free(l->lowest_idx); free(l->lowest_key);
    free(l->node_id); free(l->node_md);
}

/* ProbNode::add_neighbor (src/models/prob_node.rs:210-283); returns slot index or -1 */
static inline int bl_add_neighbor(uint32_t min_key, uint32_t max_key, blevel *l, uint32_t node, uint32_t nbr, uint32_t dkey) {
    const uint32_t lidx = l->lowest_idx[node], lkey = l->lowest_key[node];
    if (dkey <= lkey) return -1;
    uint32_t *slot = &l->adj[(size_t)node * l->nb + lidx];
    uint32_t *skey = &l->simkey[(size_t)node * l->nb + lidx];
    int ok = 0;
    uint32_t old = ORC_EMPTY;
    if (*slot == ORC_EMPTY) { *slot = nbr; *skey = dkey; ok = 1; }
    else if (dkey > *skey) { old = *slot; *slot = nbr; *skey = dkey; ok = 1; }
    /* recompute (lowest_idx, lowest_sim) */
    uint32_t nidx = 0, nkey = max_key;
    for (uint32_t s = 0; s < l->nb; ++s) {
        if (l->adj[(size_t)node * l->nb + s] == ORC_EMPTY) { nkey = min_key; nidx = s; break; }
        uint32_t k = l->simkey[(size_t)node * l->nb + s];
        if (k < nkey) { nkey = k; nidx = s; }
    }
    l->lowest_idx[node] = nidx;
    l->lowest_key[node] = nkey;
    if (!ok) return -1;
    if (old != ORC_EMPTY) { /* evicted neighbour drops its back link (remove_neighbor_by_id) */
        for (uint32_t s = 0; s < l->nb; ++s)
            if (l->adj[(size_t)old * l->nb + s] == node) { l->adj[(size_t)old * l->nb + s] = ORC_EMPTY; break; }
    }
    return (int)lidx;
}
#endif
