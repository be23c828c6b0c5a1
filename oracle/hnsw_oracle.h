/*
 * hnsw_oracle.h -- CPU restatement of cosdata's HNSW search (and a deterministic
 * single-threaded restatement of its build) on a FLAT graph.
 *
 * TEST INFRASTRUCTURE ONLY (see cosdata_oracle.h).
 *
 * Reference:
 *   ann_search              src/vector_store.rs:256-402
 *   traverse_find_nearest   src/vector_store.rs:1112-1204
 *   PerformantFixedSet      src/models/fixedset.rs:2-29
 *   remove_duplicates_and_filter  src/models/common.rs:381-412
 *   finalize_ann_results    src/vector_store.rs:404-445
 *   index_embedding / create_node_edges  src/vector_store.rs:782-1109
 *   ProbNode::add_neighbor  src/models/prob_node.rs:210-283
 *
 * The reference graph is a pointer structure built concurrently with random
 * levels, so parity is defined on an exported graph:
 * GPU search == this search on the same arrays.  Where the reference leaves
 * the order unspecified (BinaryHeap of (MetricResult, pointer), sort_unstable)
 * the oracle rule is: better score first, then smaller node id.
 *
 * Flat graph (the same arrays cdb_index_set_graph takes):
 *   vectors: rows 0..n-1 are data (id = row), row n is the root vector (id u32::MAX,
 *            vector_store.rs:57-67); codes/mags have n+1 rows.
 *   level L in 0..=num_levels has cnt[L] nodes with level-local indices:
 *     node_row[L][i]  vector row of node i           (level 0: cnt = n+1, node_row = identity)
 *     adj[L][i*nbrs(L) + s]  local index of the neighbour in slot s, 0xFFFFFFFF = empty slot
 *     child[L][i]     local index at level L-1 of the same vector (L >= 1)
 *   entry = local index of the root at level num_levels.
 */
#ifndef HNSW_ORACLE_H
#define HNSW_ORACLE_H
#include "cosdata_oracle.h"

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_EMPTY 0xFFFFFFFFu
#define ORC_ROOT_ID 0xFFFFFFFFu
#define ORC_QUERY_ID 0xFFFFFFFEu /* hnsw/mod.rs:398: u32::MAX - 1 */

typedef struct {
    uint32_t num_levels;      /* hnsw_params.num_layers: levels 0..=num_levels */
    uint32_t neighbors_count; /* slots per node at levels >= 1 */
    uint32_t level0_neighbors_count;
    uint32_t n;               /* data rows; root vector is row n */
    uint32_t entry;           /* root's local index at the top level */
    const uint32_t *cnt;      /* [num_levels+1] */
    const uint32_t *const *node_row;
    const uint32_t *const *adj;
    const uint32_t *const *child; /* child[0] unused (may be NULL) */
    int metric, storage_type;
    size_t dim;
    const void *codes;        /* [n+1] rows, orc_code_bytes each */
    const float *mags;        /* [n+1] */
} orc_graph;

typedef struct {
    uint32_t ef;              /* ef_search / ef_construction */
    uint32_t shortlist_size;  /* config.search.shortlist_size */
    uint32_t final_len;       /* 100 for search, 64 while indexing (vector_store.rs:1194) */
    uint32_t self_id;         /* id pre-inserted into the fixed set: ORC_QUERY_ID (search) / new node id (build) */
} orc_traverse_params;

/* traverse_find_nearest on one level.  out_* capacity >= final_len.  Returns status;
 * *out_n results sorted best-first.  evals/pops accumulate (may be NULL). */
int orc_traverse(const orc_graph *g, uint32_t level, uint32_t entry_local, const void *qcode, float qmag,
                 const orc_traverse_params *p, uint64_t *fixedset /* nbrs(level) words, zeroed by caller */,
                 uint32_t *out_nodes, float *out_scores, uint32_t *out_n, uint64_t *evals, uint64_t *pops);

/* ann_search: all levels; results concatenated top level first as (vector row, score).
 * cap >= (num_levels+1)*100. */
int orc_ann_search(const orc_graph *g, const void *qcode, float qmag, uint32_t ef_search, uint32_t shortlist_size,
                   uint32_t *out_rows, float *out_scores, size_t cap, size_t *out_n, uint64_t *evals, uint64_t *pops);

/* remove_duplicates_and_filter: dedup by id keeping the first occurrence, drop the root,
 * sort best-first, truncate to 5*k.  In place; returns the new length. */
size_t orc_dedup_filter(const orc_graph *g, uint32_t *rows, float *scores, size_t n, size_t k);

/* full search_internal for a batch: quantized ann_search + dedup + exact f32 re-rank (raw rows [n][dim]).
 * out_ids/out_scores [nq][k] (missing = 0xFFFFFFFF/0), out_counts[nq], err[nq]. */
int orc_hnsw_search_batch(const orc_graph *g, const float *raw, const float *queries, size_t nq, float lo, float hi,
                          uint32_t ef_search, uint32_t shortlist_size, size_t k, int threads,
                          uint32_t *out_ids, float *out_scores, uint32_t *out_counts, uint8_t *err,
                          uint64_t *evals, uint64_t *pops);

/* deterministic builder */
typedef struct orc_built orc_built;
orc_built *orc_hnsw_build(int metric, int storage_type, size_t dim, const void *codes /* n+1 rows */,
                          const float *mags, uint32_t n, uint32_t num_levels, uint32_t neighbors_count,
                          uint32_t level0_neighbors_count, uint32_t ef_construction, uint32_t shortlist_size,
                          uint64_t seed);
const orc_graph *orc_built_graph(const orc_built *b);
void orc_built_free(orc_built *b);

#ifdef __cplusplus
}
#endif
#endif
