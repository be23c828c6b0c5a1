/*
 * metadata_oracle.h -- CPU restatement of the metadata-filter arms of the hot path.  TEST INFRASTRUCTURE ONLY
 * (see cosdata_oracle.h).
 *
 * Reference:
 *   VectorData::replica_node_kind             src/models/types.rs:223-243
 *   Metadata::from(MetadataDimensions / &QueryFilterDimensions)   src/models/types.rs:111-147
 *   CosineSimilarity::calculate (replica arms) src/distance/cosine.rs:34-102
 *   cosine_similarity_mdims                    src/distance/cosine.rs:243-259
 *   ann_search, filter branch + empty fallback src/vector_store.rs:273-313, 329-380
 *   search_internal root selection             src/indexes/hnsw/mod.rs:411-420
 *   remove_duplicates_and_filter (pseudo drop) src/models/common.rs:381-412
 *   get_raw_emb_by_internal_id (replica -> base) src/models/collection.rs:368-384  (here: node_row of the flat graph)
 * Only the cosine metric looks at metadata; the other metrics score the quantized vectors alone
 * (dotproduct.rs:12-65, euclidean.rs:9-40, hamming.rs:10-58).
 */
#ifndef METADATA_ORACLE_H
#define METADATA_ORACLE_H
#include "hnsw_oracle.h"

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_UNREACHABLE 7 /* an arm the reference marks unreachable!() (it would panic) */
#define ORC_KIND_PSEUDO 0
#define ORC_KIND_BASE 1
#define ORC_KIND_METADATA 2

typedef struct {
    const void *code;
    float mag;
    int has_id;              /* VectorData.id: Option<&InternalId> */
    uint32_t id;
    const int32_t *md_bits;  /* Metadata.mbits, NULL = metadata: None */
    float md_mag;            /* Metadata.mag */
} orc_vector_data;

float orc_metadata_mag(const int32_t *dims, size_t m);        /* types.rs:111-125 */
float orc_query_filter_mag(const int8_t *dims, size_t m);     /* types.rs:127-146 */
int orc_replica_kind(const orc_vector_data *v);
/* DistanceMetric::calculate(x, y) with metadata; x is the query side (fvec_data), y the node side */
int orc_distance_md(int metric, int storage_type, size_t dim, size_t md_dims, const orc_vector_data *x,
                    const orc_vector_data *y, float *out);

/* flat graph with replica nodes: node ids and metadata per node */
typedef struct {
    orc_graph g;                      /* g.n is unused for ids here; codes/mags rows are addressed through node_row */
    size_t md_dims;
    const int32_t *md_bits;           /* [n_md][md_dims] */
    const float *md_mags;             /* [n_md] */
    const uint32_t *const *node_id;   /* [levels+1][cnt]: ProbNode::get_id() (replica id; prop_value.id without metadata) */
    const uint32_t *const *node_md;   /* [levels+1][cnt]: row of the metadata table, ORC_EMPTY = prop_metadata None */
    uint32_t pseudo_entry;            /* top-level local index of the pseudo root (get_pseudo_root_vec) */
} orc_md_graph;

/* ann_search with Option<&Vec<QueryFilterDimensions>>: n_filters == 0 and filters == NULL -> None (main root),
 * else one traversal per filter from the pseudo root.  Results concatenated top level first as level-0-independent
 * triples (node id, vector row, score); pseudo nodes are reported with row = ORC_EMPTY.  cap >= (levels+1)*100. */
int orc_ann_search_md(const orc_md_graph *mg, const void *qcode, float qmag, const int8_t *filters, size_t n_filters,
                      int has_filter, uint32_t ef_search, uint32_t shortlist_size, uint32_t *out_ids, uint32_t *out_rows,
                      float *out_scores, size_t cap, size_t *out_n, uint64_t *evals, uint64_t *pops);

/* search_internal for a batch with per-query filters: filter_offsets[nq+1] index rows of filter_dims ([total][md_dims]);
 * has_filter[q] distinguishes Some(empty) from None.  out_ids are replica ids. */
int orc_hnsw_search_batch_md(const orc_md_graph *mg, const float *raw, const float *queries, size_t nq, float lo, float hi,
                             const uint32_t *filter_offsets, const int8_t *filter_dims, const uint8_t *has_filter,
                             uint32_t ef_search, uint32_t shortlist_size, size_t k, int threads, uint32_t *out_ids,
                             float *out_scores, uint32_t *out_counts, uint8_t *err, uint64_t *evals, uint64_t *pops);

/* ------------------------------------------------------------------ builder for collections with a metadata schema
 * index_embeddings (src/vector_store.rs:714-780) over the flattened IndexableEmbeddings of preprocess_embedding
 * (:629-712), single-threaded, in list order.  Each entry is one graph node to create:
 *   row        vector row (pseudo replicas share the pseudo root's row, :661)
 *   node_id    ProbNode::get_id(): the replica id, or prop_value.id without metadata (:803-806)
 *   base_id    prop_value.id, the id fvec_data carries during the traversal (:812, 1131-1135)
 *   md_row     metadata table row or ORC_EMPTY
 *   max_level  the get_max_insert_level draw (levels_prob, or pseudo_level_probs for pseudo replicas; :749-753) -- drawn by
 *              the caller, the reference uses rand::random
 * Nodes whose metadata has mag != 0 (Pseudo and Metadata kinds) are indexed under the pseudo root, the others under the
 * main root (IndexableEmbedding::root_node_kind, :480-483; types.rs:187-193).  create_node_edges applies the replica rules
 * (:1014-1040): a Metadata node links to a Pseudo neighbour only on cs == 1.0, two Metadata nodes never on cs == -1.0.
 * Layout of every level: [0] main root, [1] pseudo root, then the created nodes in list order. */
typedef struct {
    uint32_t n_nodes;
    const uint32_t *row, *node_id, *base_id, *md_row;
    const uint8_t *max_level;
    uint32_t main_root_row, main_root_md;     /* create_root_node: id u32::MAX, base dimensions with mag 0 (vector_store.rs:79-93) */
    uint32_t pseudo_root_row, pseudo_root_md; /* create_pseudo_root_node: id u32::MAX - 257 (vector_store.rs:160-200) */
} orc_replica_list;

typedef struct orc_md_built orc_md_built;
orc_md_built *orc_hnsw_build_md(int metric, int storage_type, size_t dim, const void *codes, const float *mags, size_t md_dims,
                                const int32_t *md_bits, const float *md_mags, const orc_replica_list *rl, uint32_t num_levels,
                                uint32_t neighbors_count, uint32_t level0_neighbors_count, uint32_t ef_construction,
                                uint32_t shortlist_size, uint8_t *failed /* [n_nodes] or NULL: insert returned Err */);
const orc_md_graph *orc_md_built_graph(const orc_md_built *b);
void orc_md_built_free(orc_md_built *b);

#ifdef __cplusplus
}
#endif
#endif
