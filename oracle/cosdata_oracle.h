/*
 * cosdata_oracle.h -- CPU restatement of the cosdata ANN distance hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or
 * the timed CPU baseline.  The product path (cosdata_b200/, include/) never
 * links, imports or calls it.
 *
 * Parity status: the Rust reference cannot be compiled in this image (no
 * cargo/rustc), so this restatement is pinned against
 *   (a) the properties the reference's own tests assert
 *       (src/models/dot_product/x86_64.rs:454-505, 544-602, 673-746, 784-816;
 *        src/models/types.rs:1610-1633), re-expressed with fixed seeds in
 *       tests/test_oracle_*.py, and
 *   (b) independent numpy restatements of the same formulas (tests/).
 * The reference holds no golden vectors for this path ("parity unpinned by
 * golden data").
 *
 * Every function cites the reference file:line (relative to /root/reference)
 * that it restates.
 */
#ifndef COSDATA_ORACLE_H
#define COSDATA_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Mirrors DistanceError (src/distance/mod.rs:18-22) + extras. */
enum {
    ORC_OK = 0,
    ORC_STORAGE_MISMATCH = 1,
    ORC_CALCULATION_ERROR = 2,
    ORC_INVALID = 3,
    ORC_UNIMPLEMENTED = 6 /* reference hits unimplemented!() */
};

/* StorageType (src/quantization/mod.rs:19-25); SubByte(r) -> 1..3. */
enum {
    ORC_ST_U8 = 0,
    ORC_ST_SUB1 = 1,
    ORC_ST_SUB2 = 2,
    ORC_ST_SUB3 = 3,
    ORC_ST_F16 = 4,
    ORC_ST_F32 = 5,
    ORC_ST_BF16 = 6   /* labelled extension (bfloat16), see cosdata_oracle.c */
};

/* DistanceMetric (src/models/types.rs:460-467). */
enum {
    ORC_METRIC_COSINE = 0,
    ORC_METRIC_EUCLIDEAN = 1,
    ORC_METRIC_HAMMING = 2,
    ORC_METRIC_DOT = 3
};

/* ---- synthetic data generator shared with the CUDA side (include/cosdata_b200.h) */
float orc_synth_value(uint64_t seed, uint64_t idx);
void orc_synth_fill(uint64_t seed, uint64_t first_idx, size_t n, float *out);

/* ---- half conversions (crate `half` 2.4.1: IEEE binary16, RNE) */
uint16_t orc_f32_to_f16(float x);
float orc_f16_to_f32(uint16_t h);

/* ---- dot products: src/models/dot_product.rs + dot_product/x86_64.rs */
uint64_t orc_dot_u8_scalar(const uint8_t *a, const uint8_t *b, size_t n);
uint64_t orc_dot_u8_avx2(const uint8_t *a, const uint8_t *b, size_t n);
float orc_dot_f16(const uint16_t *a, const uint16_t *b, size_t n);
float orc_dot_f32_scalar(const float *a, const float *b, size_t n);
float orc_dot_f32_simd(const float *a, const float *b, size_t n);
/* sub-byte: planes laid out [r][nbytes] contiguous, plane p at x + p*nbytes */
float orc_dot_binary_scalar(const uint8_t *x, const uint8_t *y, size_t nbytes);
float orc_dot_binary_avx2(const uint8_t *x, const uint8_t *y, size_t nbytes);
float orc_dot_quaternary_scalar(const uint8_t *x, const uint8_t *y, size_t nbytes);
float orc_dot_quaternary_avx2(const uint8_t *x, const uint8_t *y, size_t nbytes);
float orc_dot_octal_scalar(const uint8_t *x, const uint8_t *y, size_t nbytes);
float orc_dot_octal_avx2(const uint8_t *x, const uint8_t *y, size_t nbytes);
uint64_t orc_count_ones_256(const uint8_t *p32);

/* ---- quantization: src/quantization/scalar.rs:10-52, src/models/common.rs:225-275 */
size_t orc_code_bytes(int storage_type, size_t dim);
int orc_quantize(int storage_type, float lo, float hi, const float *v, size_t dim,
                 void *out_code, float *out_mag);

/* ---- value-range sampling for `quantization: auto`: HNSWIndex::sample_embedding + finalize_sampling
 *      (src/indexes/hnsw/mod.rs:202-351).  counts[0..7) = # > {0.025,0.05,0.1,0.2,0.3,0.4,0.5}, counts[7..14) = # < -{..};
 *      counts are ADDED to (callers zero them first). */
void orc_sample_counts(const float *values, size_t n_values, uint64_t *counts14);
void orc_values_range(const uint64_t *counts14, uint64_t n_values, float clamp_margin_percent, float *range2);

/* ---- pairwise DistanceFunction::calculate for the (Base,Base) arm */
int orc_distance(int metric, int storage_type, size_t dim,
                 const void *x_code, float x_mag,
                 const void *y_code, float y_mag, float *out);

/* ---- ordering: MetricResult::cmp (src/models/types.rs:401-411) as a u32 key,
 *      larger key == "greater" == better. */
uint32_t orc_order_key(int metric, float value);

/* ---- exact re-rank formula of finalize_ann_results (src/vector_store.rs:414-439) */
float orc_mag_f32(const float *v, size_t dim);
float orc_rerank_cosine(const float *q, float mag_q, const float *v, size_t dim);

/* ---- brute-force scans (reference has none in Rust; this applies the
 *      reference's per-pair formula to every row).
 *   mode 0: finalize_ann_results formula over raw f32 rows (configs C1/C2)
 *   mode 1: DistanceMetric::calculate over quantized codes (config C4)
 * Tie rule (oracle-defined; reference is unspecified): better score first,
 * then smaller id.  out_ids/out_scores are [nq][k]; missing slots get
 * id 0xFFFFFFFF / score 0.  err_flags[q] != 0 mirrors the Err propagation. */
int orc_brute_topk_f32(const float *corpus, size_t n, size_t dim,
                       const float *queries, size_t nq, size_t k, int threads,
                       uint32_t *out_ids, float *out_scores);
int orc_brute_topk_codes(int metric, int storage_type, size_t dim,
                         const void *codes, const float *mags, size_t n,
                         const void *qcodes, const float *qmags, size_t nq,
                         size_t k, int threads,
                         uint32_t *out_ids, float *out_scores, uint8_t *err_flags);

/* S3: re-rank candidate ids with the finalize formula, sort desc, truncate k. */
int orc_rerank_f32(const float *corpus, size_t dim, const float *q,
                   const uint32_t *cand, size_t ncand, size_t k,
                   uint32_t *out_ids, float *out_scores);

/* HNSW search / build restatement lives in hnsw_oracle.h */

#ifdef __cplusplus
}
#endif
#endif
