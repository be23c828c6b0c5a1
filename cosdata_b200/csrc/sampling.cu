// sampling.cu -- value-range sampling for `quantization: auto`
// (HNSWIndex::sample_embedding + finalize_sampling, src/indexes/hnsw/mod.rs:202-351).
//
// The reference bumps 14 atomic counters per element (7 thresholds each side of zero) over the first
// `sample_threshold` embeddings, then picks the tightest range whose clipped share is <= clamp_margin_percent.
// Here: one streaming pass over the f32 matrix in HBM (4 B/element read once, 14 compares per element, counters kept
// in registers, one warp reduction + 14 global atomics per CTA).  HBM-bound; algorithmic bytes = 4 per element.
#include <algorithm>

#include "kernels.h"

namespace cdb {

constexpr int SAMPLE_THREADS = 256;

__device__ __forceinline__ void sample_bump(float v, uint32_t (&c)[CDB_SAMPLE_COUNTERS]) {
    // thresholds are the f32 literals of the reference (value: f32 > 0.025 ...)
    c[0] += v > 0.025f; c[1] += v > 0.05f; c[2] += v > 0.1f; c[3] += v > 0.2f; c[4] += v > 0.3f; c[5] += v > 0.4f; c[6] += v > 0.5f;
    c[7] += v < -0.025f; c[8] += v < -0.05f; c[9] += v < -0.1f; c[10] += v < -0.2f; c[11] += v < -0.3f; c[12] += v < -0.4f; c[13] += v < -0.5f;
}

__global__ void __launch_bounds__(SAMPLE_THREADS) sample_counts_kernel(const float *__restrict__ v, uint64_t n_values,
                                                                         unsigned long long *__restrict__ counts) {
    uint32_t c[CDB_SAMPLE_COUNTERS];
#pragma unroll
    for (int i = 0; i < CDB_SAMPLE_COUNTERS; ++i) c[i] = 0;
    const uint64_t tid = (uint64_t)blockIdx.x * SAMPLE_THREADS + threadIdx.x;
    const uint64_t nthreads = (uint64_t)gridDim.x * SAMPLE_THREADS;
    // head elements up to the first 16-byte boundary, then float4 body, then tail
    const uint64_t mis = ((16 - (reinterpret_cast<uintptr_t>(v) & 15)) & 15) / 4;
    const uint64_t head = mis < n_values ? mis : n_values;
    const uint64_t n4 = (n_values - head) / 4;
    const float4 *v4 = reinterpret_cast<const float4 *>(v + head);
    // a thread never sees more than 2^32 elements: n_values / nthreads stays far below that for any HBM-resident matrix
    for (uint64_t i = tid; i < n4; i += nthreads) {
        const float4 x = __ldcs(v4 + i);
        sample_bump(x.x, c); sample_bump(x.y, c); sample_bump(x.z, c); sample_bump(x.w, c);
    }
    if (tid < head) sample_bump(v[tid], c);
    const uint64_t tail0 = head + n4 * 4;
    if (tail0 + tid < n_values) sample_bump(v[tail0 + tid], c);

    __shared__ unsigned long long s_counts[CDB_SAMPLE_COUNTERS];
    if (threadIdx.x < CDB_SAMPLE_COUNTERS) s_counts[threadIdx.x] = 0ull;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < CDB_SAMPLE_COUNTERS; ++i) {
        uint32_t x = c[i];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, off);
        if ((threadIdx.x & 31) == 0 && x) atomicAdd(&s_counts[i], (unsigned long long)x);
    }
    __syncthreads();
    if (threadIdx.x < CDB_SAMPLE_COUNTERS && s_counts[threadIdx.x]) atomicAdd(&counts[threadIdx.x], s_counts[threadIdx.x]);
}

cdb_status sample_counts_device(const float *d_vecs, uint64_t n_values, unsigned long long *d_counts, int sm_count, cudaStream_t s) {
    if (n_values == 0) return CDB_OK;
    const uint64_t want = (n_values / 4 + SAMPLE_THREADS - 1) / SAMPLE_THREADS + 1;
    const uint32_t grid = (uint32_t)std::min<uint64_t>(want, (uint64_t)sm_count * 8);   // 8 resident CTAs of 256 threads per SM
    sample_counts_kernel<<<grid, SAMPLE_THREADS, 0, s>>>(d_vecs, n_values, d_counts);
    CDB_LAUNCH_CHECK();
    return CDB_OK;
}

// finalize_sampling (src/indexes/hnsw/mod.rs:268-342): percent = (count as f32 / values_count) * 100.0 in f32, the first
// threshold (tightest first) whose clipped share is <= clamp_margin_percent wins, else the full [-1, 1].
void values_range_from_counts(const uint64_t *counts, uint64_t n_values, float clamp_margin_percent, float *range) {
    static const float T[7] = {0.025f, 0.05f, 0.1f, 0.2f, 0.3f, 0.4f, 0.5f};
    const float values_count = (float)n_values;
    float hi = 1.0f, lo = -1.0f;
    for (int i = 0; i < 7; ++i)
        if (((float)counts[i] / values_count) * 100.0f <= clamp_margin_percent) { hi = T[i]; break; }
    for (int i = 0; i < 7; ++i)
        if (((float)counts[7 + i] / values_count) * 100.0f <= clamp_margin_percent) { lo = -T[i]; break; }
    range[0] = lo;
    range[1] = hi;
}

}  // namespace cdb
