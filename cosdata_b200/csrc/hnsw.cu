// hnsw.cu -- batched HNSW search on a flat graph (S1, CDB_MODE_HNSW):
//   ann_search              src/vector_store.rs:256-402
//   traverse_find_nearest   src/vector_store.rs:1112-1204
//   PerformantFixedSet      src/models/fixedset.rs:2-29
//   remove_duplicates_and_filter  src/models/common.rs:381-412
// One CTA per query (the data-parallel axis of IndexOps::batch_search); all queries of a batch
// are resident at once.  Per pop the CTA
//   1. walks the popped node's first `shortlist_size` slots IN ORDER through the lossy fixed set
//      (one thread: the membership test + insert is order dependent, aliasing ids included),
//   2. scores the surviving neighbours in parallel, one thread per neighbour, with the reference's
//      exact per-pair arithmetic (pair_distance), gathering the rows straight from HBM,
//   3. merges them into the candidate queue.
// The reference's BinaryHeap is unbounded but performs exactly `ef` pops, so an entry ranked
// below the number of pops still to come can never be popped: the queue is a sorted array of at
// most `ef` entries and the pop order is identical to the heap's.  Keys are
// (order_key(score) << 32 | ~id): "better score first, then smaller id" -- the oracle's tie rule.
#include "kernels.h"

namespace cdb {

constexpr int HN_THREADS = 128;
constexpr uint32_t HN_EMPTY = 0xFFFFFFFFu;
constexpr uint32_t HN_ROOT_ID = 0xFFFFFFFFu;
constexpr uint32_t HN_QUERY_ID = 0xFFFFFFFEu;  // hnsw/mod.rs:398
constexpr uint32_t HN_MAX_TAKE = 64;            // slots examined per pop (<= shortlist_size, config.toml:32)
constexpr uint32_t HN_FINAL_LEN = 100;          // vector_store.rs:1194

__device__ __forceinline__ uint32_t hn_id(const GraphDev &g, uint32_t row) { return row == g.root_row ? HN_ROOT_ID : row; }

// bitonic sort (descending) of n keys with a payload, padded to P (power of two) with zeros
__device__ void hn_sort_desc(uint64_t *keys, uint32_t *vals, uint32_t n, uint32_t P) {
    for (uint32_t i = n + threadIdx.x; i < P; i += blockDim.x) { keys[i] = 0ull; vals[i] = 0; }
    __syncthreads();
    for (uint32_t size = 2; size <= P; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = threadIdx.x; t < P / 2; t += blockDim.x) {
                const uint32_t lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                const bool desc = (lo & size) == 0;
                const uint64_t x = keys[lo], y = keys[hi];
                if ((x < y) == desc) {
                    keys[lo] = y; keys[hi] = x;
                    const uint32_t v = vals[lo]; vals[lo] = vals[hi]; vals[hi] = v;
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(HN_THREADS) hnsw_search_kernel(HnswArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t EFP = [&] { uint32_t p = 1; while (p < a.ef) p <<= 1; return p < 128 ? 128u : p; }();  // queue / result capacity
    uint8_t *qs = smem;                                                   // [row_pitch]
    uint64_t *qkeys = reinterpret_cast<uint64_t *>(qs + round_up(a.row_pitch, 16));  // [2][EFP]
    uint64_t *rkeys = qkeys + 2 * EFP;                                    // [EFP]
    uint64_t *nkeys = rkeys + EFP;                                        // [64] new candidates, then sorted copy [64]
    uint64_t *fs = nkeys + 2 * HN_MAX_TAKE;                               // [64] fixed set buckets
    uint32_t *qnodes = reinterpret_cast<uint32_t *>(fs + 64);             // [2][EFP]
    uint32_t *rnodes = qnodes + 2 * EFP;                                  // [EFP]
    uint32_t *nnodes = rnodes + EFP;                                      // [64] + sorted [64]
    __shared__ uint32_t s_qlen, s_cur, s_visited, s_rlen, s_ncand, s_err, s_best_node, s_entry;
    __shared__ uint32_t s_bitkey[HN_MAX_TAKE];

    const uint32_t qi = blockIdx.x;
    const int tid = threadIdx.x;
    for (uint32_t i = tid; i < a.row_pitch / 4; i += HN_THREADS)
        reinterpret_cast<uint32_t *>(qs)[i] = reinterpret_cast<const uint32_t *>(a.q + (size_t)qi * a.row_pitch)[i];
    const float qmag = a.qmags[qi];
    const uint32_t pp = plane_pitch(a.dim);
    if (tid == 0) { s_err = 0; s_entry = a.g.entry; }
    uint32_t out_total = 0;  // meaningful in thread 0
    unsigned long long evals = 0, pops = 0;
    __syncthreads();

    for (int level = (int)a.g.num_levels; level >= 0; --level) {
        const uint32_t nb = level == 0 ? a.g.nbrs0 : a.g.nbrs;
        const uint32_t take = min(min(a.shortlist, nb), HN_MAX_TAKE);
        const uint32_t *node_row = a.g.node_row[level];
        const uint32_t *adj = a.g.adj[level];
        // fresh fixed set per level, pre-seeded with the query id (vector_store.rs:266-271)
        if (tid < 64) fs[tid] = 0ull;
        __syncthreads();
        if (tid == 0) {
            const uint32_t mask = nb - 1u;
            fs[(HN_QUERY_ID >> 6) & mask] |= 1ull << (HN_QUERY_ID & 0x3f);
            const uint32_t entry = s_entry;
            const uint32_t erow = node_row[entry];
            float d = 0.f;
            const int rc = pair_distance(a.metric, a.st, a.dim, qs, qmag, pp, a.rows + (size_t)erow * a.row_pitch, a.mags[erow], pp, &d);
            evals++;
            if (rc != CDB_OK) s_err = rc == CDB_CALCULATION_ERROR ? CDB_ERRFLAG_CALCULATION : 2;
            const uint32_t eid = hn_id(a.g, erow);
            fs[(eid >> 6) & mask] |= 1ull << (eid & 0x3f);
            qkeys[0] = make_key64(order_key(a.metric, __float_as_uint(d)), eid);
            qnodes[0] = entry;
            s_qlen = 1; s_cur = 0; s_visited = 0; s_rlen = 0;
        }
        __syncthreads();
        if (s_err) break;

        while (true) {
            const uint32_t qlen = s_qlen, cur = s_cur, visited = s_visited;
            if (qlen == 0 || visited >= a.ef) break;
            uint64_t *Q = qkeys + cur * EFP;
            uint32_t *QN = qnodes + cur * EFP;
            // ---- pop (one thread), then the walk through the lossy fixed set for all slots at once.
            // The reference tests and inserts slot by slot: a slot is scored iff its bit is not yet set AND no
            // earlier non-empty slot of this pop maps to the same bit (that one either set it or found it set).
            const uint32_t bn = QN[0];
            if (tid == 0) {
                rkeys[s_rlen] = Q[0]; rnodes[s_rlen] = bn; s_rlen++;
                pops++;
                s_ncand = 0;
            }
            uint32_t my_nbl = HN_EMPTY, my_bitkey = 0xFFFFFFFFu;
            if ((uint32_t)tid < take) {
                my_nbl = adj[(size_t)bn * nb + tid];
                if (my_nbl != HN_EMPTY) {
                    const uint32_t id = hn_id(a.g, node_row[my_nbl]);
                    my_bitkey = (((id >> 6) & (nb - 1u)) << 6) | (id & 0x3f);
                }
                s_bitkey[tid] = my_bitkey;
            }
            __syncthreads();
            bool accept = false;
            if (my_bitkey != 0xFFFFFFFFu) {
                accept = ((fs[my_bitkey >> 6] >> (my_bitkey & 0x3f)) & 1ull) == 0;
                for (int s2 = 0; s2 < tid && accept; ++s2) accept = s_bitkey[s2] != my_bitkey;
            }
            __syncthreads();  // every thread has read the old fixed set
            if (accept) {
                atomicOr(reinterpret_cast<unsigned long long *>(&fs[my_bitkey >> 6]), 1ull << (my_bitkey & 0x3f));
                nnodes[atomicAdd(&s_ncand, 1u)] = my_nbl;
            }
            __syncthreads();
            const uint32_t nc = s_ncand;
            // ---- score the new neighbours, one thread each, reference arithmetic
            if ((uint32_t)tid < nc) {
                const uint32_t nbl = nnodes[tid];
                const uint32_t row = node_row[nbl];
                float d = 0.f;
                const int rc = pair_distance(a.metric, a.st, a.dim, qs, qmag, pp, a.rows + (size_t)row * a.row_pitch, a.mags[row], pp, &d);
                if (rc != CDB_OK) atomicOr(&s_err, rc == CDB_CALCULATION_ERROR ? (uint32_t)CDB_ERRFLAG_CALCULATION : 2u);
                nkeys[tid] = make_key64(order_key(a.metric, __float_as_uint(d)), hn_id(a.g, row));
            }
            if (tid == 0) evals += nc;
            __syncthreads();
            if (s_err) break;
            // ---- sort the new entries (rank sort, nc <= 64) into nkeys[64..], nnodes[64..]
            if ((uint32_t)tid < nc) {
                const uint64_t k = nkeys[tid];
                uint32_t r = 0;
                for (uint32_t j = 0; j < nc; ++j) r += nkeys[j] > k;
                nkeys[HN_MAX_TAKE + r] = k;
                nnodes[HN_MAX_TAKE + r] = nnodes[tid];
            }
            __syncthreads();
            // ---- merge old queue (minus the popped head) with the new entries, keep what can still be popped
            {
                const uint64_t *NK = nkeys + HN_MAX_TAKE;
                const uint32_t *NN = nnodes + HN_MAX_TAKE;
                const uint32_t oldn = qlen - 1;
                const uint32_t cap = min(a.ef - (visited + 1), EFP);  // pops still to come
                uint64_t *D = qkeys + (cur ^ 1) * EFP;
                uint32_t *DN = qnodes + (cur ^ 1) * EFP;
                for (uint32_t i = tid; i < oldn; i += HN_THREADS) {
                    const uint64_t k = Q[1 + i];
                    uint32_t lo = 0, hi = nc;  // number of new entries better than k
                    while (lo < hi) { const uint32_t m = (lo + hi) >> 1; if (NK[m] > k) lo = m + 1; else hi = m; }
                    const uint32_t pos = i + lo;
                    if (pos < cap) { D[pos] = k; DN[pos] = QN[1 + i]; }
                }
                for (uint32_t j = tid; j < nc; j += HN_THREADS) {
                    const uint64_t k = NK[j];
                    uint32_t lo = 0, hi = oldn;  // number of old entries better than k
                    while (lo < hi) { const uint32_t m = (lo + hi) >> 1; if (Q[1 + m] > k) lo = m + 1; else hi = m; }
                    const uint32_t pos = j + lo;
                    if (pos < cap) { D[pos] = k; DN[pos] = NN[j]; }
                }
                __syncthreads();
                if (tid == 0) { s_qlen = min(oldn + nc, cap); s_cur = cur ^ 1; s_visited = visited + 1; }
            }
            __syncthreads();
        }
        __syncthreads();
        if (s_err) break;
        // ---- results of this level: sort best-first, keep 100, append; child of the best is the next entry
        const uint32_t rlen = s_rlen;
        uint32_t P = 1;
        while (P < rlen) P <<= 1;
        hn_sort_desc(rkeys, rnodes, rlen, P);
        const uint32_t keep = min(rlen, HN_FINAL_LEN);
        const uint32_t base = out_total;
        for (uint32_t i = tid; i < keep; i += HN_THREADS) {
            const uint32_t slot = base + i;
            if (slot < a.out_cap) {
                a.out_rows[(size_t)qi * a.out_cap + slot] = node_row[rnodes[i]];
                a.out_scores[(size_t)qi * a.out_cap + slot] = __uint_as_float(key_to_bits(a.metric, (uint32_t)(rkeys[i] >> 32)));
            }
        }
        out_total += keep;
        if (tid == 0 && level > 0) s_entry = a.g.child[level][rnodes[0]];
        __syncthreads();
    }
    if (tid == 0) {
        a.out_n[qi] = s_err ? 0u : min(out_total, a.out_cap);
        if (s_err) atomicOr(a.err32 + qi, s_err);
        if (a.counters) { atomicAdd(a.counters, evals); atomicAdd(a.counters + 1, pops); }
    }
}

// remove_duplicates_and_filter: dedup by id keeping the first occurrence, drop the root, sort
// best-first, truncate to 5*k.  One CTA per query; writes global candidate ids for the re-rank.
__global__ void __launch_bounds__(256) hnsw_dedup_kernel(const uint32_t *__restrict__ rows, const float *__restrict__ scores,
                                                         const uint32_t *__restrict__ n_in, uint32_t in_cap, int metric,
                                                         uint32_t root_row, uint32_t id_base, uint32_t k5, uint32_t *__restrict__ cand,
                                                         uint32_t *__restrict__ cand_cnt) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint64_t *keys = reinterpret_cast<uint64_t *>(smem);
    uint32_t P = 1;
    while (P < in_cap) P <<= 1;
    uint32_t *vals = reinterpret_cast<uint32_t *>(keys + P);
    uint32_t *rws = vals + P;
    __shared__ int kept;
    const uint32_t q = blockIdx.x, n = min(n_in[q], in_cap);
    if (threadIdx.x == 0) kept = 0;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) rws[i] = rows[(size_t)q * in_cap + i];
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t r = rws[i];
        bool dup = false;
        for (uint32_t j = 0; j < i; ++j) dup |= rws[j] == r;
        uint64_t key = 0;
        if (!dup && r != root_row) {
            key = make_key64(order_key(metric, __float_as_uint(scores[(size_t)q * in_cap + i])), r);
            atomicAdd(&kept, 1);
        }
        keys[i] = key;
        vals[i] = r;
    }
    __syncthreads();
    // sort descending (0 = dropped entries last)
    for (uint32_t i = n + threadIdx.x; i < P; i += blockDim.x) { keys[i] = 0ull; vals[i] = 0; }
    __syncthreads();
    for (uint32_t size = 2; size <= P; size <<= 1)
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = threadIdx.x; t < P / 2; t += blockDim.x) {
                const uint32_t lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                const bool desc = (lo & size) == 0;
                const uint64_t x = keys[lo], y = keys[hi];
                if ((x < y) == desc) { keys[lo] = y; keys[hi] = x; const uint32_t v = vals[lo]; vals[lo] = vals[hi]; vals[hi] = v; }
            }
            __syncthreads();
        }
    const uint32_t m = min((uint32_t)kept, k5);
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) cand[(size_t)q * k5 + i] = id_base + vals[i];
    if (threadIdx.x == 0) cand_cnt[q] = m;
}

size_t hnsw_search_smem(uint32_t row_pitch, uint32_t ef) {
    uint32_t efp = 1;
    while (efp < ef) efp <<= 1;
    if (efp < 128) efp = 128;
    return round_up(row_pitch, 16) + (size_t)(3 * efp + 2 * HN_MAX_TAKE + 64) * 8 + (size_t)(3 * efp + 2 * HN_MAX_TAKE) * 4 + 64;
}

cdb_status hnsw_search_device(const HnswArgs &a, cudaStream_t s) {
    if (!a.nq) return CDB_OK;
    if (a.ef == 0 || a.ef > 4096) { set_error("hnsw: ef_search must be in 1..4096"); return CDB_INVALID_PARAMS; }
    const size_t smem = hnsw_search_smem(a.row_pitch, a.ef);
    if (smem > 200 * 1024) { set_error("hnsw: ef_search too large for shared memory"); return CDB_INVALID_PARAMS; }
    CDB_CUDA_TRY(cudaFuncSetAttribute(hnsw_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hnsw_search_kernel<<<a.nq, HN_THREADS, smem, s>>>(a);
    CDB_LAUNCH_CHECK();
    return CDB_OK;
}

cdb_status hnsw_dedup_device(const uint32_t *d_rows, const float *d_scores, const uint32_t *d_n, uint32_t in_cap, int metric,
                             uint32_t root_row, uint32_t id_base, uint32_t k5, uint32_t nq, uint32_t *d_cand, uint32_t *d_cand_cnt,
                             cudaStream_t s) {
    if (!nq) return CDB_OK;
    uint32_t P = 1;
    while (P < in_cap) P <<= 1;
    const size_t smem = (size_t)P * 12 + (size_t)in_cap * 4 + 16;
    if (smem > 200 * 1024) { set_error("hnsw dedup: too many levels"); return CDB_INVALID_PARAMS; }
    CDB_CUDA_TRY(cudaFuncSetAttribute(hnsw_dedup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hnsw_dedup_kernel<<<nq, 256, smem, s>>>(d_rows, d_scores, d_n, in_cap, metric, root_row, id_base, k5, d_cand, d_cand_cnt);
    CDB_LAUNCH_CHECK();
    return CDB_OK;
}

}  // namespace cdb
