// hnsw_build.cu -- GPU-side HNSW index build: the write-side twin of hnsw.cu.
//   index_embeddings / index_embedding   src/vector_store.rs:714-940
//   create_node_edges                    src/vector_store.rs:976-1070
//   ProbNode::add_neighbor (+ eviction)  src/models/prob_node.rs:210-283
//   get_max_insert_level, level probs    src/models/common.rs:373-379, 421-429 (factor 4, api_service.rs:109)
// The reference builds concurrently (rayon batches racing through per-node locks), so its graph is not
// reproducible and parity is defined for SEARCH on a given graph (tests: the graph built here is exported and the
// CUDA search is compared with the oracle's search on it).  The build keeps the reference's algorithm:
//   phase A (hnsw_build_search_kernel): every vector of a batch runs traverse_find_nearest with ef_construction on
//     every level from the top (fixed set pre-seeded with its own id, best 64 kept) -- the same device code as the
//     search kernel -- against the graph as of the start of the batch;
//   phase B (hnsw_build_link_kernel): create_node_edges for levels 0..max_level with the reference's add_neighbor
//     (lowest-similarity slot, strict improvement, stale lowest_index, back-link removal of the evicted neighbour,
//     remove_neighbor_by_index_and_id on a failed reverse edge) under a per-node spin lock.
// Batches grow geometrically from 1 so the upper structure is built almost sequentially.
// Level membership is decided up front from a counter RNG, so node_row / child are static and only the adjacency
// (plus each slot's similarity and the node's lowest slot) evolves.
#include <algorithm>
#include <vector>

#include "hnsw_traverse.cuh"

namespace cdb {

struct BuildLevel {
    uint32_t cnt, nb;
    uint32_t *node_row;    // [cnt] (level >= 1: [0] = root row, then ascending rows)
    uint32_t *child;       // [cnt] level-local index one level down
    uint32_t *adj;         // [cnt*nb]
    uint32_t *simkey;      // [cnt*nb] order key of the slot's similarity
    uint32_t *lowest_idx;  // [cnt]
    uint32_t *lowest_key;  // [cnt]
    int *lock;             // [cnt]
    // replica builds (collections with a metadata schema): [0] main root, [1] pseudo root, then created nodes in list order
    const uint32_t *node_id;   // [cnt] ProbNode::get_id()
    const uint32_t *node_md;   // [cnt] metadata table row or HN_EMPTY
    const uint32_t *node_ins;  // [cnt] position in the replica list (ascending from local 2; roots HN_EMPTY)
};
constexpr int BUILD_MAX_LEVELS = 16;
struct BuildGraph {
    uint32_t num_levels, entry, root_row;
    BuildLevel lv[BUILD_MAX_LEVELS];
    uint32_t min_key, max_key;  // MetricResult::min / ::max as order keys (types.rs:435-457)
    uint32_t pseudo_entry;      // replica builds: top-level index of the pseudo root
};

// the flattened IndexableEmbeddings of preprocess_embedding (vector_store.rs:629-712) on the device
struct ReplicaDev {
    const uint32_t *row, *id, *base_id, *md;   // [n_nodes]
    const int32_t *md_bits;                    // [n_md][M]
    const float *md_mags;
    uint32_t M;
    uint32_t key_one, key_minus_one;           // order keys of cs == 1.0 / -1.0 (edge rules, vector_store.rs:1014-1040)
};

struct BuildArgs {
    BuildGraph g;
    HnScoreCtx sc;
    const uint8_t *levels;  // [n] max insert level of each row
    uint32_t first, count;  // rows [first, first+count) form this batch
    uint32_t ef, shortlist;
    uint32_t *z_nodes;      // [count][num_levels+1][64]
    uint32_t *z_keys;       // [count][num_levels+1][64] order keys
    uint32_t *z_n;          // [count][num_levels+1]
    uint8_t *failed;        // [count]
    bool replicas;          // rows of the batch are positions of the replica list `rl`
    ReplicaDev rl;
};

// local index of `row` at `level` (level 0: the row itself; level >= 1: binary search in the sorted node_row)
__device__ inline uint32_t build_local(const BuildLevel &l, uint32_t level, uint32_t row, bool replicas = false) {
    if (replicas) {   // every level lists the created nodes in list order behind the two roots
        uint32_t lo = 2, hi = l.cnt;
        while (lo < hi) { const uint32_t md = (lo + hi) >> 1; if (l.node_ins[md] < row) lo = md + 1; else hi = md; }
        return lo;
    }
    if (level == 0) return row;
    uint32_t lo = 1, hi = l.cnt;
    while (lo < hi) { const uint32_t md = (lo + hi) >> 1; if (l.node_row[md] < row) lo = md + 1; else hi = md; }
    return lo;
}

__global__ void __launch_bounds__(HN_THREADS) hnsw_build_search_kernel(BuildArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const HnSmem m = hn_carve(smem, a.sc.row_pitch, a.ef);
    __shared__ HnShared sh;
    const uint32_t b = blockIdx.x, r = a.first + b;
    const int tid = threadIdx.x;
    const uint32_t vrow = a.replicas ? a.rl.row[r] : r;
    for (uint32_t i = tid; i < a.sc.row_pitch / 4; i += HN_THREADS)
        reinterpret_cast<uint32_t *>(m.qs)[i] = reinterpret_cast<const uint32_t *>(a.sc.rows + (size_t)vrow * a.sc.row_pitch)[i];
    const float qmag = a.sc.mags[vrow];
    const uint32_t max_level = a.levels[r];
    const uint32_t L1 = a.g.num_levels + 1;
    // replica builds: the traversal carries the embedding's metadata and prop_value.id (vector_store.rs:809-821); nodes whose
    // metadata has mag != 0 (Pseudo / Metadata kinds) start at the pseudo root (:461-483, 755-758)
    HnMdCtx md{};
    uint32_t self_id = r;
    bool under_pseudo = false;
    if (a.replicas) {
        const uint32_t mrow = a.rl.md[r];
        md.md_bits = a.rl.md_bits; md.md_mags = a.rl.md_mags; md.M = a.rl.M;
        md.q_bits = mrow == HN_EMPTY ? nullptr : a.rl.md_bits + (size_t)mrow * a.rl.M;
        md.q_mag = mrow == HN_EMPTY ? 0.0f : a.rl.md_mags[mrow];
        md.keep_fs = false; md.q_has_id = true; md.q_id = a.rl.base_id[r];
        self_id = a.rl.id[r];
        under_pseudo = md.q_bits && md.q_mag != 0.0f;
    }
    if (tid == 0) { sh.err = 0; sh.entry = under_pseudo ? a.g.pseudo_entry : a.g.entry; }
    unsigned long long evals = 0, pops = 0;
    __syncthreads();
    for (int level = (int)a.g.num_levels; level >= 0; --level) {
        const BuildLevel &l = a.g.lv[level];
        const uint32_t take = min(min(a.shortlist, l.nb), HN_MAX_TAKE);
        if (a.replicas) { md.node_id = l.node_id; md.node_md = l.node_md; }
        hn_traverse_level(l.node_row, l.adj, l.nb, take, a.sc, m, sh, qmag, self_id, a.ef, evals, pops, a.replicas ? &md : nullptr);
        if (sh.err) break;
        const uint32_t keep = min(sh.rlen, 64u);  // is_indexing: final_len = 64 (vector_store.rs:1194)
        if ((uint32_t)level <= max_level) {
            const size_t base = ((size_t)b * L1 + level) * 64;
            for (uint32_t i = tid; i < keep; i += HN_THREADS) {
                a.z_nodes[base + i] = m.rnodes[i];
                a.z_keys[base + i] = (uint32_t)(m.rkeys[i] >> 32);
            }
            if (tid == 0) a.z_n[(size_t)b * L1 + level] = keep;
        }
        if (tid == 0 && level > 0) sh.entry = l.child[m.rnodes[0]];
        __syncthreads();
    }
    if (tid == 0) a.failed[b] = sh.err ? 1 : 0;  // the reference returns Err: the embedding is not indexed
}

__device__ inline void node_lock(int *lk) {
    while (atomicCAS(lk, 0, 1) != 0) __nanosleep(32);
    __threadfence();
}
__device__ inline void node_unlock(int *lk) {
    __threadfence();
    atomicExch(lk, 0);
}

// ProbNode::add_neighbor; returns the slot index or -1
__device__ int build_add_neighbor(const BuildGraph &g, const BuildLevel &l, uint32_t node, uint32_t nbr, uint32_t dkey) {
    volatile uint32_t *adj = l.adj + (size_t)node * l.nb;
    volatile uint32_t *sk = l.simkey + (size_t)node * l.nb;
    node_lock(l.lock + node);
    const uint32_t lidx = ((volatile uint32_t *)l.lowest_idx)[node], lkey = ((volatile uint32_t *)l.lowest_key)[node];
    if (dkey <= lkey) { node_unlock(l.lock + node); return -1; }
    bool ok = false;
    uint32_t old = HN_EMPTY;
    if (adj[lidx] == HN_EMPTY) { adj[lidx] = nbr; sk[lidx] = dkey; ok = true; }
    else if (dkey > sk[lidx]) { old = adj[lidx]; adj[lidx] = nbr; sk[lidx] = dkey; ok = true; }
    uint32_t nidx = 0, nkey = g.max_key;
    for (uint32_t s = 0; s < l.nb; ++s) {
        if (adj[s] == HN_EMPTY) { nkey = g.min_key; nidx = s; break; }
        const uint32_t k = sk[s];
        if (k < nkey) { nkey = k; nidx = s; }
    }
    ((volatile uint32_t *)l.lowest_idx)[node] = nidx;
    ((volatile uint32_t *)l.lowest_key)[node] = nkey;
    node_unlock(l.lock + node);
    if (!ok) return -1;
    if (old != HN_EMPTY) {  // the evicted neighbour drops its back link (remove_neighbor_by_id), lowest_index left stale
        volatile uint32_t *oa = l.adj + (size_t)old * l.nb;
        node_lock(l.lock + old);
        for (uint32_t s = 0; s < l.nb; ++s)
            if (oa[s] == node) { oa[s] = HN_EMPTY; break; }
        node_unlock(l.lock + old);
    }
    return (int)lidx;
}

// ProbNode::replica_node_kind (prob_node.rs:487-496): the node's own id and metadata
__device__ inline int build_node_kind(const ReplicaDev &rl, const BuildLevel &l, uint32_t node) {
    const uint32_t mrow = l.node_md[node];
    if (mrow == HN_EMPTY || rl.md_mags[mrow] == 0.0f) return MD_KIND_BASE;
    const uint32_t id = l.node_id[node];
    return id >= MD_PSEUDO_ID_LO && id <= MD_PSEUDO_ID_HI ? MD_KIND_PSEUDO : MD_KIND_METADATA;
}

// one warp per new vector, lane 0 works (a thread spinning on a lock must not share a warp with its holder)
__global__ void __launch_bounds__(128) hnsw_build_link_kernel(BuildArgs a) {
    const uint32_t b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if ((threadIdx.x & 31) != 0 || b >= a.count || a.failed[b]) return;
    const uint32_t r = a.first + b;
    const uint32_t L1 = a.g.num_levels + 1;
    const uint32_t top = min((uint32_t)a.levels[r], a.g.num_levels);
    for (uint32_t level = 0; level <= top; ++level) {  // edges are created bottom-up as the recursion unwinds
        const BuildLevel &l = a.g.lv[level];
        const uint32_t node = build_local(l, level, r, a.replicas);
        const size_t base = ((size_t)b * L1 + level) * 64;
        const uint32_t zn = a.z_n[(size_t)b * L1 + level];
        const bool rules = a.replicas && a.sc.metric == CDB_METRIC_COSINE;   // the rules match MetricResult::CosineSimilarity only
        const int nk = rules ? build_node_kind(a.rl, l, node) : MD_KIND_BASE;
        uint32_t successful = 0;
        for (uint32_t i = 0; i < zn; ++i) {
            if (successful >= l.nb) break;
            const uint32_t nbr = a.z_nodes[base + i], dkey = a.z_keys[base + i];
            if (rules && nk == MD_KIND_METADATA) {   // vector_store.rs:1014-1040
                const int bk = build_node_kind(a.rl, l, nbr);
                if (bk == MD_KIND_PSEUDO && dkey != a.rl.key_one) continue;          // pseudo neighbour: perfect match only
                if (bk == MD_KIND_METADATA && dkey == a.rl.key_minus_one) continue;  // metadata neighbour: not on a strong mismatch
            }
            const int idx = build_add_neighbor(a.g, l, node, nbr, dkey);
            if (idx >= 0) {
                const int j = build_add_neighbor(a.g, l, nbr, node, dkey);
                if (j >= 0) successful++;
                else {  // remove_neighbor_by_index_and_id
                    volatile uint32_t *na = l.adj + (size_t)node * l.nb;
                    node_lock(l.lock + node);
                    if (na[idx] == nbr) na[idx] = HN_EMPTY;
                    node_unlock(l.lock + node);
                }
            }
        }
    }
}

__global__ void fill_u32_kernel(uint32_t *p, uint32_t v, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ------------------------------------------------------------------ host
struct BuiltGraph {
    BuildGraph g{};
    std::vector<void *> allocs;
    uint8_t *d_levels = nullptr;
};

static float unit_rng(uint64_t seed, uint64_t i) {  // stands in for rand::random::<f32>() (vector_store.rs:750-753)
    uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}

cdb_status hnsw_build_device(const HnScoreCtx &sc, uint32_t n /* data rows; root is row n */, uint32_t num_levels, uint32_t nbrs,
                             uint32_t nbrs0, uint32_t ef_construction, uint32_t shortlist, uint32_t max_batch, uint64_t seed,
                             GraphDev *out_graph, std::vector<void *> *out_allocs, std::vector<uint32_t> *out_counts,
                             std::vector<const uint32_t *> *out_nr, std::vector<const uint32_t *> *out_ad,
                             std::vector<const uint32_t *> *out_ch, cudaStream_t s) {
    if (num_levels + 1 > BUILD_MAX_LEVELS) { set_error("build: too many levels"); return CDB_INVALID_PARAMS; }
    const uint32_t L1 = num_levels + 1;
    // ---- level membership (get_max_insert_level with probabilities 1 - 4^-n)
    std::vector<uint8_t> levels(n);
    std::vector<std::vector<uint32_t>> rows(L1);
    for (uint32_t L = 1; L < L1; ++L) rows[L].push_back(n);  // root is local 0 on every upper level
    for (uint32_t r = 0; r < n; ++r) {
        const double x = (double)unit_rng(seed, r);
        uint32_t ml = 0;
        double p4 = 1.0;
        for (uint32_t lv = 1; lv <= num_levels; ++lv) { p4 *= 4.0; if (x >= 1.0 - 1.0 / p4) ml = lv; else break; }
        // x >= 1 - 4^-lv is monotone in lv, so the largest satisfied lv is the first hit of the reference's top-down scan
        levels[r] = (uint8_t)ml;
        for (uint32_t L = 1; L <= ml; ++L) rows[L].push_back(r);
    }
    BuiltGraph bg;
    bg.g.num_levels = num_levels;
    bg.g.root_row = n;
    bg.g.entry = num_levels == 0 ? n : 0;
    switch (sc.metric) {
    case CDB_METRIC_COSINE: bg.g.min_key = order_key(sc.metric, 0xBF800000u /* -1.0 */); bg.g.max_key = order_key(sc.metric, 0x40000000u /* 2.0 */); break;
    default: bg.g.min_key = order_key(sc.metric, 0xFF800000u /* -inf */); bg.g.max_key = order_key(sc.metric, 0x7F800000u /* +inf */); break;
    }
    auto dalloc = [&](void **p, size_t bytes) -> cdb_status {
        CDB_CUDA_TRY(cudaMalloc(p, bytes ? bytes : 4));
        out_allocs->push_back(*p);
        return CDB_OK;
    };
    cdb_status rc;
    std::vector<const uint32_t *> t_nr(L1), t_ad(L1), t_ch(L1);
    for (uint32_t L = 0; L < L1; ++L) {
        BuildLevel &l = bg.g.lv[L];
        l.nb = L == 0 ? nbrs0 : nbrs;
        l.cnt = L == 0 ? n + 1 : (uint32_t)rows[L].size();
        if ((rc = dalloc((void **)&l.node_row, (size_t)l.cnt * 4)) || (rc = dalloc((void **)&l.child, (size_t)l.cnt * 4)) ||
            (rc = dalloc((void **)&l.adj, (size_t)l.cnt * l.nb * 4)) || (rc = dalloc((void **)&l.simkey, (size_t)l.cnt * l.nb * 4)) ||
            (rc = dalloc((void **)&l.lowest_idx, (size_t)l.cnt * 4)) || (rc = dalloc((void **)&l.lowest_key, (size_t)l.cnt * 4)) ||
            (rc = dalloc((void **)&l.lock, (size_t)l.cnt * 4)))
            return rc;
        std::vector<uint32_t> nr(l.cnt), ch(l.cnt, HN_EMPTY);
        if (L == 0) for (uint32_t i = 0; i <= n; ++i) nr[i] = i;
        else {
            nr = rows[L];
            // child = position of the same row one level down (both lists: root first, then ascending rows)
            if (L == 1) for (uint32_t i = 0; i < l.cnt; ++i) ch[i] = nr[i];
            else {
                const std::vector<uint32_t> &below = rows[L - 1];
                uint32_t j = 1;
                ch[0] = 0;
                for (uint32_t i = 1; i < l.cnt; ++i) { while (below[j] != nr[i]) ++j; ch[i] = j; }
            }
        }
        CDB_CUDA_TRY(cudaMemcpyAsync(l.node_row, nr.data(), (size_t)l.cnt * 4, cudaMemcpyHostToDevice, s));
        CDB_CUDA_TRY(cudaMemcpyAsync(l.child, ch.data(), (size_t)l.cnt * 4, cudaMemcpyHostToDevice, s));
        CDB_CUDA_TRY(cudaStreamSynchronize(s));  // nr/ch are stack-owned
        const uint64_t slots = (uint64_t)l.cnt * l.nb;
        fill_u32_kernel<<<(uint32_t)((slots + 255) / 256), 256, 0, s>>>(l.adj, HN_EMPTY, slots);
        CDB_LAUNCH_CHECK();
        CDB_CUDA_TRY(cudaMemsetAsync(l.simkey, 0, slots * 4, s));
        CDB_CUDA_TRY(cudaMemsetAsync(l.lowest_idx, 0, (size_t)l.cnt * 4, s));  // ProbNode::new: (0, MetricResult::min)
        fill_u32_kernel<<<(l.cnt + 255) / 256, 256, 0, s>>>(l.lowest_key, bg.g.min_key, l.cnt);
        CDB_LAUNCH_CHECK();
        CDB_CUDA_TRY(cudaMemsetAsync(l.lock, 0, (size_t)l.cnt * 4, s));
        t_nr[L] = l.node_row; t_ad[L] = l.adj; t_ch[L] = l.child;
    }
    uint8_t *d_levels = nullptr;
    if ((rc = dalloc((void **)&d_levels, n))) return rc;
    CDB_CUDA_TRY(cudaMemcpyAsync(d_levels, levels.data(), n, cudaMemcpyHostToDevice, s));
    // ---- batches
    if (max_batch == 0) max_batch = 4096;
    uint32_t *z_nodes = nullptr, *z_keys = nullptr, *z_n = nullptr;
    uint8_t *failed = nullptr;
    CDB_CUDA_TRY(cudaMalloc(&z_nodes, (size_t)max_batch * L1 * 64 * 4));
    CDB_CUDA_TRY(cudaMalloc(&z_keys, (size_t)max_batch * L1 * 64 * 4));
    CDB_CUDA_TRY(cudaMalloc(&z_n, (size_t)max_batch * L1 * 4));
    CDB_CUDA_TRY(cudaMalloc(&failed, max_batch));
    const size_t smem = hn_smem_bytes(sc.row_pitch, ef_construction);
    CDB_ALLOW_SMEM(hnsw_build_search_kernel, smem);
    BuildArgs a{};
    a.g = bg.g;
    a.sc = sc;
    a.sc.root_row = n;
    a.levels = d_levels;
    a.ef = ef_construction;
    a.shortlist = shortlist;
    a.z_nodes = z_nodes; a.z_keys = z_keys; a.z_n = z_n; a.failed = failed;
    cdb_status result = CDB_OK;
    for (uint32_t done = 0; done < n && result == CDB_OK;) {
        uint32_t batch = std::max<uint32_t>(1, std::min<uint32_t>(max_batch, done / 8));
        batch = std::min<uint32_t>(batch, n - done);
        a.first = done;
        a.count = batch;
        hnsw_build_search_kernel<<<batch, HN_THREADS, smem, s>>>(a);
        g_launch_count.fetch_add(1, std::memory_order_relaxed);
        hnsw_build_link_kernel<<<(batch * 32 + 127) / 128, 128, 0, s>>>(a);
        g_launch_count.fetch_add(1, std::memory_order_relaxed);
        done += batch;
    }
    cudaError_t e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaFree(z_nodes); cudaFree(z_keys); cudaFree(z_n); cudaFree(failed);
    if (e != cudaSuccess) { set_error(std::string("hnsw build: ") + cudaGetErrorString(e)); return CDB_CUDA_ERROR; }
    // ---- publish as a search graph
    const std::vector<const uint32_t *> *src[3] = {&t_nr, &t_ad, &t_ch};
    const uint32_t *tbl[3];
    for (int t = 0; t < 3; ++t) {
        void *p = nullptr;
        if ((rc = dalloc(&p, L1 * sizeof(void *)))) return rc;
        CDB_CUDA_TRY(cudaMemcpy(p, src[t]->data(), L1 * sizeof(void *), cudaMemcpyHostToDevice));
        tbl[t] = (const uint32_t *)p;
    }
    out_counts->clear();
    for (uint32_t L = 0; L < L1; ++L) out_counts->push_back(bg.g.lv[L].cnt);
    *out_nr = t_nr; *out_ad = t_ad; *out_ch = t_ch;
    out_graph->num_levels = num_levels;
    out_graph->nbrs = nbrs;
    out_graph->nbrs0 = nbrs0;
    out_graph->entry = bg.g.entry;
    out_graph->root_row = n;
    out_graph->identity_mask = 1u;   // level 0: node i is row i (lv_local), the search skips the node_row lookup there
    out_graph->node_row = reinterpret_cast<const uint32_t *const *>(tbl[0]);
    out_graph->adj = reinterpret_cast<const uint32_t *const *>(tbl[1]);
    out_graph->child = reinterpret_cast<const uint32_t *const *>(tbl[2]);
    return CDB_OK;
}

// ------------------------------------------------------------------ replica lists (collections with a metadata schema)
// index_embeddings over the flattened IndexableEmbeddings (vector_store.rs:714-780): host arrays, one entry per node to create.
// Level layout: [0] main root, [1] pseudo root (both on every level, create_root_node / create_pseudo_root_node), then the
// created nodes in list order.  The same two kernels as above do the work; level draws come from the caller.
cdb_status hnsw_build_replicas_device(const HnScoreCtx &sc, const ReplicaHost &rh, uint32_t num_levels, uint32_t nbrs, uint32_t nbrs0,
                                      uint32_t ef_construction, uint32_t shortlist, uint32_t max_batch, GraphDev *out_graph,
                                      std::vector<void *> *out_allocs, std::vector<uint32_t> *out_counts,
                                      std::vector<const uint32_t *> *out_nr, std::vector<const uint32_t *> *out_ad,
                                      std::vector<const uint32_t *> *out_ch, ReplicaGraphDev *out_md, std::vector<uint8_t> *out_failed,
                                      cudaStream_t s) {
    if (num_levels + 1 > BUILD_MAX_LEVELS) { set_error("build: too many levels"); return CDB_INVALID_PARAMS; }
    const uint32_t L1 = num_levels + 1, n = rh.n_nodes;
    std::vector<uint8_t> levels(n);
    std::vector<std::vector<uint32_t>> members(L1);   // list positions present on each level
    for (uint32_t t = 0; t < n; ++t) {
        levels[t] = (uint8_t)std::min<uint32_t>(rh.max_level[t], num_levels);
        for (uint32_t L = 0; L <= levels[t]; ++L) members[L].push_back(t);
    }
    BuiltGraph bg;
    bg.g.num_levels = num_levels;
    bg.g.root_row = rh.main_root_row;
    bg.g.entry = 0;
    bg.g.pseudo_entry = 1;
    switch (sc.metric) {
    case CDB_METRIC_COSINE: bg.g.min_key = order_key(sc.metric, 0xBF800000u); bg.g.max_key = order_key(sc.metric, 0x40000000u); break;
    default: bg.g.min_key = order_key(sc.metric, 0xFF800000u); bg.g.max_key = order_key(sc.metric, 0x7F800000u); break;
    }
    auto dalloc = [&](void **p, size_t bytes) -> cdb_status {
        CDB_CUDA_TRY(cudaMalloc(p, bytes ? bytes : 4));
        out_allocs->push_back(*p);
        return CDB_OK;
    };
    auto upload = [&](const void *src, size_t bytes, void **dst) -> cdb_status {
        cdb_status rc = dalloc(dst, bytes);
        if (rc) return rc;
        if (bytes) CDB_CUDA_TRY(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
        return CDB_OK;
    };
    cdb_status rc;
    ReplicaDev rl{};
    void *p = nullptr;
    if ((rc = upload(rh.row, (size_t)n * 4, &p))) return rc;
    rl.row = (const uint32_t *)p;
    if ((rc = upload(rh.node_id, (size_t)n * 4, &p))) return rc;
    rl.id = (const uint32_t *)p;
    if ((rc = upload(rh.base_id, (size_t)n * 4, &p))) return rc;
    rl.base_id = (const uint32_t *)p;
    if ((rc = upload(rh.md_row, (size_t)n * 4, &p))) return rc;
    rl.md = (const uint32_t *)p;
    if ((rc = upload(rh.md_bits, (size_t)rh.n_md * rh.md_dims * 4, &p))) return rc;
    rl.md_bits = (const int32_t *)p;
    if ((rc = upload(rh.md_mags, (size_t)rh.n_md * 4, &p))) return rc;
    rl.md_mags = (const float *)p;
    rl.M = rh.md_dims;
    rl.key_one = order_key(CDB_METRIC_COSINE, 0x3F800000u);
    rl.key_minus_one = order_key(CDB_METRIC_COSINE, 0xBF800000u);
    std::vector<const uint32_t *> t_nr(L1), t_ad(L1), t_ch(L1), t_id(L1), t_md(L1);
    for (uint32_t L = 0; L < L1; ++L) {
        BuildLevel &l = bg.g.lv[L];
        const std::vector<uint32_t> &mem = members[L];
        l.nb = L == 0 ? nbrs0 : nbrs;
        l.cnt = 2 + (uint32_t)mem.size();
        std::vector<uint32_t> nr(l.cnt), ch(l.cnt, HN_EMPTY), ni(l.cnt), nm(l.cnt), ins(l.cnt, HN_EMPTY);
        nr[0] = rh.main_root_row; ni[0] = 0xFFFFFFFFu; nm[0] = rh.main_root_md;
        nr[1] = rh.pseudo_root_row; ni[1] = MD_PSEUDO_ID_LO; nm[1] = rh.pseudo_root_md;
        if (L > 0) { ch[0] = 0; ch[1] = 1; }
        uint32_t j = 0;   // position in the level below (both lists ascend)
        for (uint32_t i = 0; i < mem.size(); ++i) {
            const uint32_t t = mem[i];
            nr[2 + i] = rh.row[t]; ni[2 + i] = rh.node_id[t]; nm[2 + i] = rh.md_row[t]; ins[2 + i] = t;
            if (L > 0) { while (members[L - 1][j] != t) ++j; ch[2 + i] = 2 + j; }
        }
        void *d_nr, *d_ch, *d_ni, *d_nm, *d_ins;
        if ((rc = upload(nr.data(), (size_t)l.cnt * 4, &d_nr)) || (rc = upload(ch.data(), (size_t)l.cnt * 4, &d_ch)) ||
            (rc = upload(ni.data(), (size_t)l.cnt * 4, &d_ni)) || (rc = upload(nm.data(), (size_t)l.cnt * 4, &d_nm)) ||
            (rc = upload(ins.data(), (size_t)l.cnt * 4, &d_ins)))
            return rc;
        l.node_row = (uint32_t *)d_nr; l.child = (uint32_t *)d_ch;
        l.node_id = (const uint32_t *)d_ni; l.node_md = (const uint32_t *)d_nm; l.node_ins = (const uint32_t *)d_ins;
        if ((rc = dalloc((void **)&l.adj, (size_t)l.cnt * l.nb * 4)) || (rc = dalloc((void **)&l.simkey, (size_t)l.cnt * l.nb * 4)) ||
            (rc = dalloc((void **)&l.lowest_idx, (size_t)l.cnt * 4)) || (rc = dalloc((void **)&l.lowest_key, (size_t)l.cnt * 4)) ||
            (rc = dalloc((void **)&l.lock, (size_t)l.cnt * 4)))
            return rc;
        const uint64_t slots = (uint64_t)l.cnt * l.nb;
        fill_u32_kernel<<<(uint32_t)((slots + 255) / 256), 256, 0, s>>>(l.adj, HN_EMPTY, slots);
        CDB_LAUNCH_CHECK();
        CDB_CUDA_TRY(cudaMemsetAsync(l.simkey, 0, slots * 4, s));
        CDB_CUDA_TRY(cudaMemsetAsync(l.lowest_idx, 0, (size_t)l.cnt * 4, s));
        fill_u32_kernel<<<(l.cnt + 255) / 256, 256, 0, s>>>(l.lowest_key, bg.g.min_key, l.cnt);
        CDB_LAUNCH_CHECK();
        CDB_CUDA_TRY(cudaMemsetAsync(l.lock, 0, (size_t)l.cnt * 4, s));
        t_nr[L] = l.node_row; t_ad[L] = l.adj; t_ch[L] = l.child; t_id[L] = l.node_id; t_md[L] = l.node_md;
    }
    uint8_t *d_levels = nullptr;
    if ((rc = upload(levels.data(), n, (void **)&d_levels))) return rc;
    if (max_batch == 0) max_batch = 4096;
    max_batch = std::min<uint32_t>(max_batch, std::max<uint32_t>(n, 1));
    uint32_t *z_nodes = nullptr, *z_keys = nullptr, *z_n = nullptr;
    uint8_t *failed = nullptr;
    CDB_CUDA_TRY(cudaMalloc(&z_nodes, (size_t)max_batch * L1 * 64 * 4));
    CDB_CUDA_TRY(cudaMalloc(&z_keys, (size_t)max_batch * L1 * 64 * 4));
    CDB_CUDA_TRY(cudaMalloc(&z_n, (size_t)max_batch * L1 * 4));
    CDB_CUDA_TRY(cudaMalloc(&failed, (size_t)std::max<uint32_t>(n, 1)));
    const size_t smem = hn_smem_bytes(sc.row_pitch, ef_construction);
    CDB_ALLOW_SMEM(hnsw_build_search_kernel, smem);
    BuildArgs a{};
    a.g = bg.g;
    a.sc = sc;
    a.sc.root_row = rh.main_root_row;
    a.levels = d_levels;
    a.ef = ef_construction;
    a.shortlist = shortlist;
    a.z_nodes = z_nodes; a.z_keys = z_keys; a.z_n = z_n;
    a.replicas = true;
    a.rl = rl;
    for (uint32_t done = 0; done < n;) {
        uint32_t batch = std::max<uint32_t>(1, std::min<uint32_t>(max_batch, done / 8));
        batch = std::min<uint32_t>(batch, n - done);
        a.first = done;
        a.count = batch;
        a.failed = failed + done;   // per list position, read back below
        hnsw_build_search_kernel<<<batch, HN_THREADS, smem, s>>>(a);
        g_launch_count.fetch_add(1, std::memory_order_relaxed);
        hnsw_build_link_kernel<<<(batch * 32 + 127) / 128, 128, 0, s>>>(a);
        g_launch_count.fetch_add(1, std::memory_order_relaxed);
        done += batch;
    }
    cudaError_t e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess && out_failed) {
        out_failed->assign(n, 0);
        if (n) e = cudaMemcpy(out_failed->data(), failed, n, cudaMemcpyDeviceToHost);
    }
    cudaFree(z_nodes); cudaFree(z_keys); cudaFree(z_n); cudaFree(failed);
    if (e != cudaSuccess) { set_error(std::string("hnsw replica build: ") + cudaGetErrorString(e)); return CDB_CUDA_ERROR; }
    const std::vector<const uint32_t *> *src[5] = {&t_nr, &t_ad, &t_ch, &t_id, &t_md};
    const uint32_t *tbl[5];
    for (int t = 0; t < 5; ++t) {
        void *q = nullptr;
        if ((rc = upload(src[t]->data(), L1 * sizeof(void *), &q))) return rc;
        tbl[t] = (const uint32_t *)q;
    }
    out_counts->clear();
    for (uint32_t L = 0; L < L1; ++L) out_counts->push_back(bg.g.lv[L].cnt);
    *out_nr = t_nr; *out_ad = t_ad; *out_ch = t_ch;
    out_graph->num_levels = num_levels;
    out_graph->nbrs = nbrs;
    out_graph->nbrs0 = nbrs0;
    out_graph->entry = 0;
    out_graph->root_row = rh.main_root_row;
    out_graph->identity_mask = 0u;
    out_graph->node_row = reinterpret_cast<const uint32_t *const *>(tbl[0]);
    out_graph->adj = reinterpret_cast<const uint32_t *const *>(tbl[1]);
    out_graph->child = reinterpret_cast<const uint32_t *const *>(tbl[2]);
    out_md->node_id = reinterpret_cast<const uint32_t *const *>(tbl[3]);
    out_md->node_md = reinterpret_cast<const uint32_t *const *>(tbl[4]);
    out_md->h_node_id = t_id;
    out_md->h_node_md = t_md;
    out_md->md_bits = rl.md_bits;
    out_md->md_mags = rl.md_mags;
    out_md->md_dims = rh.md_dims;
    out_md->pseudo_entry = 1;
    return CDB_OK;
}

}  // namespace cdb
