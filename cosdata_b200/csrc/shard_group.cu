// shard_group.cu -- row-sharded search over the GPUs of one node behind the C ABI (the reference's call site is
// IndexOps::batch_search, src/indexes/mod.rs:260-272 -- one process, rayon over queries; here one index shard per GPU).
//
// Every shard searches the same query batch; its final kernel (exact re-rank / merge / key selection) writes the packed
// 64-bit selection keys (order_key(score) << 32 | ~global_id, best first, 0 = empty) of every query STRAIGHT INTO ITS SLOT of
// the gather buffer, the per-query error bytes right behind them; ONE ncclAllGather of B*k*8 (+ B) bytes per rank makes
// every rank hold all slots; one merge kernel picks the global top-k per query with the common ordering rule.  No
// id/score arrays are exchanged, no pack kernel, no second collective.
//
// Two deployments, same code:
//   cdb_shard_group_create       one process, n devices (ncclCommInitAll); the host calls cdb_search_batch_sharded once
//   cdb_shard_group_create_rank  one process per GPU (torchrun style; rank 0 distributes cdb_nccl_unique_id())
// NCCL is bound at run time (dlopen of libnccl.so.2, the copy already in the process if there is one), so the library
// itself has no NCCL link dependency and a single-GPU host never needs it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "kernels.h"

struct cdb_index;
namespace cdb {
cdb_status index_search_device_keys(cdb_index *ix, const float *d_queries, uint32_t nq, const cdb_search_params *p, uint32_t *d_ids,
                                    float *d_scores, uint32_t *d_counts, uint8_t *d_err, uint64_t *d_keys, cudaStream_t s);
int index_device(const cdb_index *ix);
uint32_t index_dim(const cdb_index *ix);
int index_result_metric(const cdb_index *ix, int mode);

namespace {
struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi *nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy a host framework already loaded, if any
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        api.h = h;
#define CDB_NCCL_SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name))
        CDB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
        CDB_NCCL_SYM(CommInitRank, "ncclCommInitRank");
        CDB_NCCL_SYM(CommInitAll, "ncclCommInitAll");
        CDB_NCCL_SYM(CommDestroy, "ncclCommDestroy");
        CDB_NCCL_SYM(AllGather, "ncclAllGather");
        CDB_NCCL_SYM(GroupStart, "ncclGroupStart");
        CDB_NCCL_SYM(GroupEnd, "ncclGroupEnd");
        CDB_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef CDB_NCCL_SYM
        if (!api.GetUniqueId || !api.CommInitRank || !api.CommInitAll || !api.CommDestroy || !api.AllGather || !api.GroupStart ||
            !api.GroupEnd || !api.GetErrorString)
            api.h = nullptr;
    });
    return api.h ? &api : nullptr;
}
#define CDB_NCCL_TRY(api, expr)                                                                       \
    do {                                                                                              \
        ncclResult_t _r = (expr);                                                                     \
        if (_r != ncclSuccess) {                                                                      \
            set_error(std::string(#expr) + ": " + (api)->GetErrorString(_r));                         \
            return CDB_NCCL_ERROR;                                                                    \
        }                                                                                             \
    } while (0)

struct Buf {
    void *p = nullptr;
    size_t cap = 0;
    cdb_status ensure(size_t bytes) {
        if (bytes <= cap) return CDB_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        CDB_CUDA_TRY(cudaMalloc(&p, bytes + bytes / 4 + 256));
        cap = bytes + bytes / 4 + 256;
        return CDB_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// per-rank slot of the gather buffer, in u64 words: nq*k keys, then the nq error bytes
__host__ __device__ inline size_t slot_words(uint32_t nq, uint32_t k) { return (size_t)nq * k + ((size_t)nq + 7) / 8; }

// one CTA per query: the best k of world*k keys (rank selection; keys are unique), error bytes OR-ed over the ranks
__global__ void merge_gathered_kernel(const uint64_t *__restrict__ gathered, uint32_t world, uint32_t nq, uint32_t k, size_t words,
                                      int metric, uint32_t *__restrict__ ids, float *__restrict__ scores,
                                      uint32_t *__restrict__ counts, uint8_t *__restrict__ err) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint64_t *keys = reinterpret_cast<uint64_t *>(smem);
    __shared__ int nvalid;
    const uint32_t q = blockIdx.x, M = world * k;
    if (threadIdx.x == 0) nvalid = 0;
    for (uint32_t j = threadIdx.x; j < k; j += blockDim.x) { ids[(size_t)q * k + j] = CDB_INVALID_ID; scores[(size_t)q * k + j] = 0.0f; }
    __syncthreads();
    int local = 0;
    for (uint32_t i = threadIdx.x; i < M; i += blockDim.x) {
        const uint32_t r = i / k, j = i - r * k;
        const uint64_t v = gathered[(size_t)r * words + (size_t)q * k + j];
        keys[i] = v;
        local += v != 0;
    }
    if (local) atomicAdd(&nvalid, local);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < M; i += blockDim.x) {
        const uint64_t key = keys[i];
        if (!key) continue;
        uint32_t rank = 0;
        for (uint32_t j = 0; j < M; ++j) rank += keys[j] > key;
        if (rank < k) {
            ids[(size_t)q * k + rank] = key64_id(key);
            scores[(size_t)q * k + rank] = __uint_as_float(key_to_bits(metric, (uint32_t)(key >> 32)));
        }
    }
    if (threadIdx.x == 0) {
        if (counts) counts[q] = (uint32_t)nvalid < k ? (uint32_t)nvalid : k;
        if (err) {
            uint8_t e = 0;
            for (uint32_t r = 0; r < world; ++r) e |= reinterpret_cast<const uint8_t *>(gathered + (size_t)r * words + (size_t)nq * k)[q];
            err[q] = e;
        }
    }
}
}  // namespace
}  // namespace cdb

using namespace cdb;

struct cdb_shard_group {
    uint32_t world = 1, rank0 = 0;        // ranks in the communicator; global rank of local slot 0
    bool loopback = false;                // local group with a repeated device: gather by device copies (tests on one GPU)
    std::vector<int> dev;                 // per local slot
    std::vector<ncclComm_t> comm;
    std::vector<cdb_index *> shard;
    std::vector<cudaStream_t> stream;
    std::vector<cudaEvent_t> ev;
    std::vector<Buf> gather, q, ids, scores, counts, err;
    std::mutex mu;
};

#define CDB_REQUIRE_G(cond, msg)                                   \
    do {                                                           \
        if (!(cond)) { set_error(msg); return CDB_INVALID_PARAMS; } \
    } while (0)

static cdb_status group_alloc_slots(cdb_shard_group *g, size_t n) {
    g->comm.assign(n, nullptr);
    g->shard.assign(n, nullptr);
    g->stream.assign(n, nullptr);
    g->ev.assign(n, nullptr);
    g->gather.resize(n); g->q.resize(n); g->ids.resize(n); g->scores.resize(n); g->counts.resize(n); g->err.resize(n);
    for (size_t i = 0; i < n; ++i) {
        CDB_CUDA_TRY(cudaSetDevice(g->dev[i]));
        CDB_CUDA_TRY(cudaStreamCreateWithFlags(&g->stream[i], cudaStreamNonBlocking));
        CDB_CUDA_TRY(cudaEventCreateWithFlags(&g->ev[i], cudaEventDisableTiming));
    }
    return CDB_OK;
}

extern "C" {

cdb_status cdb_nccl_unique_id(uint8_t *out_id128) {
    CDB_REQUIRE_G(out_id128, "null argument");
    NcclApi *n = nccl_api();
    if (!n) { set_error("libnccl.so.2 not found"); return CDB_NCCL_ERROR; }
    ncclUniqueId id;
    CDB_NCCL_TRY(n, n->GetUniqueId(&id));
    static_assert(sizeof(id) == CDB_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId size");
    memcpy(out_id128, &id, sizeof(id));
    return CDB_OK;
}

cdb_status cdb_shard_group_create(const int32_t *device_ordinals, uint32_t n_devices, cdb_shard_group **out) {
    CDB_REQUIRE_G(device_ordinals && out && n_devices >= 1 && n_devices <= 64, "bad argument");
    std::unique_ptr<cdb_shard_group> g(new cdb_shard_group());
    g->world = n_devices;
    g->dev.assign(device_ordinals, device_ordinals + n_devices);
    for (uint32_t i = 0; i < n_devices; ++i)
        for (uint32_t j = 0; j < i; ++j) g->loopback |= g->dev[i] == g->dev[j];
    cdb_status rc = group_alloc_slots(g.get(), n_devices);
    if (rc) { cdb_shard_group_destroy(g.release()); return rc; }
    if (n_devices > 1 && !g->loopback) {
        NcclApi *n = nccl_api();
        if (!n) { set_error("libnccl.so.2 not found"); cdb_shard_group_destroy(g.release()); return CDB_NCCL_ERROR; }
        ncclResult_t r = n->CommInitAll(g->comm.data(), (int)n_devices, g->dev.data());
        if (r != ncclSuccess) {
            set_error(std::string("ncclCommInitAll: ") + n->GetErrorString(r));
            for (auto &c : g->comm) c = nullptr;
            cdb_shard_group_destroy(g.release());
            return CDB_NCCL_ERROR;
        }
    }
    *out = g.release();
    return CDB_OK;
}

cdb_status cdb_shard_group_create_rank(const uint8_t *id128, uint32_t world, uint32_t rank, int32_t device, cdb_shard_group **out) {
    CDB_REQUIRE_G(out && world >= 1 && rank < world && (id128 || world == 1), "bad argument");
    std::unique_ptr<cdb_shard_group> g(new cdb_shard_group());
    g->world = world;
    g->rank0 = rank;
    g->dev.assign(1, device);
    cdb_status rc = group_alloc_slots(g.get(), 1);
    if (rc) { cdb_shard_group_destroy(g.release()); return rc; }
    if (world > 1) {
        NcclApi *n = nccl_api();
        if (!n) { set_error("libnccl.so.2 not found"); cdb_shard_group_destroy(g.release()); return CDB_NCCL_ERROR; }
        ncclUniqueId id;
        memcpy(&id, id128, sizeof(id));
        ncclResult_t r = n->CommInitRank(&g->comm[0], (int)world, id, (int)rank);
        if (r != ncclSuccess) {
            set_error(std::string("ncclCommInitRank: ") + n->GetErrorString(r));
            g->comm[0] = nullptr;
            cdb_shard_group_destroy(g.release());
            return CDB_NCCL_ERROR;
        }
    }
    *out = g.release();
    return CDB_OK;
}

cdb_status cdb_shard_group_destroy(cdb_shard_group *g) {
    if (!g) return CDB_OK;
    NcclApi *n = nccl_api();
    for (size_t i = 0; i < g->dev.size(); ++i) {
        cudaSetDevice(g->dev[i]);
        if (i < g->stream.size() && g->stream[i]) cudaStreamSynchronize(g->stream[i]);
        if (i < g->comm.size() && g->comm[i] && n) n->CommDestroy(g->comm[i]);
        for (auto *v : {&g->gather, &g->q, &g->ids, &g->scores, &g->counts, &g->err})
            if (i < v->size()) (*v)[i].release();
        if (i < g->ev.size() && g->ev[i]) cudaEventDestroy(g->ev[i]);
        if (i < g->stream.size() && g->stream[i]) cudaStreamDestroy(g->stream[i]);
    }
    delete g;
    return CDB_OK;
}

cdb_status cdb_shard_group_attach(cdb_shard_group *g, uint32_t local_slot, cdb_index *shard) {
    CDB_REQUIRE_G(g && shard && local_slot < g->dev.size(), "bad argument");
    CDB_REQUIRE_G(index_device(shard) == g->dev[local_slot], "the shard lives on another device than the group slot");
    CDB_REQUIRE_G(g->shard[0] == nullptr || local_slot == 0 || index_dim(shard) == index_dim(g->shard[0]), "shards differ in dimension");
    g->shard[local_slot] = shard;
    return CDB_OK;
}

uint32_t cdb_shard_group_world(const cdb_shard_group *g) { return g ? g->world : 0; }

// local slot i: search its shard into its slot of its gather buffer, then the collective, then (slot 0 / this rank) the merge
static cdb_status sharded_enqueue(cdb_shard_group *g, uint32_t nq, const cdb_search_params *p, const std::vector<const float *> &d_q,
                                  uint32_t *d_ids, float *d_scores, uint32_t *d_counts, uint8_t *d_err, cudaStream_t out_stream) {
    const size_t nl = g->dev.size();
    const uint32_t k = p->k;
    const size_t words = slot_words(nq, k);
    cdb_status rc;
    for (size_t i = 0; i < nl; ++i) {
        CDB_REQUIRE_G(g->shard[i], "a group slot has no shard attached");
        CDB_CUDA_TRY(cudaSetDevice(g->dev[i]));
        if ((rc = g->gather[i].ensure(g->world * words * 8)) || (rc = g->ids[i].ensure((size_t)nq * k * 4)) ||
            (rc = g->scores[i].ensure((size_t)nq * k * 4)) || (rc = g->counts[i].ensure((size_t)nq * 4)))
            return rc;
        uint64_t *slot = reinterpret_cast<uint64_t *>(g->gather[i].p) + (size_t)(g->rank0 + i) * words;
        cudaStream_t s = (nl == 1 && out_stream) ? out_stream : g->stream[i];
        if ((rc = index_search_device_keys(g->shard[i], d_q[i], nq, p, reinterpret_cast<uint32_t *>(g->ids[i].p),
                                           reinterpret_cast<float *>(g->scores[i].p), reinterpret_cast<uint32_t *>(g->counts[i].p),
                                           reinterpret_cast<uint8_t *>(slot + (size_t)nq * k), slot, s)))
            return rc;
    }
    if (g->world > 1) {
        if (g->loopback) {
            // test mode (a device listed twice): every slot copies its part into slot 0's buffer, ordered by events
            CDB_CUDA_TRY(cudaSetDevice(g->dev[0]));
            for (size_t i = 1; i < nl; ++i) {
                CDB_CUDA_TRY(cudaEventRecord(g->ev[i], g->stream[i]));
                CDB_CUDA_TRY(cudaStreamWaitEvent(g->stream[0], g->ev[i], 0));
                CDB_CUDA_TRY(cudaMemcpyAsync(reinterpret_cast<uint64_t *>(g->gather[0].p) + i * words,
                                             reinterpret_cast<uint64_t *>(g->gather[i].p) + i * words, words * 8, cudaMemcpyDeviceToDevice,
                                             g->stream[0]));
            }
        } else {
            NcclApi *n = nccl_api();
            if (!n) { set_error("libnccl.so.2 not found"); return CDB_NCCL_ERROR; }
            CDB_NCCL_TRY(n, n->GroupStart());
            for (size_t i = 0; i < nl; ++i) {
                uint64_t *base = reinterpret_cast<uint64_t *>(g->gather[i].p);
                cudaStream_t s = (nl == 1 && out_stream) ? out_stream : g->stream[i];
                ncclResult_t r = n->AllGather(base + (size_t)(g->rank0 + i) * words, base, words, ncclUint64, g->comm[i], s);
                if (r != ncclSuccess) { n->GroupEnd(); set_error(std::string("ncclAllGather: ") + n->GetErrorString(r)); return CDB_NCCL_ERROR; }
            }
            CDB_NCCL_TRY(n, n->GroupEnd());
        }
    }
    // merge on local slot 0
    CDB_CUDA_TRY(cudaSetDevice(g->dev[0]));
    cudaStream_t s0 = (nl == 1 && out_stream) ? out_stream : g->stream[0];
    const size_t smem = (size_t)g->world * k * 8;
    if (smem > 200 * 1024) { set_error("sharded merge: world * k too large"); return CDB_INVALID_PARAMS; }
    CDB_ALLOW_SMEM(merge_gathered_kernel, smem);
    merge_gathered_kernel<<<nq, 128, smem, s0>>>(reinterpret_cast<const uint64_t *>(g->gather[0].p), g->world, nq, k, words,
                                                  index_result_metric(g->shard[0], p->mode), d_ids, d_scores, d_counts, d_err);
    CDB_LAUNCH_CHECK();
    return CDB_OK;
}

cdb_status cdb_search_batch_sharded(cdb_shard_group *g, const float *queries, uint32_t nq, const cdb_search_params *p,
                                    uint32_t *out_ids, float *out_scores, uint32_t *out_counts, uint8_t *err_flags) {
    CDB_REQUIRE_G(g && p && (queries || !nq) && (out_ids || !nq) && (out_scores || !nq), "null argument");
    CDB_REQUIRE_G(p->k >= 1 && p->k <= 1024, "k must be in 1..1024");
    if (!nq) return CDB_OK;
    std::lock_guard<std::mutex> lock(g->mu);
    const size_t nl = g->dev.size();
    CDB_REQUIRE_G(g->shard[0], "a group slot has no shard attached");
    const uint32_t dim = index_dim(g->shard[0]);
    const size_t nk = (size_t)nq * p->k;
    cdb_status rc;
    std::vector<const float *> d_q(nl);
    for (size_t i = 0; i < nl; ++i) {   // the same host batch goes to every local device
        CDB_CUDA_TRY(cudaSetDevice(g->dev[i]));
        if ((rc = g->q[i].ensure((size_t)nq * dim * 4))) return rc;
        CDB_CUDA_TRY(cudaMemcpyAsync(g->q[i].p, queries, (size_t)nq * dim * 4, cudaMemcpyHostToDevice, g->stream[i]));
        d_q[i] = reinterpret_cast<const float *>(g->q[i].p);
    }
    CDB_CUDA_TRY(cudaSetDevice(g->dev[0]));
    // final results land in slot 0's ids/scores buffers AFTER the local search wrote them: separate output buffers
    Buf &o_err = g->err[0];
    if ((rc = o_err.ensure(nk * 8 + (size_t)nq * 8))) return rc;
    uint32_t *f_ids = reinterpret_cast<uint32_t *>(o_err.p);
    float *f_scores = reinterpret_cast<float *>(f_ids + nk);
    uint32_t *f_counts = reinterpret_cast<uint32_t *>(f_scores + nk);
    uint8_t *f_err = reinterpret_cast<uint8_t *>(f_counts + nq);
    if ((rc = sharded_enqueue(g, nq, p, d_q, f_ids, f_scores, f_counts, f_err, nullptr))) return rc;
    CDB_CUDA_TRY(cudaSetDevice(g->dev[0]));
    cudaStream_t s0 = g->stream[0];
    CDB_CUDA_TRY(cudaMemcpyAsync(out_ids, f_ids, nk * 4, cudaMemcpyDeviceToHost, s0));
    CDB_CUDA_TRY(cudaMemcpyAsync(out_scores, f_scores, nk * 4, cudaMemcpyDeviceToHost, s0));
    if (out_counts) CDB_CUDA_TRY(cudaMemcpyAsync(out_counts, f_counts, (size_t)nq * 4, cudaMemcpyDeviceToHost, s0));
    if (err_flags) CDB_CUDA_TRY(cudaMemcpyAsync(err_flags, f_err, nq, cudaMemcpyDeviceToHost, s0));
    for (size_t i = 0; i < nl; ++i) {   // every local stream drains (collective included) before the host buffers are reused
        CDB_CUDA_TRY(cudaSetDevice(g->dev[i]));
        CDB_CUDA_TRY(cudaStreamSynchronize(g->stream[i]));
    }
    return CDB_OK;
}

cdb_status cdb_search_batch_sharded_device(cdb_shard_group *g, const float *d_queries, uint32_t nq, const cdb_search_params *p,
                                           uint32_t *d_out_ids, float *d_out_scores, uint32_t *d_out_counts, uint8_t *d_err_flags,
                                           void *stream) {
    CDB_REQUIRE_G(g && p && (d_queries || !nq) && (d_out_ids || !nq) && (d_out_scores || !nq), "null argument");
    CDB_REQUIRE_G(g->dev.size() == 1, "the device form serves one-process-per-GPU groups (cdb_shard_group_create_rank)");
    CDB_REQUIRE_G(p->k >= 1 && p->k <= 1024, "k must be in 1..1024");
    if (!nq) return CDB_OK;
    std::lock_guard<std::mutex> lock(g->mu);
    std::vector<const float *> d_q(1, d_queries);
    return sharded_enqueue(g, nq, p, d_q, d_out_ids, d_out_scores, d_out_counts, d_err_flags, (cudaStream_t)stream);
}

}  // extern "C"
