// Sharded search through the plain C ABI, no torch / no Python: the host program a Rust (or any FFI) caller would write.
//   shard_smoke <n_devices>      n_devices >= 2 and that many GPUs present: one NCCL communicator over devices 0..n-1
//   shard_smoke 0                single GPU: two shards on device 0 (copy-based loopback gather, same code path otherwise)
// Checks that the merged top-k over the shards equals the top-k of one index holding the whole corpus.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "cosdata_b200.h"

#define CHECK(x)                                                                              \
    do {                                                                                      \
        cdb_status _rc = (x);                                                                 \
        if (_rc != CDB_OK) { std::printf("FAILED %s -> %d: %s\n", #x, _rc, cdb_last_error_string()); return 1; } \
    } while (0)

int main(int argc, char **argv) {
    int32_t have = 0;
    if (cdb_device_count(&have) != CDB_OK || have == 0) { std::printf("no device\n"); return 0; }
    int want = argc > 1 ? std::atoi(argv[1]) : 0;
    const uint32_t shards = want >= 2 ? (uint32_t)want : 2;
    if (want >= 2 && have < want) { std::printf("skipped: %d devices present\n", have); return 0; }
    const uint32_t dim = 96, n = 40000, nq = 37, k = 10;
    const uint64_t seed_c = 901, seed_q = 902;
    std::vector<float> q((size_t)nq * dim);
    CHECK(cdb_synth_fill_host(seed_q, 0, q.size(), q.data()));

    cdb_index_desc d{};
    d.dim = dim; d.storage_type = CDB_ST_F32; d.metric = CDB_METRIC_COSINE; d.range_lo = -1.f; d.range_hi = 1.f;
    d.capacity = n; d.device = 0; d.keep_raw_f32 = 0; d.id_base = 0; d.tensor_prefilter = 1;
    cdb_index *whole = nullptr;
    CHECK(cdb_index_create(&d, &whole));
    CHECK(cdb_index_append_synthetic(whole, seed_c, 0, n));
    cdb_search_params p{};
    p.k = k; p.mode = CDB_MODE_BRUTE_RAW; p.ef_search = 64; p.shortlist_size = 64;
    std::vector<uint32_t> want_ids((size_t)nq * k), got_ids((size_t)nq * k), counts(nq);
    std::vector<float> want_sc((size_t)nq * k), got_sc((size_t)nq * k);
    std::vector<uint8_t> err(nq);
    CHECK(cdb_search_batch(whole, q.data(), nq, &p, want_ids.data(), want_sc.data(), counts.data(), err.data()));

    std::vector<int32_t> devs(shards);
    for (uint32_t i = 0; i < shards; ++i) devs[i] = want >= 2 ? (int32_t)i : 0;
    cdb_shard_group *g = nullptr;
    CHECK(cdb_shard_group_create(devs.data(), shards, &g));
    std::vector<cdb_index *> part(shards, nullptr);
    for (uint32_t i = 0; i < shards; ++i) {
        const uint64_t row0 = (uint64_t)n * i / shards, row1 = (uint64_t)n * (i + 1) / shards;
        cdb_index_desc sd = d;
        sd.capacity = row1 - row0; sd.device = devs[i]; sd.id_base = (uint32_t)row0;
        CHECK(cdb_index_create(&sd, &part[i]));
        CHECK(cdb_index_append_synthetic(part[i], seed_c, row0, row1 - row0));
        CHECK(cdb_shard_group_attach(g, i, part[i]));
    }
    for (int rep = 0; rep < 3; ++rep)
        CHECK(cdb_search_batch_sharded(g, q.data(), nq, &p, got_ids.data(), got_sc.data(), counts.data(), err.data()));
    const bool same = std::memcmp(want_ids.data(), got_ids.data(), want_ids.size() * 4) == 0 &&
                      std::memcmp(want_sc.data(), got_sc.data(), want_sc.size() * 4) == 0;
    std::printf("shards=%u world=%u mode=%s merged_equals_whole=%d first=%u\n", shards, cdb_shard_group_world(g),
                want >= 2 ? "nccl" : "loopback", (int)same, got_ids[0]);
    CHECK(cdb_shard_group_destroy(g));
    for (cdb_index *ix : part) CHECK(cdb_index_destroy(ix));
    CHECK(cdb_index_destroy(whole));
    return same ? 0 : 2;
}
