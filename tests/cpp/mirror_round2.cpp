// Compile-and-link check of the round-2 part of the C++ host mirror (include/cosdata_b200.hpp): bf16 storage id, GPU-side
// graph builds, raw-row fill, ShardGroup.  Without a device every call must fail loudly (CDB_CUDA_ERROR), never compute on the CPU.
#include <cstdio>
#include <vector>

#include "cosdata_b200.hpp"

int main() {
    using namespace cosdata;
    static_assert((int)StorageType::BFloat16 == CDB_ST_BF16, "bf16 storage id");
    int32_t ndev = 0;
    cdb_device_count(&ndev);
    std::printf("devices=%d\n", ndev);
    int loud = 0;
    try {
        ShardGroup g(std::vector<int32_t>{0, 0});            // loopback form; needs a device
        DenseIndex a(16, StorageType::BFloat16, DistanceMetricKind::Cosine, 64, 0, {-1.f, 1.f}, true), b(16, StorageType::BFloat16,
                     DistanceMetricKind::Cosine, 64, 0, {-1.f, 1.f}, true, 32);
        std::vector<float> v(32 * 16);
        for (size_t i = 0; i < v.size(); ++i) v[i] = (float)((i * 37) % 101) / 101.0f - 0.5f;
        a.append(v.data(), 32);
        b.append(v.data(), 32);
        g.attach(0, a);
        g.attach(1, b);
        SearchResults r = g.batch_search(v.data(), 2, 3);
        std::printf("world=%u top=%u,%u\n", g.world(), r.ids[0], r.ids[1]);
        if (r.ids[0] != 0 || r.ids[1] != 32) return 2;       // row 0 of both shards is the query itself: ids 0 and 32 (tie -> smaller id first)
        a.build_graph(2, 4, 8, 8, 64, 8, 1);
        std::printf("size after build=%llu raw_missing=%llu\n", (unsigned long long)a.size(), (unsigned long long)a.raw_missing());
    } catch (const Error &e) {
        loud = 1;
        std::printf("error: %s\n", e.what());
    }
    if (ndev == 0 && !loud) return 3;                         // no device and no error: a silent fallback
    if (ndev > 0 && loud) return 4;
    return 0;
}
