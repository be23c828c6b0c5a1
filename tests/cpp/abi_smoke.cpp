// Compiled and run by tests/test_cpp_host_mirror.py: the C++ host mirror builds against the C ABI, links the
// shared object, and (without a CUDA device) reports CDB_CUDA_ERROR instead of computing anything on the CPU.
#include <cstdio>
#include <vector>

#include "cosdata_b200.hpp"

int main() {
    using namespace cosdata;
    int32_t ndev = 0;
    cdb_device_count(&ndev);
    std::vector<float> v(64, 0.25f), w(64, -0.5f);
    w[3] = 0.75f;
    ScalarQuantization q;
    try {
        Storage a = q.quantize(v, StorageType::FullPrecisionFP), b = q.quantize(w, StorageType::FullPrecisionFP);
        float cs = DistanceMetric{DistanceMetricKind::Cosine}.calculate(a, b);
        std::printf("devices=%d cosine=%.9g\n", ndev, cs);
        try {
            DistanceMetric{DistanceMetricKind::DotProduct}.calculate(a, b);  // no f32 arm: dotproduct.rs:62
            return 3;
        } catch (const DistanceError &e) {
            if (e.status != CDB_STORAGE_MISMATCH) return 4;
        }
        DenseIndex ix(64, StorageType::FullPrecisionFP, DistanceMetricKind::Cosine, 4);
        std::vector<float> rows;
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 64; ++c) rows.push_back((float)((r * 7 + c) % 5) - 2.0f);
        ix.append(rows.data(), 4);
        SearchResults r = ix.batch_search(rows.data() + 64, 1, 2);
        std::printf("top=%u count=%u\n", r.ids[0], r.counts[0]);
        return (r.ids[0] == 1 && r.counts[0] == 2) ? 0 : 5;
    } catch (const Error &e) {
        std::printf("devices=%d error status=%d: %s\n", ndev, e.status, e.what());
        return (ndev == 0 && e.status == CDB_CUDA_ERROR) ? 0 : 2;   // no device -> loud failure, never a CPU result
    }
}
