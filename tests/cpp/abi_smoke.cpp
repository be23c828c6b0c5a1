// Compiled and run by tests/test_cpp_host_mirror.py: the C++ host mirror builds against the C ABI, links the
// shared object, and (without a CUDA device) reports CDB_CUDA_ERROR instead of computing anything on the CPU.
#include <cstdio>
#include <cstdlib>
#include <unistd.h>
#include <vector>

#include "cosdata_b200.hpp"

// host-only part of the ABI: one serde_cbor record {"id": 7, "value": {"UnsignedByte": {"mag": 1.5, "quant_vec": [0, 23, 24, 255]}}}
static int check_prop_reader() {
    static const unsigned char rec[] = {0xa2, 0x62, 'i', 'd', 0x07, 0x65, 'v', 'a', 'l', 'u', 'e', 0xa1, 0x6c, 'U', 'n', 's', 'i', 'g', 'n',
                                        'e', 'd', 'B', 'y', 't', 'e', 0xa2, 0x63, 'm', 'a', 'g', 0xf9, 0x3e, 0x00, 0x69, 'q', 'u', 'a',
                                        'n', 't', '_', 'v', 'e', 'c', 0x84, 0x00, 0x17, 0x18, 0x18, 0x18, 0xff};
    char path[] = "/tmp/cdb_abi_smoke_XXXXXX";
    const int fd = mkstemp(path);
    if (fd < 0) return 10;
    const bool ok = write(fd, rec, sizeof(rec)) == (ssize_t)sizeof(rec) && write(fd, rec, sizeof(rec)) == (ssize_t)sizeof(rec);
    close(fd);
    int rc = 11;
    if (ok) {
        try {
            cosdata::PropFile pf(path);
            rc = (pf.records == 2 && pf.storage_type == cosdata::StorageType::UnsignedByte && pf.elems == 4 && pf.ids[1] == 7 &&
                  pf.mags[0] == 1.5f && pf.codes[3] == 255 && pf.codes[6] == 24 && pf.offsets[1] == sizeof(rec) && pf.lengths[0] == sizeof(rec))
                     ? 0 : 12;
        } catch (const cosdata::Error &) { rc = 13; }
    }
    unlink(path);
    return rc;
}

int main() {
    using namespace cosdata;
    if (int rc = check_prop_reader()) return rc;
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
