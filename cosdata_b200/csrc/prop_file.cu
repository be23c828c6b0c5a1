// prop_file.cu -- host-side reader for the reference's `prop.data` (no device code here).
//
// write_prop_value_to_file (src/models/file_persist.rs:58-87) appends one serde_cbor record per node:
//     NodePropValueSerialize { id: &InternalId, value: &Storage }                      (file_persist.rs:58-62)
// and ProbNode keeps (FileOffset, BytesToRead) of its record (serializer/hnsw/node.rs:51-54).  The file therefore is a
// plain concatenation of CBOR items and can be walked front to back without the index files.
//
// serde_cbor = 0.11.2 (Cargo.toml:28; not vendored under /root/reference) encodes, per its published format (RFC 8949
// subset): structs as definite-length maps keyed by field-name text strings; a struct enum variant as a 1-entry map
// { "Variant": { fields } }; newtype structs (InternalId(u32), half::f16(u16) with the `serde` feature, Cargo.toml:17)
// as their inner value; Vec<T> as a definite-length array (Vec<u8> is an array of unsigned ints, not a byte string);
// unsigned ints in the shortest form; f32 as half precision (0xf9) when that is lossless, else 0xfa.  The decoder below
// accepts every width / definite and indefinite lengths / byte strings for Vec<u8>, so it does not depend on those choices.
#include "ref_files.h"

namespace cdb {
namespace {
using namespace reffiles;

cdb_status record_error(const Cur &c, uint64_t index, size_t offset) {
    set_error("prop file record " + std::to_string(index) + " at byte " + std::to_string(offset) + ": " + c.why);
    return CDB_INVALID_PARAMS;
}

}  // namespace
}  // namespace cdb

using namespace cdb;

extern "C" {

cdb_status cdb_prop_file_scan(const char *path, uint64_t *out_records, int32_t *out_storage_type, uint32_t *out_elems,
                              uint64_t *out_code_bytes) {
    if (!path) { set_error("null path"); return CDB_INVALID_PARAMS; }
    Mapped f;
    cdb_status rc = f.open(path);
    if (rc) return rc;
    Cur c{f.p, f.p + f.len};
    uint64_t n = 0;
    Record r, first;
    while (c.p < c.end) {
        const size_t off = (size_t)(c.p - f.p);
        if (peek_is_metadata(c)) { if (!skip_item(c)) return record_error(c, n, off); continue; }   // replica Metadata: not a row
        if (!parse_record(c, r)) return record_error(c, n, off);
        if (n == 0) first = r;
        else if (r.st != first.st || r.code.size() != first.code.size()) {
            set_error("prop file record " + std::to_string(n) + ": storage variant or length differs from record 0");
            return CDB_STORAGE_MISMATCH;
        }
        ++n;
    }
    if (out_records) *out_records = n;
    if (out_storage_type) *out_storage_type = n ? first.st : -1;
    if (out_elems) *out_elems = n ? first.elems : 0;
    if (out_code_bytes) *out_code_bytes = n ? first.code.size() : 0;
    return CDB_OK;
}

cdb_status cdb_prop_file_load(const char *path, uint64_t first_record, uint64_t max_records, uint32_t *out_ids, void *out_codes,
                              float *out_mags, uint64_t *out_offsets, uint32_t *out_lengths, uint64_t *out_read) {
    if (!path || !out_read) { set_error("null argument"); return CDB_INVALID_PARAMS; }
    Mapped f;
    cdb_status rc = f.open(path);
    if (rc) return rc;
    Cur c{f.p, f.p + f.len};
    uint64_t idx = 0, got = 0;
    Record r;
    int st = -1;
    size_t bytes = 0;
    while (c.p < c.end && got < max_records) {
        const size_t off = (size_t)(c.p - f.p);
        if (peek_is_metadata(c)) { if (!skip_item(c)) return record_error(c, idx, off); continue; }
        if (idx < first_record) { if (!skip_item(c)) return record_error(c, idx, off); ++idx; continue; }
        if (!parse_record(c, r)) return record_error(c, idx, off);
        if (got == 0) { st = r.st; bytes = r.code.size(); }
        else if (r.st != st || r.code.size() != bytes) {
            set_error("prop file record " + std::to_string(idx) + ": storage variant or length differs");
            return CDB_STORAGE_MISMATCH;
        }
        if (out_ids) out_ids[got] = r.id;
        if (out_mags) out_mags[got] = r.mag;
        if (out_codes) memcpy(static_cast<uint8_t *>(out_codes) + got * bytes, r.code.data(), bytes);
        if (out_offsets) out_offsets[got] = off;
        if (out_lengths) out_lengths[got] = (uint32_t)((size_t)(c.p - f.p) - off);
        ++got; ++idx;
    }
    *out_read = got;
    return CDB_OK;
}

cdb_status cdb_prop_file_scan_metadata(const char *path, uint64_t *out_records, uint32_t *out_md_dims) {
    if (!path) { set_error("null path"); return CDB_INVALID_PARAMS; }
    Mapped f;
    cdb_status rc = f.open(path);
    if (rc) return rc;
    Cur c{f.p, f.p + f.len};
    uint64_t n = 0, idx = 0;
    size_t dims = 0;
    Record r;
    while (c.p < c.end) {
        const size_t off = (size_t)(c.p - f.p);
        if (!peek_is_metadata(c)) { if (!skip_item(c)) return record_error(c, idx, off); ++idx; continue; }
        if (!parse_record(c, r)) return record_error(c, idx, off);
        if (n == 0) dims = r.mbits.size();
        else if (r.mbits.size() != dims) { set_error("prop file: Metadata records differ in length"); return CDB_STORAGE_MISMATCH; }
        ++n; ++idx;
    }
    if (out_records) *out_records = n;
    if (out_md_dims) *out_md_dims = (uint32_t)dims;
    return CDB_OK;
}

cdb_status cdb_prop_file_load_metadata(const char *path, uint64_t max_records, uint32_t md_dims, uint32_t *out_replica_ids,
                                       float *out_mags, int32_t *out_mbits, uint64_t *out_offsets, uint32_t *out_lengths,
                                       uint64_t *out_read) {
    if (!path || !out_read) { set_error("null argument"); return CDB_INVALID_PARAMS; }
    Mapped f;
    cdb_status rc = f.open(path);
    if (rc) return rc;
    Cur c{f.p, f.p + f.len};
    uint64_t got = 0, idx = 0;
    Record r;
    while (c.p < c.end && got < max_records) {
        const size_t off = (size_t)(c.p - f.p);
        if (!peek_is_metadata(c)) { if (!skip_item(c)) return record_error(c, idx, off); ++idx; continue; }
        if (!parse_record(c, r)) return record_error(c, idx, off);
        if (r.mbits.size() != md_dims) { set_error("prop file: Metadata record length differs from md_dims"); return CDB_STORAGE_MISMATCH; }
        if (out_replica_ids) out_replica_ids[got] = r.id;
        if (out_mags) out_mags[got] = r.mag;
        if (out_mbits) memcpy(out_mbits + got * md_dims, r.mbits.data(), (size_t)md_dims * 4);
        if (out_offsets) out_offsets[got] = off;
        if (out_lengths) out_lengths[got] = (uint32_t)((size_t)(c.p - f.p) - off);
        ++got; ++idx;
    }
    *out_read = got;
    return CDB_OK;
}

cdb_status cdb_index_append_prop_file(cdb_index *index, const char *path, uint32_t *out_ids, uint64_t max_ids, uint64_t *out_appended) {
    if (!index || !path) { set_error("null argument"); return CDB_INVALID_PARAMS; }
    cdb_index_desc desc;
    cdb_status rc = cdb_index_describe(index, &desc);
    if (rc) return rc;
    const int32_t st_index = desc.storage_type;
    const uint32_t dim = desc.dim;
    Mapped f;
    if ((rc = f.open(path))) return rc;
    const size_t want = cdb_code_bytes(st_index, dim);
    if (out_appended) *out_appended = 0;
    // pass 1: validate the whole file -- nothing is committed to the index unless every record parses and matches it
    // (a half-loaded index would silently break the row = record-ordinal mapping the graph flattener relies on)
    {
        Cur c{f.p, f.p + f.len};
        Record r;
        uint64_t idx = 0;
        while (c.p < c.end) {
            const size_t off = (size_t)(c.p - f.p);
            if (peek_is_metadata(c)) { if (!skip_item(c)) return record_error(c, idx, off); continue; }
            if (!parse_record(c, r)) return record_error(c, idx, off);
            if (r.st != st_index || r.code.size() != want) {
                set_error("prop file record " + std::to_string(idx) + ": Storage variant / length does not match the index (StorageMismatch)");
                return CDB_STORAGE_MISMATCH;
            }
            ++idx;
        }
    }
    // pass 2: append in chunks; *out_appended always reports the rows committed so far (a device error can still stop it)
    const uint64_t CH = 16384;
    std::vector<uint8_t> codes;
    std::vector<float> mags;
    codes.reserve(CH * want); mags.reserve(CH);
    Cur c{f.p, f.p + f.len};
    Record r;
    uint64_t idx = 0, committed = 0;
    auto flush = [&]() -> cdb_status {
        if (mags.empty()) return CDB_OK;
        cdb_status e = cdb_index_append_codes(index, codes.data(), mags.data(), mags.size());
        if (e == CDB_OK) { committed += mags.size(); if (out_appended) *out_appended = committed; }
        codes.clear(); mags.clear();
        return e;
    };
    while (c.p < c.end) {
        if (peek_is_metadata(c)) { skip_item(c); continue; }
        parse_record(c, r);
        codes.insert(codes.end(), r.code.begin(), r.code.end());
        mags.push_back(r.mag);
        if (out_ids && idx < max_ids) out_ids[idx] = r.id;
        ++idx;
        if (mags.size() == CH && (rc = flush())) return rc;
    }
    return flush();
}

}  // extern "C"
