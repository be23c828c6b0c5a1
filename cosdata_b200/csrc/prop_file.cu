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
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstring>
#include <string>
#include <vector>

#include "kernels.h"

namespace cdb {
namespace {

struct Cur {
    const uint8_t *p, *end;
    bool ok = true;
    const char *why = "";
    bool fail(const char *w) { if (ok) { ok = false; why = w; } return false; }
    bool need(size_t n) { return (size_t)(end - p) >= n ? true : fail("truncated record"); }
};

struct Head { uint8_t major, info; uint64_t arg; bool indefinite; };

bool read_head(Cur &c, Head &h) {
    if (!c.need(1)) return false;
    const uint8_t b = *c.p++;
    h.major = b >> 5; h.info = b & 31; h.arg = 0; h.indefinite = false;
    if (h.info < 24) { h.arg = h.info; return true; }
    if (h.info == 31) { h.indefinite = true; return true; }
    if (h.info > 27) return c.fail("reserved CBOR additional info");
    const int nb = 1 << (h.info - 24);
    if (!c.need(nb)) return false;
    for (int i = 0; i < nb; ++i) h.arg = (h.arg << 8) | *c.p++;
    return true;
}
bool at_break(Cur &c) { return c.need(1) && *c.p == 0xFF; }

float half_to_float(uint16_t h) {   // IEEE binary16 -> binary32, exact
    const uint32_t s = (uint32_t)(h >> 15) << 31, e = (h >> 10) & 31, m = h & 1023;
    uint32_t bits;
    if (e == 0) {
        if (m == 0) bits = s;
        else { int sh = 0; uint32_t mm = m; while (!(mm & 1024)) { mm <<= 1; ++sh; } bits = s | ((uint32_t)(113 - sh) << 23) | ((mm & 1023) << 13); }
    } else if (e == 31) bits = s | 0x7F800000u | (m << 13);
    else bits = s | ((e + 112) << 23) | (m << 13);
    float f; memcpy(&f, &bits, 4); return f;
}

bool skip_item(Cur &c, int depth = 0) {
    if (depth > 64) return c.fail("CBOR nesting too deep");
    Head h;
    if (!read_head(c, h)) return false;
    switch (h.major) {
    case 0: case 1: return !h.indefinite || c.fail("bad integer");
    case 2: case 3:
        if (h.indefinite) { while (c.ok && !at_break(c)) skip_item(c, depth + 1); if (c.ok) c.p++; return c.ok; }
        if (!c.need(h.arg)) return false;
        c.p += h.arg; return true;
    case 4: case 5: {
        const uint64_t per = h.major == 5 ? 2 : 1;
        if (h.indefinite) { while (c.ok && !at_break(c)) skip_item(c, depth + 1); if (c.ok) c.p++; return c.ok; }
        if (h.arg > (uint64_t)(c.end - c.p)) return c.fail("truncated record");
        for (uint64_t i = 0; i < h.arg * per && c.ok; ++i) skip_item(c, depth + 1);
        return c.ok;
    }
    case 6: return skip_item(c, depth + 1);
    default: return !h.indefinite || c.fail("unexpected break");   // simple values / floats: argument already consumed
    }
}

bool read_uint(Cur &c, uint64_t &v) {
    Head h;
    if (!read_head(c, h)) return false;
    if (h.major != 0 || h.indefinite) return c.fail("expected unsigned integer");
    v = h.arg; return true;
}
bool read_float(Cur &c, float &f) {
    Head h;
    if (!read_head(c, h)) return false;
    if (h.major == 7 && h.info == 25) { f = half_to_float((uint16_t)h.arg); return true; }
    if (h.major == 7 && h.info == 26) { const uint32_t b = (uint32_t)h.arg; memcpy(&f, &b, 4); return true; }
    if (h.major == 7 && h.info == 27) { double d; memcpy(&d, &h.arg, 8); f = (float)d; return true; }
    if (h.major == 0) { f = (float)h.arg; return true; }
    if (h.major == 1) { f = -1.0f - (float)h.arg; return true; }
    return c.fail("expected float");
}
bool read_text(Cur &c, std::string &s) {
    Head h;
    if (!read_head(c, h)) return false;
    if (h.major != 3 || h.indefinite) return c.fail("expected text key");
    if (!c.need(h.arg)) return false;
    s.assign(reinterpret_cast<const char *>(c.p), h.arg);
    c.p += h.arg; return true;
}
// iterate a map/array: returns false on error; `n` = definite count or UINT64_MAX for indefinite
bool open_container(Cur &c, uint8_t major, uint64_t &n) {
    Head h;
    if (!read_head(c, h)) return false;
    if (h.major != major) return c.fail(major == 5 ? "expected map" : "expected array");
    n = h.indefinite ? UINT64_MAX : h.arg;
    if (!h.indefinite && h.arg > (uint64_t)(c.end - c.p)) return c.fail("truncated record");
    return true;
}
bool next_entry(Cur &c, uint64_t &n) {   // true while another entry follows
    if (!c.ok) return false;
    if (n == UINT64_MAX) { if (at_break(c)) { c.p++; return false; } return c.ok; }
    if (n == 0) return false;
    --n; return true;
}

// Vec<u8> (array of small ints, or a byte string) appended to `out`
bool read_u8_vec(Cur &c, std::vector<uint8_t> &out) {
    if (!c.need(1)) return false;
    if ((*c.p >> 5) == 2) {
        Head h; read_head(c, h);
        if (h.indefinite) return c.fail("indefinite byte string");
        if (!c.need(h.arg)) return false;
        out.insert(out.end(), c.p, c.p + h.arg); c.p += h.arg; return true;
    }
    uint64_t n;
    if (!open_container(c, 4, n)) return false;
    while (next_entry(c, n)) {
        // fast path: value < 24 is the byte itself, 0x18 xx is one more byte
        if (c.p < c.end && *c.p < 24) { out.push_back(*c.p++); continue; }
        uint64_t v;
        if (!read_uint(c, v)) return false;
        if (v > 255) return c.fail("u8 element out of range");
        out.push_back((uint8_t)v);
    }
    return c.ok;
}

struct Record {
    uint32_t id = 0;
    int st = -1;              // cdb_storage_type
    float mag = 0.f;
    uint32_t elems = 0;       // u8/f16/f32: elements; sub-byte: bytes per plane
    std::vector<uint8_t> code;  // tight ABI layout
};

bool parse_storage_fields(Cur &c, const std::string &variant, Record &r) {
    int kind;   // 0 u8, 1 sub, 2 f16, 3 f32
    if (variant == "UnsignedByte") kind = 0;
    else if (variant == "SubByte") kind = 1;
    else if (variant == "HalfPrecisionFP") kind = 2;
    else if (variant == "FullPrecisionFP") kind = 3;
    else return c.fail("unknown Storage variant");
    uint64_t n, resolution = 0, planes = 0;
    bool have_mag = false, have_vec = false;
    if (!open_container(c, 5, n)) return false;
    std::string key;
    while (next_entry(c, n)) {
        if (!read_text(c, key)) return false;
        if (key == "mag") { if (!read_float(c, r.mag)) return false; have_mag = true; }
        else if (key == "resolution" && kind == 1) { if (!read_uint(c, resolution)) return false; }
        else if ((key == "quant_vec" && kind != 3) || (key == "vec" && kind == 3)) {
            have_vec = true;
            if (kind == 0) { if (!read_u8_vec(c, r.code)) return false; r.elems = (uint32_t)r.code.size(); }
            else if (kind == 1) {
                uint64_t np;
                if (!open_container(c, 4, np)) return false;
                size_t plane_len = 0;
                while (next_entry(c, np)) {
                    const size_t before = r.code.size();
                    if (!read_u8_vec(c, r.code)) return false;
                    if (planes && r.code.size() - before != plane_len) return c.fail("SubByte planes differ in length");
                    plane_len = r.code.size() - before;
                    ++planes;
                }
                r.elems = (uint32_t)plane_len;
            } else {
                uint64_t ne;
                if (!open_container(c, 4, ne)) return false;
                while (next_entry(c, ne)) {
                    if (kind == 2) {     // half::f16 serializes as its u16 bit pattern
                        uint64_t v;
                        if (!read_uint(c, v)) return false;
                        if (v > 0xFFFF) return c.fail("f16 bits out of range");
                        r.code.push_back((uint8_t)(v & 255)); r.code.push_back((uint8_t)(v >> 8));
                    } else {
                        float f;
                        if (!read_float(c, f)) return false;
                        uint8_t b[4]; memcpy(b, &f, 4);
                        r.code.insert(r.code.end(), b, b + 4);
                    }
                    r.elems++;
                }
            }
        } else if (!skip_item(c)) return false;
    }
    if (!c.ok) return false;
    if (!have_mag || !have_vec) return c.fail("Storage record misses mag or vector");
    if (kind == 1) {
        if (resolution < 1 || resolution > 3 || planes != resolution) return c.fail("SubByte resolution / plane count mismatch");
        r.st = CDB_ST_SUB1 + (int)resolution - 1;
    } else r.st = kind == 0 ? CDB_ST_U8 : kind == 2 ? CDB_ST_F16 : CDB_ST_F32;
    return true;
}

bool parse_record(Cur &c, Record &r) {
    r = Record();
    uint64_t n;
    if (!open_container(c, 5, n)) return false;
    bool have_id = false, have_value = false;
    std::string key, variant;
    while (next_entry(c, n)) {
        if (!read_text(c, key)) return false;
        if (key == "id") { uint64_t v; if (!read_uint(c, v)) return false; if (v > 0xFFFFFFFFull) return c.fail("id out of range"); r.id = (uint32_t)v; have_id = true; }
        else if (key == "value") {
            uint64_t one;
            if (!open_container(c, 5, one)) return false;
            if (!next_entry(c, one) || !read_text(c, variant) || !parse_storage_fields(c, variant, r)) return c.fail(c.ok ? "empty enum map" : c.why);
            if (next_entry(c, one)) return c.fail("enum map with more than one variant");
            have_value = true;
        } else if (!skip_item(c)) return false;
    }
    if (!c.ok) return false;
    return (have_id && have_value) || c.fail("record misses id or value");
}

struct Mapped {
    const uint8_t *p = nullptr;
    size_t len = 0;
    int fd = -1;
    cdb_status open(const char *path) {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) { set_error(std::string("cannot open ") + path); return CDB_INVALID_PARAMS; }
        struct stat st;
        if (fstat(fd, &st) != 0) { set_error("fstat failed"); return CDB_INVALID_PARAMS; }
        len = (size_t)st.st_size;
        if (len) {
            void *m = mmap(nullptr, len, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m == MAP_FAILED) { set_error("mmap failed"); return CDB_INVALID_PARAMS; }
            p = static_cast<const uint8_t *>(m);
        }
        return CDB_OK;
    }
    ~Mapped() { if (p) munmap(const_cast<uint8_t *>(p), len); if (fd >= 0) ::close(fd); }
};

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
    const uint64_t CH = 16384;
    std::vector<uint8_t> codes;
    std::vector<float> mags;
    codes.reserve(CH * want); mags.reserve(CH);
    Cur c{f.p, f.p + f.len};
    Record r;
    uint64_t idx = 0;
    auto flush = [&]() -> cdb_status {
        if (mags.empty()) return CDB_OK;
        cdb_status e = cdb_index_append_codes(index, codes.data(), mags.data(), mags.size());
        codes.clear(); mags.clear();
        return e;
    };
    while (c.p < c.end) {
        const size_t off = (size_t)(c.p - f.p);
        if (!parse_record(c, r)) { flush(); return record_error(c, idx, off); }
        if (r.st != st_index || r.code.size() != want) {
            flush();
            set_error("prop file record " + std::to_string(idx) + ": Storage variant / length does not match the index (StorageMismatch)");
            return CDB_STORAGE_MISMATCH;
        }
        codes.insert(codes.end(), r.code.begin(), r.code.end());
        mags.push_back(r.mag);
        if (out_ids && idx < max_ids) out_ids[idx] = r.id;
        ++idx;
        if (mags.size() == CH && (rc = flush())) return rc;
    }
    if ((rc = flush())) return rc;
    if (out_appended) *out_appended = idx;
    return CDB_OK;
}

}  // extern "C"
