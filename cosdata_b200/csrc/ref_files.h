// ref_files.h -- host-only helpers shared by the readers of the reference's on-disk files (prop_file.cu, index_file.cu):
// a small CBOR decoder for the serde_cbor records of prop.data and a read-only memory map.  No device code.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <climits>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "kernels.h"

namespace cdb {
namespace reffiles {


struct Cur {
    const uint8_t *p, *end;
    bool ok = true;
    const char *why = "";
    bool fail(const char *w) { if (ok) { ok = false; why = w; } return false; }
    bool need(size_t n) { return (size_t)(end - p) >= n ? true : fail("truncated record"); }
};

struct Head { uint8_t major, info; uint64_t arg; bool indefinite; };

inline bool read_head(Cur &c, Head &h) {
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
inline bool at_break(Cur &c) { return c.need(1) && *c.p == 0xFF; }

inline float half_to_float(uint16_t h) {   // IEEE binary16 -> binary32, exact
    const uint32_t s = (uint32_t)(h >> 15) << 31, e = (h >> 10) & 31, m = h & 1023;
    uint32_t bits;
    if (e == 0) {
        if (m == 0) bits = s;
        else { int sh = 0; uint32_t mm = m; while (!(mm & 1024)) { mm <<= 1; ++sh; } bits = s | ((uint32_t)(113 - sh) << 23) | ((mm & 1023) << 13); }
    } else if (e == 31) bits = s | 0x7F800000u | (m << 13);
    else bits = s | ((e + 112) << 23) | (m << 13);
    float f; memcpy(&f, &bits, 4); return f;
}

inline bool skip_item(Cur &c, int depth = 0) {
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

inline bool read_uint(Cur &c, uint64_t &v) {
    Head h;
    if (!read_head(c, h)) return false;
    if (h.major != 0 || h.indefinite) return c.fail("expected unsigned integer");
    v = h.arg; return true;
}
inline bool read_float(Cur &c, float &f) {
    Head h;
    if (!read_head(c, h)) return false;
    if (h.major == 7 && h.info == 25) { f = half_to_float((uint16_t)h.arg); return true; }
    if (h.major == 7 && h.info == 26) { const uint32_t b = (uint32_t)h.arg; memcpy(&f, &b, 4); return true; }
    if (h.major == 7 && h.info == 27) { double d; memcpy(&d, &h.arg, 8); f = (float)d; return true; }
    if (h.major == 0) { f = (float)h.arg; return true; }
    if (h.major == 1) { f = -1.0f - (float)h.arg; return true; }
    return c.fail("expected float");
}
inline bool read_text(Cur &c, std::string &s) {
    Head h;
    if (!read_head(c, h)) return false;
    if (h.major != 3 || h.indefinite) return c.fail("expected text key");
    if (!c.need(h.arg)) return false;
    s.assign(reinterpret_cast<const char *>(c.p), h.arg);
    c.p += h.arg; return true;
}
// iterate a map/array: returns false on error; `n` = definite count or UINT64_MAX for indefinite
inline bool open_container(Cur &c, uint8_t major, uint64_t &n) {
    Head h;
    if (!read_head(c, h)) return false;
    if (h.major != major) return c.fail(major == 5 ? "expected map" : "expected array");
    n = h.indefinite ? UINT64_MAX : h.arg;
    if (!h.indefinite && h.arg > (uint64_t)(c.end - c.p)) return c.fail("truncated record");
    return true;
}
inline bool next_entry(Cur &c, uint64_t &n) {   // true while another entry follows
    if (!c.ok) return false;
    if (n == UINT64_MAX) { if (at_break(c)) { c.p++; return false; } return c.ok; }
    if (n == 0) return false;
    --n; return true;
}

// Vec<u8> (array of small ints, or a byte string) appended to `out`
inline bool read_u8_vec(Cur &c, std::vector<uint8_t> &out) {
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

// one record of prop.data: a node's Storage (write_prop_value_to_file) or a replica's Metadata (write_prop_metadata_to_file,
// file_persist.rs:110-139: { replica_id: InternalId, vec: Metadata { mag: f32, mbits: Vec<i32> } })
struct Record {
    bool is_metadata = false;
    std::vector<int32_t> mbits;   // metadata records
    uint32_t id = 0;              // value records: prop id; metadata records: replica id
    int st = -1;              // cdb_storage_type
    float mag = 0.f;
    uint32_t elems = 0;       // u8/f16/f32: elements; sub-byte: bytes per plane
    std::vector<uint8_t> code;  // tight ABI layout
};

inline bool parse_storage_fields(Cur &c, const std::string &variant, Record &r) {
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

inline bool read_int(Cur &c, int64_t &v) {
    Head h;
    if (!read_head(c, h)) return false;
    if ((h.major != 0 && h.major != 1) || h.indefinite || h.arg > 0x7FFFFFFFFFFFFFFFull) return c.fail("expected integer");
    v = h.major == 0 ? (int64_t)h.arg : -1 - (int64_t)h.arg;
    return true;
}

inline bool parse_metadata_fields(Cur &c, Record &r) {
    uint64_t n;
    if (!open_container(c, 5, n)) return false;
    bool have_mag = false, have_bits = false;
    std::string key;
    while (next_entry(c, n)) {
        if (!read_text(c, key)) return false;
        if (key == "mag") { if (!read_float(c, r.mag)) return false; have_mag = true; }
        else if (key == "mbits") {
            uint64_t ne;
            if (!open_container(c, 4, ne)) return false;
            while (next_entry(c, ne)) {
                int64_t v;
                if (!read_int(c, v)) return false;
                if (v < INT32_MIN || v > INT32_MAX) return c.fail("mbits element out of i32 range");
                r.mbits.push_back((int32_t)v);
            }
            have_bits = true;
        } else if (!skip_item(c)) return false;
    }
    if (!c.ok) return false;
    return (have_mag && have_bits) || c.fail("Metadata record misses mag or mbits");
}

// true when the record at the cursor is a Metadata record (first key replica_id / vec); the cursor is not advanced
inline bool peek_is_metadata(Cur c) {
    uint64_t n;
    std::string key;
    if (!open_container(c, 5, n) || !next_entry(c, n) || !read_text(c, key)) return false;
    return key == "replica_id" || key == "vec";
}

inline bool parse_record(Cur &c, Record &r) {
    r = Record();
    uint64_t n;
    if (!open_container(c, 5, n)) return false;
    bool have_id = false, have_value = false, have_replica = false, have_md = false;
    std::string key, variant;
    while (next_entry(c, n)) {
        if (!read_text(c, key)) return false;
        if (key == "replica_id") { uint64_t v; if (!read_uint(c, v)) return false; if (v > 0xFFFFFFFFull) return c.fail("replica_id out of range"); r.id = (uint32_t)v; have_replica = true; }
        else if (key == "vec") { if (!parse_metadata_fields(c, r)) return false; have_md = true; }
        else if (key == "id") { uint64_t v; if (!read_uint(c, v)) return false; if (v > 0xFFFFFFFFull) return c.fail("id out of range"); r.id = (uint32_t)v; have_id = true; }
        else if (key == "value") {
            uint64_t one;
            if (!open_container(c, 5, one)) return false;
            if (!next_entry(c, one) || !read_text(c, variant) || !parse_storage_fields(c, variant, r)) return c.fail(c.ok ? "empty enum map" : c.why);
            if (next_entry(c, one)) return c.fail("enum map with more than one variant");
            have_value = true;
        } else if (!skip_item(c)) return false;
    }
    if (!c.ok) return false;
    if (have_replica && have_md && !have_id && !have_value) { r.is_metadata = true; return true; }
    return (have_id && have_value && !have_replica && !have_md) || c.fail("record is neither { id, value } nor { replica_id, vec }");
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


}  // namespace reffiles
}  // namespace cdb
