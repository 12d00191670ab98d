// "Next" row N3: the on-disk format either side of the path.
//
// The reference stores its data sets as GZIP-compressed TFRecord files holding one tf.train.SequenceExample per
// record (writer: convert_data.py:247-279; parser: dataloader/outdoor_data_mfcc.py:260-344,
// dataloader/frames.py:246-341): int64 context features ('classes', 'location', 'audio_image/height', ...; for
// FlickrSoundNet also the 'xmin/xmax/ymin/ymax' boxes) and byte-string feature lists ('audio/image' = raw float32
// [H, W, D] per step, 'audio/data' = raw int32 samples, 'video/image' = raw uint8 frames).
//
// This is a dependency-free reader for exactly that: zlib for the container, CRC-32C (Castagnoli, masked as
// TFRecord does) for the framing, and a minimal protobuf wire-format walk for the message.  Host side only - it feeds
// the device path, it is not part of it.
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/aig.h"

namespace {

thread_local std::string g_reader_error;

// ---- CRC-32C ----------------------------------------------------------------------------------------
struct Crc32cTable {
    uint32_t t[256];
    Crc32cTable() {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ 0x82F63B78u : (c >> 1);
            t[i] = c;
        }
    }
};
#if defined(__x86_64__) && defined(__GNUC__)
// SSE4.2 has the Castagnoli polynomial in hardware: 8 bytes per instruction instead of one table look-up per byte
// (a 10 MB record: 2 ms instead of 45 ms).
__attribute__((target("sse4.2"))) uint32_t crc32c_sse42(const uint8_t* p, size_t n) {
    uint64_t c = 0xFFFFFFFFu;
    while (n > 0 && (reinterpret_cast<uintptr_t>(p) & 7u)) { c = __builtin_ia32_crc32qi(static_cast<uint32_t>(c), *p++); --n; }
    for (; n >= 8; n -= 8, p += 8) {
        uint64_t v;
        std::memcpy(&v, p, 8);
        c = __builtin_ia32_crc32di(c, v);
    }
    for (; n > 0; --n) c = __builtin_ia32_crc32qi(static_cast<uint32_t>(c), *p++);
    return static_cast<uint32_t>(c) ^ 0xFFFFFFFFu;
}
const bool g_has_sse42 = __builtin_cpu_supports("sse4.2");
#else
uint32_t crc32c_sse42(const uint8_t*, size_t) { return 0; }
const bool g_has_sse42 = false;
#endif

uint32_t crc32c_table(const uint8_t* p, size_t n) {
    static const Crc32cTable table;
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; ++i) c = table.t[(c ^ p[i]) & 0xFFu] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}
uint32_t crc32c(const uint8_t* p, size_t n) { return g_has_sse42 ? crc32c_sse42(p, n) : crc32c_table(p, n); }
uint32_t masked_crc(const uint8_t* p, size_t n) {
    const uint32_t c = crc32c(p, n);
    return ((c >> 15) | (c << 17)) + 0xA282EAD8u;
}

// ---- protobuf wire format ---------------------------------------------------------------------------
struct Span { const uint8_t* p = nullptr; size_t n = 0; };

bool read_varint(const uint8_t*& p, const uint8_t* end, uint64_t* out) {
    uint64_t v = 0;
    for (int shift = 0; shift < 64 && p < end; shift += 7) {
        const uint8_t b = *p++;
        v |= static_cast<uint64_t>(b & 0x7F) << shift;
        if (!(b & 0x80)) { *out = v; return true; }
    }
    return false;
}
// Visit the fields of one message; fn(field, wire_type, payload span / varint) returns false to stop.
template <typename Fn>
bool walk(Span msg, Fn fn) {
    const uint8_t* p = msg.p;
    const uint8_t* end = msg.p + msg.n;
    while (p < end) {
        uint64_t key;
        if (!read_varint(p, end, &key)) return false;
        const uint32_t field = static_cast<uint32_t>(key >> 3), wire = static_cast<uint32_t>(key & 7);
        Span payload;
        uint64_t value = 0;
        switch (wire) {
            case 0: if (!read_varint(p, end, &value)) return false; break;
            case 1: if (end - p < 8) return false; payload = {p, 8}; p += 8; break;
            case 2: {
                uint64_t len;
                if (!read_varint(p, end, &len) || len > static_cast<uint64_t>(end - p)) return false;
                payload = {p, static_cast<size_t>(len)};
                p += len;
                break;
            }
            case 5: if (end - p < 4) return false; payload = {p, 4}; p += 4; break;
            default: return false;
        }
        if (!fn(field, wire, payload, value)) return true;
    }
    return true;
}

// map<string, V> entry with the given key inside `container` (field 1 = repeated map entries): returns V's bytes.
bool find_map_value(Span container, const char* key, Span* out) {
    bool found = false;
    const size_t klen = std::strlen(key);
    const bool ok = walk(container, [&](uint32_t field, uint32_t wire, Span entry, uint64_t) {
        if (field != 1 || wire != 2) return true;
        Span k, v;
        walk(entry, [&](uint32_t f, uint32_t w, Span s, uint64_t) {
            if (f == 1 && w == 2) k = s;
            if (f == 2 && w == 2) v = s;
            return true;
        });
        if (k.n == klen && std::memcmp(k.p, key, klen) == 0) { *out = v; found = true; return false; }
        return true;
    });
    return ok && found;
}

struct Record { size_t offset, length; };

}  // namespace

struct aig_record_reader {
    std::vector<uint8_t> data;        // inflated file
    std::vector<Record> records;
    Span context(int r) const { return part(r, 1); }
    Span feature_lists(int r) const { return part(r, 2); }
    Span part(int r, uint32_t want) const {
        Span out;
        walk({data.data() + records[r].offset, records[r].length}, [&](uint32_t f, uint32_t w, Span s, uint64_t) {
            if (f == want && w == 2) { out = s; return false; }
            return true;
        });
        return out;
    }
};

namespace {

int reader_fail(int code, const std::string& msg) {
    g_reader_error = msg;
    return code;
}

bool check_record(const aig_record_reader* r, int record) {
    return r != nullptr && record >= 0 && static_cast<size_t>(record) < r->records.size();
}

}  // namespace

extern "C" {

const char* aig_records_last_error(void) { return g_reader_error.c_str(); }

uint32_t aig_crc32c(const void* data, size_t n, int force_table) {
    const uint8_t* p = static_cast<const uint8_t*>(data);
    if (p == nullptr || n == 0) return 0;
    return force_table ? crc32c_table(p, n) : crc32c(p, n);
}

int aig_records_open(const char* path, aig_record_reader** out) {
    if (path == nullptr || out == nullptr) return reader_fail(AIG_ERR_ARGUMENT, "aig_records_open: null argument");
    *out = nullptr;
    gzFile f = gzopen(path, "rb");                      // transparently reads plain (uncompressed) files too
    if (f == nullptr) return reader_fail(AIG_ERR_ARGUMENT, std::string("aig_records_open: cannot open ") + path);
    aig_record_reader* r = new aig_record_reader();
    std::vector<uint8_t> chunk(1 << 20);
    for (;;) {
        const int got = gzread(f, chunk.data(), static_cast<unsigned>(chunk.size()));
        if (got < 0) {
            gzclose(f);
            delete r;
            return reader_fail(AIG_ERR_ARGUMENT, std::string("aig_records_open: corrupt gzip stream in ") + path);
        }
        if (got == 0) break;
        r->data.insert(r->data.end(), chunk.begin(), chunk.begin() + got);
    }
    gzclose(f);
    // TFRecord framing: u64 length | u32 masked crc32c(length) | data | u32 masked crc32c(data), little endian
    size_t pos = 0;
    const size_t n = r->data.size();
    while (pos < n) {
        if (n - pos < 12) { delete r; return reader_fail(AIG_ERR_ARGUMENT, "aig_records_open: truncated record header"); }
        uint64_t len;
        uint32_t crc;
        std::memcpy(&len, r->data.data() + pos, 8);
        std::memcpy(&crc, r->data.data() + pos + 8, 4);
        if (crc != masked_crc(r->data.data() + pos, 8)) { delete r; return reader_fail(AIG_ERR_ARGUMENT, "aig_records_open: length checksum mismatch"); }
        if (len > n - pos - 12 || n - pos - 12 - len < 4) { delete r; return reader_fail(AIG_ERR_ARGUMENT, "aig_records_open: truncated record body"); }
        std::memcpy(&crc, r->data.data() + pos + 12 + len, 4);
        if (crc != masked_crc(r->data.data() + pos + 12, static_cast<size_t>(len))) { delete r; return reader_fail(AIG_ERR_ARGUMENT, "aig_records_open: data checksum mismatch"); }
        r->records.push_back({pos + 12, static_cast<size_t>(len)});
        pos += 12 + len + 4;
    }
    *out = r;
    return AIG_OK;
}

int aig_records_close(aig_record_reader* r) {
    delete r;
    return AIG_OK;
}

int64_t aig_records_count(const aig_record_reader* r) { return r ? static_cast<int64_t>(r->records.size()) : -1; }

// Context feature `key` of record `record` as int64 values (Int64List, packed or not).  Writes up to `capacity`
// values and returns the number present via *count_out.
int aig_record_context_int64(const aig_record_reader* r, int record, const char* key, int64_t* values_out,
                             int capacity, int* count_out) {
    if (!check_record(r, record) || key == nullptr || count_out == nullptr || capacity < 0)
        return reader_fail(AIG_ERR_ARGUMENT, "aig_record_context_int64: bad argument");
    Span feature;
    if (!find_map_value(r->context(record), key, &feature))
        return reader_fail(AIG_ERR_ARGUMENT, std::string("context feature not found: ") + key);
    Span list;
    bool is_int = false;
    walk(feature, [&](uint32_t f, uint32_t w, Span s, uint64_t) {
        if (f == 3 && w == 2) { list = s; is_int = true; return false; }
        return true;
    });
    if (!is_int) return reader_fail(AIG_ERR_ARGUMENT, std::string("context feature is not an int64 list: ") + key);
    int count = 0;
    walk(list, [&](uint32_t f, uint32_t w, Span s, uint64_t v) {
        if (f != 1) return true;
        if (w == 0) {
            if (count < capacity && values_out) values_out[count] = static_cast<int64_t>(v);
            ++count;
        } else if (w == 2) {                                   // packed
            const uint8_t* p = s.p;
            uint64_t x;
            while (p < s.p + s.n && read_varint(p, s.p + s.n, &x)) {
                if (count < capacity && values_out) values_out[count] = static_cast<int64_t>(x);
                ++count;
            }
        }
        return true;
    });
    *count_out = count;
    return AIG_OK;
}

// Feature list `key` (byte strings, one per step): number of steps and total payload bytes.
int aig_record_sequence_size(const aig_record_reader* r, int record, const char* key, int64_t* steps_out,
                             int64_t* bytes_out) {
    if (!check_record(r, record) || key == nullptr || steps_out == nullptr || bytes_out == nullptr)
        return reader_fail(AIG_ERR_ARGUMENT, "aig_record_sequence_size: bad argument");
    Span list;
    if (!find_map_value(r->feature_lists(record), key, &list))
        return reader_fail(AIG_ERR_ARGUMENT, std::string("feature list not found: ") + key);
    int64_t steps = 0, bytes = 0;
    walk(list, [&](uint32_t f, uint32_t w, Span feature, uint64_t) {
        if (f != 1 || w != 2) return true;
        ++steps;
        walk(feature, [&](uint32_t ff, uint32_t ww, Span bl, uint64_t) {
            if (ff == 1 && ww == 2)
                walk(bl, [&](uint32_t f3, uint32_t w3, Span value, uint64_t) {
                    if (f3 == 1 && w3 == 2) bytes += static_cast<int64_t>(value.n);
                    return true;
                });
            return true;
        });
        return true;
    });
    *steps_out = steps;
    *bytes_out = bytes;
    return AIG_OK;
}

// Concatenate the byte strings of all steps of feature list `key` into dst (host memory, dst_bytes >= total).
int aig_record_sequence_read(const aig_record_reader* r, int record, const char* key, void* dst, int64_t dst_bytes) {
    int64_t steps = 0, bytes = 0;
    int rc = aig_record_sequence_size(r, record, key, &steps, &bytes);
    if (rc != AIG_OK) return rc;
    if (dst == nullptr || dst_bytes < bytes) return reader_fail(AIG_ERR_ARGUMENT, "aig_record_sequence_read: destination too small");
    Span list;
    find_map_value(r->feature_lists(record), key, &list);
    uint8_t* out = static_cast<uint8_t*>(dst);
    walk(list, [&](uint32_t f, uint32_t w, Span feature, uint64_t) {
        if (f != 1 || w != 2) return true;
        walk(feature, [&](uint32_t ff, uint32_t ww, Span bl, uint64_t) {
            if (ff == 1 && ww == 2)
                walk(bl, [&](uint32_t f3, uint32_t w3, Span value, uint64_t) {
                    if (f3 == 1 && w3 == 2) { std::memcpy(out, value.p, value.n); out += value.n; }
                    return true;
                });
            return true;
        });
        return true;
    });
    return AIG_OK;
}

}  // extern "C"
