// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the byte/integer codecs on YugabyteDB's DocDB compaction path.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, link or
// call anything under oracle/.  The product (yugabyte-db_b200/) never includes these headers.
//
// Every function cites the reference file:line it restates (paths relative to
// /root/reference/src/yb/).  Nothing here is copied from the reference; it is rewritten from the
// behaviour read there and pinned by the reference's own golden vectors (tests/test_oracle_*.py) and
// by the reference's debug dumps replayed verbatim (tests/test_reference_dumps.py): parity PINNED for
// the codecs, the merge / CompactionIterator rules and the retention predicate; see DESIGN.md §2 for the
// pieces no reference test fixes byte-wise (metadata file layout, three_shared_parts entry bytes).
#pragma once

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <stdexcept>
#include <algorithm>

namespace orc {

struct Slice {
  const uint8_t* p = nullptr;
  size_t n = 0;
  Slice() {}
  Slice(const uint8_t* p_, size_t n_) : p(p_), n(n_) {}
  Slice(const std::string& s) : p(reinterpret_cast<const uint8_t*>(s.data())), n(s.size()) {}
  bool empty() const { return n == 0; }
  uint8_t operator[](size_t i) const { return p[i]; }
  void remove_prefix(size_t k) { p += k; n -= k; }
  std::string str() const { return std::string(reinterpret_cast<const char*>(p), n); }
  // util/slice.h compare(): memcmp over the common length, then the shorter one is smaller.
  int compare(const Slice& o) const {
    size_t m = n < o.n ? n : o.n;
    int r = m ? memcmp(p, o.p, m) : 0;
    if (r == 0) { if (n < o.n) r = -1; else if (n > o.n) r = 1; }
    return r;
  }
  bool operator==(const Slice& o) const { return n == o.n && (n == 0 || memcmp(p, o.p, n) == 0); }
};

struct Corruption : std::runtime_error { using std::runtime_error::runtime_error; };
struct NotSupported : std::runtime_error { using std::runtime_error::runtime_error; };

// ---------------------------------------------------------------------------------------------
// rocksdb/util/coding.h: little-endian fixed ints and LEB128 varints (:224-253).
inline void PutFixed32(std::string* d, uint32_t v) { char b[4]; memcpy(b, &v, 4); d->append(b, 4); }
inline void PutFixed64(std::string* d, uint64_t v) { char b[8]; memcpy(b, &v, 8); d->append(b, 8); }
inline uint32_t DecodeFixed32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline uint64_t DecodeFixed64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
inline void PutVarint64(std::string* d, uint64_t v) {
  while (v >= 128) { d->push_back(static_cast<char>(v | 128)); v >>= 7; }
  d->push_back(static_cast<char>(v));
}
inline void PutVarint32(std::string* d, uint32_t v) { PutVarint64(d, v); }
inline int VarintLength(uint64_t v) { int n = 1; while (v >= 128) { v >>= 7; n++; } return n; }
inline const uint8_t* GetVarint64Ptr(const uint8_t* p, const uint8_t* limit, uint64_t* v) {
  uint64_t r = 0;
  for (uint32_t shift = 0; shift <= 63 && p < limit; shift += 7) {
    uint64_t b = *p++;
    if (b & 128) r |= (b & 127) << shift; else { r |= b << shift; *v = r; return p; }
  }
  return nullptr;
}
inline const uint8_t* GetVarint32Ptr(const uint8_t* p, const uint8_t* limit, uint32_t* v) {
  uint32_t r = 0;
  for (uint32_t shift = 0; shift <= 28 && p < limit; shift += 7) {
    uint32_t b = *p++;
    if (b & 128) r |= (b & 127) << shift; else { r |= b << shift; *v = r; return p; }
  }
  return nullptr;
}

// ---------------------------------------------------------------------------------------------
// rocksdb/util/crc32c.{h,cc}: CRC-32C (Castagnoli), Mask = rotr15 + 0xa282ead8 (crc32c.h:51-60).
uint32_t Crc32cExtend(uint32_t crc, const uint8_t* data, size_t n);
inline uint32_t Crc32cValue(const uint8_t* data, size_t n) { return Crc32cExtend(0, data, n); }
inline uint32_t Crc32cMask(uint32_t crc) { return ((crc >> 15) | (crc << 17)) + 0xa282ead8u; }
inline uint32_t Crc32cUnmask(uint32_t m) { uint32_t r = m - 0xa282ead8u; return (r >> 17) | (r << 15); }

// ---------------------------------------------------------------------------------------------
// util/fast_varint.cc:60-150 (signed), :260-330 (unsigned).
int FastEncodeSignedVarInt(int64_t v, uint8_t* dest);                    // returns size
inline int FastEncodeDescendingSignedVarInt(int64_t v, uint8_t* dest) { return FastEncodeSignedVarInt(-v, dest); }
// returns bytes consumed; throws Corruption
size_t FastDecodeSignedVarInt(const uint8_t* src, size_t n, int64_t* v);
int FastEncodeUnsignedVarInt(uint64_t v, uint8_t* dest);
size_t FastDecodeUnsignedVarInt(const uint8_t* src, size_t n, uint64_t* v);
size_t FastDecodeDescendingSignedVarIntSize(const uint8_t* src, size_t n);  // fast_varint.cc:171-179

// ---------------------------------------------------------------------------------------------
// common/hybrid_time.h:68-97,213-228; common/doc_hybrid_time.{h,cc}.
constexpr uint64_t kYugaByteMicrosecondEpoch = 1500000000ull * 1000000;   // doc_hybrid_time.h:107
constexpr int kBitsForLogical = 12;
constexpr uint64_t kLogicalMask = (1u << kBitsForLogical) - 1;
constexpr uint64_t kHtMin = 0, kHtMax = ~0ull, kHtInvalid = ~0ull - 1;
constexpr uint32_t kMinWriteId = 0, kMaxWriteId = 0xffffffffu;
constexpr size_t kMaxBytesPerEncodedHybridTime = 30;

inline uint64_t HtFromMicros(uint64_t micros, uint32_t logical = 0) { return (micros << kBitsForLogical) + logical; }
inline uint64_t HtMicros(uint64_t ht) { return ht >> kBitsForLogical; }
inline uint32_t HtLogical(uint64_t ht) { return static_cast<uint32_t>(ht & kLogicalMask); }

// doc_hybrid_time.cc:39-76. Returns encoded size (<= 30).
int EncodeDocHt(uint64_t ht_repr, uint32_t write_id, uint8_t* dest);
// Decode from exact encoded slice; doc_hybrid_time.cc DecodeFrom.
void DecodeDocHt(Slice enc, uint64_t* ht_repr, uint32_t* write_id);
// doc_hybrid_time.cc:194-231: size from the low 5 bits of the last byte, with the same checks.
size_t DocHtEncodedSizeFromEnd(Slice key);

struct EncodedDocHt {
  uint8_t b[kMaxBytesPerEncodedHybridTime];
  uint8_t n = 0;
  EncodedDocHt() {}
  EncodedDocHt(uint64_t ht, uint32_t wid) { n = static_cast<uint8_t>(EncodeDocHt(ht, wid, b)); }
  explicit EncodedDocHt(Slice s) { n = static_cast<uint8_t>(s.n); memcpy(b, s.p, s.n); }
  Slice slice() const { return Slice(b, n); }
  bool empty() const { return n == 0; }
};
// doc_hybrid_time.h:86-92: ordering is the REVERSED bytewise order of the encodings.
inline int CompareEncHt(const EncodedDocHt& l, const EncodedDocHt& r) { return r.slice().compare(l.slice()); }

// ---------------------------------------------------------------------------------------------
// dockv/value_type.h:30-216 — the key-entry type bytes this path looks at.
namespace kt {
constexpr uint8_t kLowest = 0, kVectorIndexMetadata = 6, kTransactionApplyState = 7,
    kExternalTransactionId = 8, kObsoleteIntentPrefix = 10, kIntentTypeSet = 13,
    kObsoleteIntentTypeSet = 15, kObsoleteIntentType = 20, kGreaterThanIntentType = 21,
    kGroupEnd = '!', kHybridTime = '#', kNullLow = '$', kCounter = '%', kSSForward = '&',
    kSSReverse = '\'', kInetaddress = '-', kInetaddressDescending = '.', kColocationId = '0',
    kWeakObjectLock = '3', kStrongObjectLock = '4', kFrozen = '<', kFrozenDescending = '>',
    kVarInt = 'B', kFloat = 'C', kDouble = 'D', kDecimal = 'E', kFalse = 'F', kUInt16Hash = 'G',
    kInt32 = 'H', kInt64 = 'I', kSystemColumnId = 'J', kColumnId = 'K', kDoubleDescending = 'L',
    kFloatDescending = 'M', kUInt32 = 'O', kString = 'S', kTrue = 'T', kUInt64 = 'U',
    kVectorId = 'V', kExternalIntents = 'Z', kArrayIndex = '[', kCollString = '\\',
    kCollStringDescending = ']', kUuid = '_', kUuidDescending = '`', kStringDescending = 'a',
    kInt64Descending = 'b', kTimestampDescending = 'c', kDecimalDescending = 'd',
    kInt32Descending = 'e', kVarIntDescending = 'f', kUInt32Descending = 'g',
    kTrueDescending = 'h', kFalseDescending = 'i', kUInt64Descending = 'j', kMergeFlags = 'k',
    kBitSet = 'm', kSubTransactionId = 'n', kBson = 'o', kBsonDescending = 'p', kTimestamp = 's',
    kTtl = 't', kUserTimestamp = 'u', kGinNull = 'v', kTransactionId = 'x', kTableId = 'y',
    kObject = '{', kNullHigh = '|', kGroupEndDescending = '}', kHighest = '~', kInvalid = 127,
    kMaxByte = 0xff;
}
namespace vt {  // ValueEntryType (dockv/value_type.h:160-216)
constexpr uint8_t kTombstone = 'X', kPackedRowV1 = 'z', kPackedRowV2 = '|', kObject = '{';
}

// dockv/primitive_value.cc:1232-1626 KeyEntryValue::DecodeKey(slice, nullptr): consume one key
// entry (type byte + payload). Throws Corruption / NotSupported.
size_t VarIntComparableSize(Slice slice, size_t num_reserved_bits);   // util/varint.cc:159-205
size_t DecimalComparableSize(Slice slice);                            // util/decimal.cc:339-367
void SkipKeyEntry(Slice* s);
// dockv/doc_key.cc:417-422,543-590 DocKey::EncodedSize(slice, part): part 0 = kUpToId,
// 1 = kWholeDocKey, 2 = kUpToHashOrFirstRange (hashed components, or the first range component
// of a key without a hash code).
size_t DocKeyEncodedSize(Slice s, int part);
// DocKey::PartiallyDecode (doc_key.cc:398-406 with DecodeDocKeyCallback :280-300): the encoded range-group components of the
// DocKey at the front of `key` (hashed components are not reported).
void DocKeyRangeComponents(Slice key, std::vector<Slice>* out);
// dockv/doc_key.cc:963-996 SubDocKey::DecodeDocKeyAndSubKeyEnds (incremental on *out).
void DecodeDocKeyAndSubKeyEnds(Slice key, std::vector<size_t>* out);

// ---------------------------------------------------------------------------------------------
// dockv/value.cc:77-143 ValueControlFields.
constexpr int64_t kMaxTtlNs = INT64_MAX;         // MonoDelta::kMax (util/monotime.cc:87)
constexpr int64_t kInvalidTimestamp = INT64_MIN; // common/table_properties_constants.h:27
struct ControlFields {
  uint64_t merge_flags = 0;
  int64_t ttl_ns = kMaxTtlNs;
  int64_t timestamp = kInvalidTimestamp;
};
// Decodes the optional prefixes; advances *v; *intent_doc_ht gets the encoded intent HT (may be
// empty).
ControlFields DecodeControlFields(Slice* v, Slice* intent_doc_ht);
void AppendControlFields(const ControlFields& f, std::string* out);   // value.cc:118-132

}  // namespace orc
