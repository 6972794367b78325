// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_codec.h).
#include "oracle_codec.h"

#if defined(__SSE4_2__)
#include <nmmintrin.h>
#endif

namespace orc {

// ---------------------------------------------------------------------------------------------
// CRC-32C, reflected polynomial 0x82F63B78. rocksdb/util/crc32c.cc uses the SSE4.2 instruction
// when the CPU has it and a table fallback otherwise; both give the standard CRC-32C.
static uint32_t g_crc_table[8][256];
static bool g_crc_init = [] {
  for (uint32_t i = 0; i < 256; i++) {
    uint32_t c = i;
    for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
    g_crc_table[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; i++)
    for (int t = 1; t < 8; t++)
      g_crc_table[t][i] = (g_crc_table[t - 1][i] >> 8) ^ g_crc_table[0][g_crc_table[t - 1][i] & 0xff];
  return true;
}();

uint32_t Crc32cExtend(uint32_t crc, const uint8_t* p, size_t n) {
  uint32_t c = crc ^ 0xffffffffu;
#if defined(__SSE4_2__)
  uint64_t c64 = c;
  while (n >= 8) { uint64_t v; memcpy(&v, p, 8); c64 = _mm_crc32_u64(c64, v); p += 8; n -= 8; }
  c = static_cast<uint32_t>(c64);
  while (n--) c = _mm_crc32_u8(c, *p++);
#else
  while (n >= 8) {
    uint64_t v; memcpy(&v, p, 8); v ^= c;
    c = g_crc_table[7][v & 0xff] ^ g_crc_table[6][(v >> 8) & 0xff] ^ g_crc_table[5][(v >> 16) & 0xff] ^
        g_crc_table[4][(v >> 24) & 0xff] ^ g_crc_table[3][(v >> 32) & 0xff] ^
        g_crc_table[2][(v >> 40) & 0xff] ^ g_crc_table[1][(v >> 48) & 0xff] ^ g_crc_table[0][v >> 56];
    p += 8; n -= 8;
  }
  while (n--) c = g_crc_table[0][(c ^ *p++) & 0xff] ^ (c >> 8);
#endif
  return c ^ 0xffffffffu;
}

// ---------------------------------------------------------------------------------------------
// util/fast_varint.cc:49-150. First bit = sign (1 positive), then a unary length prefix, then the
// magnitude; negatives are the one's complement of the positive encoding of |v|.
static int SignedPositiveVarIntLength(uint64_t v) {
  v >>= 6;
  int n = 1;
  while (v) { v >>= 7; n++; }
  return n;
}

int FastEncodeSignedVarInt(int64_t v, uint8_t* dest) {
  bool neg = v < 0;
  uint64_t uv = static_cast<uint64_t>(v);
  if (neg) uv = 1 + ~uv;
  const int n = SignedPositiveVarIntLength(uv);
  int i;
  if (n == 10) { dest[0] = 0xff; dest[1] = 0xc0; i = 2; }
  else if (n == 9) { dest[0] = 0xff; dest[1] = static_cast<uint8_t>(0x80 | (uv >> 56)); i = 2; }
  else { dest[0] = static_cast<uint8_t>(~((1 << (8 - n)) - 1) | (uv >> (8 * (n - 1)))); i = 1; }
  for (; i < n; i++) dest[i] = static_cast<uint8_t>(uv >> (8 * (n - 1 - i)));
  if (neg) for (i = 0; i < n; i++) dest[i] = ~dest[i];
  return n;
}

static const uint64_t kVarIntMasks[] = {
    0, 0x3fULL, 0x1fffULL, 0xfffffULL, 0x7ffffffULL, 0x3ffffffffULL, 0x1ffffffffffULL,
    0xffffffffffffULL, 0x7fffffffffffffULL, 0x3fffffffffffffffULL, 0xffffffffffffffffULL};

size_t FastDecodeDescendingSignedVarIntSize(const uint8_t* src, size_t n) {
  if (n == 0) return 0;
  uint16_t header = static_cast<uint16_t>(src[0] << 8 | (n > 1 ? src[1] : 0));
  uint64_t negative = -static_cast<uint64_t>((header & 0x8000) == 0);
  header ^= static_cast<uint16_t>(negative);
  return __builtin_clz((~header & 0x7fff) | 0x20) - 16;
}

size_t FastDecodeSignedVarInt(const uint8_t* src, size_t n, int64_t* v) {
  if (n == 0) throw Corruption("Cannot decode a variable-length integer of zero size");
  uint16_t header = static_cast<uint16_t>(src[0] << 8 | (n > 1 ? src[1] : 0));
  uint64_t negative = -static_cast<uint64_t>((header & 0x8000) == 0);
  header ^= static_cast<uint16_t>(negative);
  const size_t nb = __builtin_clz((~header & 0x7fff) | 0x20) - 16;
  if (n < nb) throw Corruption("Decoded VarInt size larger than bytes provided");
  uint64_t mask = kVarIntMasks[nb];
  // Big-endian load of the nb bytes; for nb > 8 only the last 8 carry magnitude bits
  // (fast_varint.cc:212-229).
  uint64_t temp = 0;
  size_t start = nb > 8 ? nb - 8 : 0;
  for (size_t i = start; i < nb; i++) temp = (temp << 8) | src[i];
  *v = static_cast<int64_t>(((temp & mask) | (~mask & negative)) - negative);
  return nb;
}

int FastEncodeUnsignedVarInt(uint64_t v, uint8_t* dest) {
  int n = 1;
  for (uint64_t t = v >> 7; t; t >>= 7) n++;
  int i;
  if (n == 10) { dest[0] = 0xff; dest[1] = 0x80; i = 2; }
  else if (n == 9) { dest[0] = 0xff; dest[1] = static_cast<uint8_t>(v >> 56); i = 2; }
  else { dest[0] = static_cast<uint8_t>(~((1 << (9 - n)) - 1) | (v >> (8 * (n - 1)))); i = 1; }
  for (; i < n; i++) dest[i] = static_cast<uint8_t>(v >> (8 * (n - 1 - i)));
  return n;
}

size_t FastDecodeUnsignedVarInt(const uint8_t* src, size_t n, uint64_t* v) {
  if (n == 0) throw Corruption("Cannot decode a variable-length integer of zero size");
  uint8_t first = src[0];
  size_t nb = __builtin_clz((static_cast<unsigned>(first) << 1) ^ 0x1ff) - 23 + 1;
  if (n < nb) throw Corruption("unsigned varint truncated");
  if (nb == 1) { *v = first & 0x7f; return 1; }
  uint64_t r = 0;
  size_t i = 0;
  if (nb == 9) {
    if (src[1] & 0x80) { nb = 10; r = src[1] & 0x3f; i = 2; }
    if (n < nb) throw Corruption("unsigned varint truncated");
  } else {
    r = src[0] & ((1 << (8 - nb)) - 1);
    i = 1;
  }
  for (; i < nb; i++) r = (r << 8) | src[i];
  *v = r;
  return nb;
}

// ---------------------------------------------------------------------------------------------
// common/doc_hybrid_time.cc:39-76.
int EncodeDocHt(uint64_t ht, uint32_t write_id, uint8_t* dest) {
  uint8_t* out = dest;
  out += FastEncodeDescendingSignedVarInt(0, out);   // generation
  out += FastEncodeDescendingSignedVarInt(
      static_cast<int64_t>(HtMicros(ht) - kYugaByteMicrosecondEpoch), out);
  out += FastEncodeDescendingSignedVarInt(HtLogical(ht), out);
  out += FastEncodeDescendingSignedVarInt((static_cast<int64_t>(write_id) + 1) << 5, out);
  const uint8_t size = static_cast<uint8_t>(out - dest);
  out[-1] = static_cast<uint8_t>((out[-1] & ~0x1f) | size);
  return size;
}

void DecodeDocHt(Slice enc, uint64_t* ht, uint32_t* write_id) {
  const size_t total = enc.n;
  int64_t v;
  size_t k = FastDecodeSignedVarInt(enc.p, enc.n, &v); enc.remove_prefix(k);           // generation
  k = FastDecodeSignedVarInt(enc.p, enc.n, &v); enc.remove_prefix(k);
  int64_t micros = static_cast<int64_t>(kYugaByteMicrosecondEpoch) + (-v);
  k = FastDecodeSignedVarInt(enc.p, enc.n, &v); enc.remove_prefix(k);
  int64_t logical = -v;
  if (logical < 0 || logical > 0xffffffffLL) throw Corruption("bad logical");
  k = FastDecodeSignedVarInt(enc.p, enc.n, &v); enc.remove_prefix(k);
  int64_t shifted = -v;
  if (shifted < 0) throw Corruption("Negative decoded_shifted_write_id");
  int64_t wid = (shifted >> 5) - 1;
  if (wid < 0 || wid > 0xffffffffLL) throw Corruption("bad write id");
  size_t decoded = total - enc.n;
  if (((enc.p[-1]) & 0x1f) != decoded) throw Corruption("Wrong encoded DocHybridTime size at the end");
  *ht = (static_cast<uint64_t>(micros) << kBitsForLogical) + static_cast<uint64_t>(logical);
  *write_id = static_cast<uint32_t>(wid);
}

size_t DocHtEncodedSizeFromEnd(Slice key) {
  if (key.n == 0) throw Corruption("empty key when looking for DocHybridTime at the end");
  size_t r = key.p[key.n - 1] & 0x1f;
  if (r < 1) throw Corruption("Encoded HybridTime must be at least one byte");
  if (r > kMaxBytesPerEncodedHybridTime) throw Corruption("Encoded HybridTime too long");
  if (r >= key.n) throw Corruption("Encoded HybridTime does not leave room for the value type");
  return r;
}

// ---------------------------------------------------------------------------------------------
// util/varint.cc:159-205 VarInt::DecodeFromComparable, the num_decoded_bytes it reports: sign bit
// (after `num_reserved_bits` reserved bits), then a unary count of bytes, then the magnitude;
// negative numbers are stored complemented.
size_t VarIntComparableSize(Slice slice, size_t num_reserved_bits) {
  const size_t len = slice.n;
  if (len == 0) throw Corruption("Cannot decode varint from empty slice");
  const bool negative = (slice[0] & (0x80 >> num_reserved_bits)) == 0;
  std::vector<uint8_t> buffer(slice.p, slice.p + slice.n);
  if (negative) for (auto& b : buffer) b = static_cast<uint8_t>(~b);
  if (num_reserved_bits) buffer[0] |= static_cast<uint8_t>(~((1 << (8 - num_reserved_bits)) - 1));
  size_t idx = 0, num_ones = 0;
  while (buffer[idx] == 0xff) {
    ++idx;
    if (idx >= len) throw Corruption("Encoded varint failure, no prefix termination");
    num_ones += 8;
  }
  for (uint8_t temp = 0x80; buffer[idx] & temp; temp >>= 1) ++num_ones;
  num_ones -= num_reserved_bits;
  if (num_ones > len) throw Corruption("Not enough data in encoded varint");
  return num_ones;
}

// util/decimal.cc:311-367 Decimal::DecodeFromComparable, the num_decoded_bytes it reports: 0x80 is
// zero; otherwise sign from the first bit (negatives complemented), exponent as a varint with two
// reserved bits, then mantissa digit pairs, the last of which has its low bit clear.
size_t DecimalComparableSize(Slice slice) {
  if (slice.empty()) throw Corruption("Cannot decode Decimal from empty slice.");
  if (slice[0] == 128) return 1;
  const bool positive = slice[0] >= 128;
  std::vector<uint8_t> encoded(slice.p, slice.p + slice.n);
  if (!positive) for (auto& b : encoded) b = static_cast<uint8_t>(~b);
  const size_t num_exponent_bytes = VarIntComparableSize(Slice(encoded.data(), encoded.size()), 2);
  for (size_t i = num_exponent_bytes; i < encoded.size(); i++)
    if (!(encoded[i] & 1)) return i + 1;
  throw Corruption("Decoded the whole slice but didn't find the ending");
}

// ---------------------------------------------------------------------------------------------
// dockv/doc_kv_util.cc:60-101 DecodeEncodedStr<kEnd>: scan to the terminator pair (kEnd kEnd);
// (kEnd kEnd^1) is an escaped kEnd byte.
static void SkipEncodedStr(Slice* s, uint8_t end_byte) {
  const uint8_t* p = s->p;
  const uint8_t* end = s->p + s->n;
  if (p == end) throw Corruption("Encoded string is empty");
  for (;;) {
    const uint8_t* stop = static_cast<const uint8_t*>(memchr(p, end_byte, end - p));
    if (!stop) throw Corruption("Encoded string is not terminated");
    if (stop >= end - 1) throw Corruption("Encoded string ends with only one terminator byte");
    if (stop[1] == end_byte) { p = stop + 2; break; }
    if (stop[1] != (end_byte ^ 1)) throw Corruption("Invalid sequence in encoded string");
    p = stop + 2;
    if (p == end) break;   // reference loop condition `while (p != end)`
  }
  s->remove_prefix(p - s->p);
}

static void NeedBytes(const Slice& s, size_t n, const char* what) {
  if (s.n < n) throw Corruption(std::string("Not enough bytes to decode ") + what);
}

void SkipKeyEntry(Slice* s) {
  if (s->empty()) throw Corruption("Cannot decode a primitive value from an empty slice");
  uint8_t t = (*s)[0];
  s->remove_prefix(1);
  switch (t) {
    // CASE_EMPTY_KEY_ENTRY_TYPES (primitive_value.cc:750-765)
    case kt::kVectorIndexMetadata: case kt::kCounter: case kt::kFalse: case kt::kFalseDescending:
    case kt::kHighest: case kt::kLowest: case kt::kNullHigh: case kt::kNullLow:
    case kt::kSSForward: case kt::kSSReverse: case kt::kTrue: case kt::kObject:
    case kt::kWeakObjectLock: case kt::kStrongObjectLock: case kt::kTrueDescending:
      return;
    case kt::kCollStringDescending: case kt::kStringDescending:
      SkipEncodedStr(s, 0xff); return;
    case kt::kCollString: case kt::kString:
      SkipEncodedStr(s, 0x00); return;
    case kt::kFrozenDescending: case kt::kFrozen: {
      uint8_t end_marker = t == kt::kFrozenDescending ? kt::kGroupEndDescending : kt::kGroupEnd;
      while (!s->empty()) {
        if ((*s)[0] == end_marker) { s->remove_prefix(1); return; }
        SkipKeyEntry(s);
      }
      throw Corruption("Reached end of slice looking for frozen group end marker");
    }
    case kt::kDecimalDescending: case kt::kDecimal:
      // primitive_value.cc:1314-1332: both orders decode the bytes as they are (the descending
      // form is the comparable encoding of the negated number)
      s->remove_prefix(DecimalComparableSize(*s)); return;
    case kt::kVarIntDescending: case kt::kVarInt:
      s->remove_prefix(VarIntComparableSize(*s, 0)); return;        // primitive_value.cc:1334-1349
    case kt::kGinNull:
      NeedBytes(*s, 1, "gin null"); s->remove_prefix(1); return;
    case kt::kInt32Descending: case kt::kInt32: case kt::kColocationId: case kt::kUInt32Descending:
    case kt::kSubTransactionId: case kt::kUInt32: case kt::kFloatDescending: case kt::kFloat:
      NeedBytes(*s, 4, "32-bit value"); s->remove_prefix(4); return;
    case kt::kUInt64Descending: case kt::kUInt64: case kt::kInt64Descending: case kt::kInt64:
    case kt::kArrayIndex: case kt::kTimestampDescending: case kt::kTimestamp:
    case kt::kDoubleDescending: case kt::kDouble:
      NeedBytes(*s, 8, "64-bit value"); s->remove_prefix(8); return;
    case kt::kUInt16Hash:
      NeedBytes(*s, 2, "16-bit hash"); s->remove_prefix(2); return;
    case kt::kInetaddress: SkipEncodedStr(s, 0x00); return;
    case kt::kInetaddressDescending: SkipEncodedStr(s, 0xff); return;
    case kt::kTransactionApplyState: case kt::kExternalTransactionId: case kt::kVectorId:
      NeedBytes(*s, 16, "UUID"); s->remove_prefix(16); return;
    case kt::kTransactionId: case kt::kTableId: case kt::kUuid:
      SkipEncodedStr(s, 0x00); return;
    case kt::kUuidDescending:
      SkipEncodedStr(s, 0xff); return;
    case kt::kColumnId: case kt::kSystemColumnId: {
      int64_t v; size_t k = FastDecodeSignedVarInt(s->p, s->n, &v);   // common/column_id.cc:40-43
      if (v < 0 || v > 0x7fffffff) throw Corruption("not valid for column id representation");
      s->remove_prefix(k); return;
    }
    case kt::kHybridTime: {
      // DocHybridTime::DecodeFrom: four varints.
      for (int i = 0; i < 4; i++) { int64_t v; size_t k = FastDecodeSignedVarInt(s->p, s->n, &v); s->remove_prefix(k); }
      return;
    }
    case kt::kIntentTypeSet: case kt::kObsoleteIntentTypeSet: case kt::kObsoleteIntentType:
      NeedBytes(*s, 1, "TypeSet"); s->remove_prefix(1); return;
    case kt::kBson: SkipEncodedStr(s, 0x00); return;             // dockv/doc_bson.cc:33-35: a zero-encoded string
    case kt::kBsonDescending: SkipEncodedStr(s, 0xff); return;   // :46-48: its complement
    default:
      throw Corruption("Cannot decode value type from the key encoding format");
  }
}

static bool IsSpecialKeyEntryType(uint8_t t) {   // value_type.h:280-284
  return t == kt::kLowest || t == kt::kHighest || t == kt::kMaxByte || t == kt::kIntentTypeSet ||
         t == kt::kGreaterThanIntentType;
}

// doc_key.cc:52-89 HasPrimitiveValue + ConsumePrimitiveValuesFromKey (allow_special = false).
static void ConsumePrimitiveValues(Slice* s) {
  for (;;) {
    if (s->empty()) throw Corruption("Unexpected end of key when decoding document key");
    uint8_t t = (*s)[0];
    if (t == kt::kGroupEnd) { s->remove_prefix(1); return; }
    if (IsSpecialKeyEntryType(t)) throw Corruption("Expected a primitive value type");
    SkipKeyEntry(s);
  }
}

size_t DocKeyEncodedSize(Slice s, int part) {
  const uint8_t* begin = s.p;
  // doc_key.cc:1229-1270: cotable id 'y' + 16 bytes, else colocation id '0' + 4 bytes.
  if (!s.empty() && s[0] == kt::kTableId) {
    s.remove_prefix(1);
    NeedBytes(s, 16, "cotable id"); s.remove_prefix(16);
  } else if (!s.empty() && s[0] == kt::kColocationId) {
    s.remove_prefix(1);
    NeedBytes(s, 4, "colocation id"); s.remove_prefix(4);
  }
  if (part == 0) return s.p - begin;
  // DecodeHashCode (doc_key.cc:1272-1310)
  bool hash_present = false;
  if (!s.empty()) {
    uint8_t t = s[0];
    if (t != kt::kGroupEnd) {
      if (IsSpecialKeyEntryType(t)) throw Corruption("Expected first value type to be primitive or GroupEnd");
      if (t == kt::kUInt16Hash) {
        NeedBytes(s, 3, "16-bit hash component"); s.remove_prefix(3);
        hash_present = true;
      }
    }
  }
  if (hash_present) ConsumePrimitiveValues(&s);
  if (s.empty()) return s.p - begin;
  if (part == 2) {
    // MaxRangeComponentsToDecode (doc_key.cc:523-538): none after hashed components, else one
    if (!hash_present) {
      if (s[0] == kt::kGroupEnd) s.remove_prefix(1);
      else if (IsSpecialKeyEntryType(s[0])) throw Corruption("Expected a primitive value type");
      else SkipKeyEntry(&s);
    }
    return s.p - begin;
  }
  ConsumePrimitiveValues(&s);   // range group
  return s.p - begin;
}

void DocKeyRangeComponents(Slice key, std::vector<Slice>* out) {
  Slice s = key;
  s.remove_prefix(DocKeyEncodedSize(s, 0));                 // cotable / colocation id
  bool hash_present = false;
  if (!s.empty()) {
    uint8_t t = s[0];
    if (t != kt::kGroupEnd) {
      if (IsSpecialKeyEntryType(t)) throw Corruption("Expected first value type to be primitive or GroupEnd");
      if (t == kt::kUInt16Hash) { NeedBytes(s, 3, "16-bit hash component"); s.remove_prefix(3); hash_present = true; }
    }
  }
  if (hash_present) ConsumePrimitiveValues(&s);             // hashed group: no slices (hashed_group() == nullptr)
  if (s.empty()) return;
  for (;;) {                                                // range group, one slice per component
    if (s.empty()) throw Corruption("Unexpected end of key when decoding document key");
    uint8_t t = s[0];
    if (t == kt::kGroupEnd) return;
    if (IsSpecialKeyEntryType(t)) throw Corruption("Expected a primitive value type");
    const uint8_t* b = s.p;
    SkipKeyEntry(&s);
    out->push_back(Slice(b, s.p - b));
  }
}

void DecodeDocKeyAndSubKeyEnds(Slice key, std::vector<size_t>* out) {
  Slice s = key;
  if (out->empty()) out->push_back(DocKeyEncodedSize(s, 0));
  if (out->size() == 1) {
    size_t id_size = out->front();
    if (s.n < id_size + 1) throw Corruption("Cannot have exclusively ID in key");
    if ((s[0] == kt::kColocationId || s[0] == kt::kTableId) && s[id_size] == kt::kGroupEnd) {
      // Table tombstone: id ! # HT.
      if (s.n < id_size + 2) throw Corruption("Space for kHybridTime expected in key");
      if (s[id_size + 1] != kt::kHybridTime) throw Corruption("Hybrid time expected in key");
      s.remove_prefix(id_size + 1);
    } else {
      s.remove_prefix(id_size);
      size_t dk = DocKeyEncodedSize(s, 1);
      s.remove_prefix(dk);
      out->push_back(id_size + dk);
    }
  } else {
    s.remove_prefix(out->back());
  }
  // SubDocKey::DecodeSubkey (doc_key.cc:827-838): subkeys until '#'.
  while (!s.empty() && s[0] != kt::kHybridTime) {
    SkipKeyEntry(&s);
    out->push_back(s.p - key.p);
  }
}

// ---------------------------------------------------------------------------------------------
ControlFields DecodeControlFields(Slice* v, Slice* intent_doc_ht) {
  ControlFields r;
  if (intent_doc_ht) *intent_doc_ht = Slice();
  if (v->empty()) return r;
  auto cur = [&]() -> uint8_t { return v->empty() ? kt::kInvalid : (*v)[0]; };
  uint8_t t = (*v)[0];
  if (t == kt::kMergeFlags) {
    v->remove_prefix(1);
    v->remove_prefix(FastDecodeUnsignedVarInt(v->p, v->n, &r.merge_flags));
    t = cur();
  }
  if (t == kt::kHybridTime) {
    v->remove_prefix(1);
    // DocHybridTime::EncodedFromStart (doc_hybrid_time.cc:86-104): four varints by size.
    const uint8_t* b = v->p; const uint8_t* e = v->p + v->n;
    for (int i = 0; i < 4; i++) {
      size_t sz = FastDecodeDescendingSignedVarIntSize(b, e - b);
      if (sz == 0 || b + sz > e) throw Corruption("Bad doc hybrid time");
      b += sz;
    }
    if (intent_doc_ht) *intent_doc_ht = Slice(v->p, b - v->p);
    v->remove_prefix(b - v->p);
    t = cur();
  }
  if (t == kt::kTtl) {
    v->remove_prefix(1);
    int64_t ms; v->remove_prefix(FastDecodeSignedVarInt(v->p, v->n, &ms));
    r.ttl_ns = ms * 1000000;   // MonoDelta::FromMilliseconds
    t = cur();
  }
  if (t == kt::kUserTimestamp) {
    v->remove_prefix(1);
    NeedBytes(*v, 8, "user timestamp");
    uint64_t be = 0; for (int i = 0; i < 8; i++) be = (be << 8) | (*v)[i];
    r.timestamp = static_cast<int64_t>(be);
    v->remove_prefix(8);
  }
  return r;
}

void AppendControlFields(const ControlFields& f, std::string* out) {
  uint8_t buf[16];
  if (f.merge_flags) {
    out->push_back(static_cast<char>(kt::kMergeFlags));
    int n = FastEncodeUnsignedVarInt(f.merge_flags, buf);
    out->append(reinterpret_cast<char*>(buf), n);
  }
  if (f.ttl_ns != kMaxTtlNs) {
    out->push_back(static_cast<char>(kt::kTtl));
    int n = FastEncodeSignedVarInt(f.ttl_ns / 1000000, buf);   // MonoDelta::ToMilliseconds
    out->append(reinterpret_cast<char*>(buf), n);
  }
  if (f.timestamp != kInvalidTimestamp) {
    out->push_back(static_cast<char>(kt::kUserTimestamp));
    uint64_t v = static_cast<uint64_t>(f.timestamp);
    for (int i = 7; i >= 0; i--) out->push_back(static_cast<char>(v >> (8 * i)));
  }
}

}  // namespace orc
