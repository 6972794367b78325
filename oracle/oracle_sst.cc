// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_codec.h).
#include "oracle_sst.h"
#include <cmath>
#include <cstring>

namespace orc {

// ---------------------------------------------------------------------------------------------
// table/block_builder.cc:119-141 FindMaxSharedSubstringAtTheSamePos. Note (as in the reference)
// a run that reaches the end of the compared range is NOT considered — only runs terminated by
// a mismatch update the maximum.
static std::pair<size_t, size_t> MaxSharedSamePos(const uint8_t* l, const uint8_t* r, size_t n) {
  size_t best = 0, best_off = 0, cur = 0;
  for (size_t i = 0; i < n; i++) {
    if (l[i] == r[i]) {
      cur++;
    } else {
      if (cur > best) { best = cur; best_off = i - cur; }
      cur = 0;
    }
  }
  return {best_off, best};
}

struct ComponentSizes {
  size_t prev_ns1 = 0, ns1 = 0, mid = 0, prev_ns2 = 0, ns2 = 0;
};

// table/block_builder.cc:163-220 FindMaxSharedMiddle.
static ComponentSizes FindMaxSharedMiddle(Slice lhs, Slice rhs) {
  size_t min_len;
  std::pair<size_t, size_t> ms;
  bool from_left = true;
  if (lhs.n == rhs.n) {
    min_len = rhs.n;
    ms = MaxSharedSamePos(lhs.p, rhs.p, min_len);
  } else {
    const uint8_t *ls, *rs;
    if (lhs.n > rhs.n) { min_len = rhs.n; ls = lhs.p + lhs.n - min_len; rs = rhs.p; }
    else { min_len = lhs.n; ls = lhs.p; rs = rhs.p + rhs.n - min_len; }
    ms = MaxSharedSamePos(lhs.p, rhs.p, min_len);
    auto mr = MaxSharedSamePos(ls, rs, min_len);
    if (mr.second > ms.second) { from_left = false; ms = mr; }
  }
  ComponentSizes c;
  if (ms.second == 0) { c.prev_ns1 = lhs.n; c.ns1 = rhs.n; return c; }
  if (from_left) {
    size_t a = ms.first + ms.second;
    c.prev_ns1 = ms.first; c.ns1 = ms.first; c.mid = ms.second;
    c.prev_ns2 = lhs.n - a; c.ns2 = rhs.n - a;
  } else {
    size_t mid_plus_ns2 = min_len - ms.first;
    size_t ns2 = mid_plus_ns2 - ms.second;
    c.prev_ns1 = lhs.n - mid_plus_ns2; c.ns1 = rhs.n - mid_plus_ns2; c.mid = ms.second;
    c.prev_ns2 = ns2; c.ns2 = ns2;
  }
  return c;
}

// table/block_builder_internal.h:101-239 EncodeThreeSharedPartsSizes.
static void EncodeThreeSharedPartsSizes(size_t shared_prefix, size_t last_reuse, bool last_inc,
                                        const ComponentSizes& c, size_t key_size, size_t value_size,
                                        std::string* buf) {
  const int64_t d1 = static_cast<int64_t>(c.ns1) - static_cast<int64_t>(c.prev_ns1);
  const int64_t d2 = static_cast<int64_t>(c.ns2) - static_cast<int64_t>(c.prev_ns2);
  const bool frequent = last_reuse > 0 && c.ns1 == 1 && c.ns2 == 1 && d1 == 0 && d2 == 0;
  PutVarint64(buf, (static_cast<uint64_t>(value_size) << 2) | (static_cast<uint64_t>(last_inc) << 1) | frequent);
  if (frequent) { PutVarint32(buf, static_cast<uint32_t>(shared_prefix)); return; }
  const bool reused = c.ns1 < key_size;
  uint8_t tmp[16];
  if (reused) {
    if (last_reuse > 0 && d1 == 0 && (d2 == 0 || d2 == 1) && c.ns1 < 8 && c.ns2 < 4) {
      buf->push_back(static_cast<char>(0b01 | ((d2 == 1) << 2) | (c.ns1 << 3) | (c.ns2 << 6)));
    } else {
      buf->push_back(static_cast<char>(0b11 | ((last_reuse > 0) << 2) | ((d1 != 0) << 3) |
                                       ((c.ns2 != 0) << 4) | ((d2 != 0) << 5)));
      PutVarint32(buf, static_cast<uint32_t>(c.ns1));
      if (d1 != 0) { int n = FastEncodeSignedVarInt(d1, tmp); buf->append(reinterpret_cast<char*>(tmp), n); }
      if (c.ns2 != 0) PutVarint32(buf, static_cast<uint32_t>(c.ns2));
      if (d2 != 0) { int n = FastEncodeSignedVarInt(d2, tmp); buf->append(reinterpret_cast<char*>(tmp), n); }
    }
    PutVarint32(buf, static_cast<uint32_t>(shared_prefix));
  } else {
    if (key_size < 128 && key_size > 0) {
      buf->push_back(static_cast<char>(key_size << 1));
    } else {
      buf->push_back(0);
      PutVarint32(buf, static_cast<uint32_t>(key_size));
    }
  }
}

void BlockBuilder::Add(Slice key, Slice value) {
  Slice prev(last_key_);
  size_t shared = 0;
  // Defaults used on restarts / no delta encoding (block_builder.cc:355-356, ctor :270-276).
  ComponentSizes comp; comp.prev_ns1 = prev.n; comp.ns1 = key.n;
  size_t last_reuse = 0; bool last_inc = false;
  if (counter_ >= restart_interval_) {
    restarts_.push_back(static_cast<uint32_t>(buf_.size()));
    counter_ = 0;
  } else if (use_delta_) {
    const size_t min_len = std::min(prev.n, key.n);
    while (shared < min_len && prev.p[shared] == key.p[shared]) shared++;
    if (enc_ == kThreeSharedParts) {
      // CalculateLastInternalComponentReuse (block_builder.cc:222-246)
      if (min_len >= shared + kLastInternalComponentSize) {
        uint64_t pl = DecodeFixed64(prev.p + prev.n - 8), cl = DecodeFixed64(key.p + key.n - 8);
        if (cl == pl + 0x100) { last_inc = true; last_reuse = 8; }
        else if (cl == pl) last_reuse = 8;
      }
      comp = FindMaxSharedMiddle(Slice(prev.p + shared, prev.n - shared - last_reuse),
                                 Slice(key.p + shared, key.n - shared - last_reuse));
    }
  }
  const size_t non_shared = key.n - shared;
  if (enc_ == kSharedPrefix) {
    PutVarint32(&buf_, static_cast<uint32_t>(shared));
    PutVarint32(&buf_, static_cast<uint32_t>(non_shared));
    PutVarint32(&buf_, static_cast<uint32_t>(value.n));
    buf_.append(reinterpret_cast<const char*>(key.p + shared), non_shared);
    buf_.append(reinterpret_cast<const char*>(value.p), value.n);
  } else {
    EncodeThreeSharedPartsSizes(shared, last_reuse, last_inc, comp, key.n, value.n, &buf_);
    buf_.append(reinterpret_cast<const char*>(key.p + shared), comp.ns1);
    buf_.append(reinterpret_cast<const char*>(key.p + key.n - last_reuse - comp.ns2), comp.ns2);
    buf_.append(reinterpret_cast<const char*>(value.p), value.n);
  }
  last_key_.resize(shared);
  last_key_.append(reinterpret_cast<const char*>(key.p + shared), non_shared);
  counter_++;
}

Slice BlockBuilder::Finish() {
  for (uint32_t r : restarts_) PutFixed32(&buf_, r);
  PutFixed32(&buf_, static_cast<uint32_t>(restarts_.size()));
  finished_ = true;
  return Slice(buf_);
}

// ---------------------------------------------------------------------------------------------
BlockIter::BlockIter(Slice block, int key_encoding) : data_(block.p), enc_(key_encoding) {
  if (block.n < 4) throw Corruption("bad block contents");
  num_restarts_ = DecodeFixed32(block.p + block.n - 4);
  if (static_cast<uint64_t>(num_restarts_) * 4 + 4 > block.n) throw Corruption("bad block contents");
  restarts_off_ = static_cast<uint32_t>(block.n - 4 - 4 * num_restarts_);
}

void BlockIter::SeekToFirst() { next_ = 0; key_.clear(); Parse(); }
void BlockIter::Next() { Parse(); }

bool BlockIter::Parse() {
  const uint8_t* p = data_ + next_;
  const uint8_t* limit = data_ + restarts_off_;
  if (p >= limit) { valid_ = false; return false; }
  if (enc_ == kSharedPrefix) {
    // table/block.cc:65-87 DecodeEntry
    uint32_t shared, non_shared, vlen;
    if (limit - p < 3) throw Corruption("bad entry in block");
    shared = p[0]; non_shared = p[1]; vlen = p[2];
    if ((shared | non_shared | vlen) < 128) {
      p += 3;
    } else {
      if (!(p = GetVarint32Ptr(p, limit, &shared)) || !(p = GetVarint32Ptr(p, limit, &non_shared)) ||
          !(p = GetVarint32Ptr(p, limit, &vlen)))
        throw Corruption("bad entry in block");
    }
    if (static_cast<uint32_t>(limit - p) < non_shared + vlen || key_.size() < shared)
      throw Corruption("bad entry in block");
    key_.resize(shared);
    key_.append(reinterpret_cast<const char*>(p), non_shared);
    value_ = Slice(p + non_shared, vlen);
    next_ = static_cast<uint32_t>(p + non_shared + vlen - data_);
  } else {
    // table/block_internal.h:51-162 + table/block.cc:294-343 + db/dbformat.h:413-470.
    if (limit - p < 2) throw Corruption("bad entry in block");
    uint64_t e1;
    if (!(p = GetVarint64Ptr(p, limit, &e1))) throw Corruption("bad entry in block");
    uint32_t vlen = static_cast<uint32_t>(e1 >> 2);
    uint64_t last_inc = (e1 & 2) << 7;
    uint32_t shared_prefix = 0, ns1 = 0, ns2 = 0, last_sz = 0;
    int64_t d1 = 0, d2 = 0;
    bool something_shared;
    if (e1 & 1) {
      if (!(p = GetVarint32Ptr(p, limit, &shared_prefix))) throw Corruption("bad entry in block");
      last_sz = 8; something_shared = true; ns1 = 1; ns2 = 1;
    } else {
      uint8_t e2 = *p++;
      if ((e2 & 1) == 0) {
        something_shared = false;
        if (e2 == 0) { if (!(p = GetVarint32Ptr(p, limit, &ns1))) throw Corruption("bad entry in block"); }
        else ns1 = e2 >> 1;
      } else {
        something_shared = true;
        if ((e2 & 2) == 0) {
          last_sz = 8; d2 = (e2 >> 2) & 1; ns1 = (e2 >> 3) & 7; ns2 = (e2 >> 6) & 3;
        } else {
          last_sz = (e2 & 4) ? 8 : 0;
          if (!(p = GetVarint32Ptr(p, limit, &ns1))) throw Corruption("bad entry in block");
          if (e2 & 8) p += FastDecodeSignedVarInt(p, limit - p, &d1);
          if (e2 & 16) { if (!(p = GetVarint32Ptr(p, limit, &ns2))) throw Corruption("bad entry in block"); }
          if (e2 & 32) p += FastDecodeSignedVarInt(p, limit - p, &d2);
        }
        if (!(p = GetVarint32Ptr(p, limit, &shared_prefix))) throw Corruption("bad entry in block");
      }
    }
    if (limit - p < static_cast<int64_t>(ns1) + ns2 + vlen) throw Corruption("bad entry in block");
    if (!something_shared) {
      key_.assign(reinterpret_cast<const char*>(p), ns1);
      value_ = Slice(p + ns1, vlen);
      next_ = static_cast<uint32_t>(p + ns1 + vlen - data_);
    } else {
      const int64_t prev_mid_start = static_cast<int64_t>(shared_prefix) + ns1 - d1;
      const int64_t prev_ns2 = static_cast<int64_t>(ns2) - d2;
      const int64_t prev_except_mid = prev_mid_start + prev_ns2 + last_sz;
      if (static_cast<int64_t>(key_.size()) < prev_except_mid) throw Corruption("bad entry in block");
      const size_t mid = key_.size() - prev_except_mid;
      if (shared_prefix + mid + last_sz == 0) throw Corruption("bad entry in block");
      std::string nk;
      nk.reserve(shared_prefix + ns1 + mid + ns2 + last_sz);
      nk.append(key_, 0, shared_prefix);
      nk.append(reinterpret_cast<const char*>(p), ns1);
      nk.append(key_, prev_mid_start, mid);
      nk.append(reinterpret_cast<const char*>(p + ns1), ns2);
      if (last_sz) {
        uint64_t last = DecodeFixed64(reinterpret_cast<const uint8_t*>(key_.data()) + key_.size() - 8) + last_inc;
        PutFixed64(&nk, last);
      }
      key_.swap(nk);
      value_ = Slice(p + ns1 + ns2, vlen);
      next_ = static_cast<uint32_t>(p + ns1 + ns2 + vlen - data_);
    }
  }
  valid_ = true;
  return true;
}

// ---------------------------------------------------------------------------------------------
void BytewiseFindShortestSeparator(std::string* start, Slice limit) {
  size_t min_len = std::min(start->size(), limit.n);
  size_t d = 0;
  const uint8_t* sb = reinterpret_cast<const uint8_t*>(start->data());
  while (d < min_len && sb[d] == limit.p[d]) d++;
  if (d >= min_len) return;
  uint8_t s = sb[d], l = limit.p[d];
  if (s > l) return;
  if (d == limit.n - 1 && s + 1 == l) {
    ++d;
    while (d < start->size() && sb[d] == 0xff) ++d;
    if (d == start->size()) return;
  }
  (*start)[d]++;
  start->resize(d + 1);
}

void BytewiseFindShortSuccessor(std::string* key) {
  size_t n = key->size();
  for (size_t i = 0; i < n; i++) {
    uint8_t b = static_cast<uint8_t>((*key)[i]);
    if (b != 0xff) { (*key)[i] = static_cast<char>(b + 1); key->resize(i + 1); return; }
  }
}

void InternalFindShortestSeparator(std::string* start, Slice limit) {
  Slice us(reinterpret_cast<const uint8_t*>(start->data()), start->size() - 8);
  Slice ul(limit.p, limit.n - 8);
  std::string tmp = us.str();
  BytewiseFindShortestSeparator(&tmp, ul);
  if (tmp.size() < us.n && us.compare(Slice(tmp)) < 0) {
    PutFixed64(&tmp, PackSeqAndType(kMaxSequenceNumber, kValueTypeForSeek));
    start->swap(tmp);
  }
}

void InternalFindShortSuccessor(std::string* key) {
  Slice uk(reinterpret_cast<const uint8_t*>(key->data()), key->size() - 8);
  std::string tmp = uk.str();
  BytewiseFindShortSuccessor(&tmp);
  if (tmp.size() < uk.n && uk.compare(Slice(tmp)) < 0) {
    PutFixed64(&tmp, PackSeqAndType(kMaxSequenceNumber, kValueTypeForSeek));
    key->swap(tmp);
  }
}

// ---------------------------------------------------------------------------------------------
// table/index_builder.cc:143-289. Each level is a ShortenedIndexBuilder (BlockBuilder with
// shared-prefix encoding and index_block_restart_interval) cut by the size policy; a finished
// block's (last_key, next_first_key, handle) is added to the next level on the following
// FlushNextBlock call.
class MultiLevelIndexBuilder {
 public:
  explicit MultiLevelIndexBuilder(const TableOptions& o) : o_(o) {}
  void Ensure() {
    if (!cur_) {
      cur_.reset(new BlockBuilder(o_.index_block_restart_interval, kSharedPrefix));
      policy_.reset(new FlushBySize{o_.index_block_size, static_cast<uint64_t>(o_.block_size_deviation),
                                    o_.min_keys_per_index_block, cur_.get()});
    }
  }
  void AddIndexEntry(std::string* last_key, const Slice* next_first, const BlockHandle& h, bool shorten) {
    Ensure();
    std::string enc; PutVarint64(&enc, h.offset); PutVarint64(&enc, h.size);
    if (shorten) {
      if (!next_first) InternalFindShortSuccessor(last_key);
      else InternalFindShortestSeparator(last_key, *next_first);
    }
    cur_->Add(Slice(*last_key), Slice(enc));
    // NB: the policy is consulted AFTER the entry was added (index_builder.cc:186).
    if (policy_->Update(Slice(*last_key), Slice(enc)) || !next_first) {
      ready_.is_ready = true;
      ready_.last_key = *last_key;
      ready_.has_next = next_first != nullptr;
      if (next_first) ready_.next_first = next_first->str();
    }
  }
  bool ShouldFlush() const {
    return ready_.is_ready || (next_add_.is_ready && !next_add_.has_next) ||
           (next_ && next_->ShouldFlush());
  }
  // Returns true if *contents now holds a finished index block to be written.
  bool FlushNextBlock(std::string* contents, const BlockHandle& last_handle, bool last_handle_set) {
    if (next_just_flushed_) {
      next_last_flushed_ = last_handle;
      next_last_flushed_set_ = last_handle_set;
      next_just_flushed_ = false;
    }
    if (flushing_) {
      if (next_add_.is_ready) {
        Slice nf(next_add_.next_first);
        next_->AddIndexEntry(&next_add_.last_key, next_add_.has_next ? &nf : nullptr, last_handle, false);
        next_add_.is_ready = false;
      }
      if (next_ && next_->ShouldFlush()) {
        bool r = next_->FlushNextBlock(contents, next_last_flushed_, next_last_flushed_set_);
        next_just_flushed_ = true;
        return r;
      }
    }
    flushing_ = true;
    if (ready_.is_ready) {
      FlushCurrent(contents);
      if (!next_ && ready_.has_next) next_.reset(new MultiLevelIndexBuilder(o_));
      if (next_) next_add_ = ready_;
      ready_.is_ready = false;
      return true;
    } else if (!last_handle_set) {
      Ensure();
      FlushCurrent(contents);
      return true;
    }
    return false;
  }
  size_t EstimatedSize() const {
    return size_ + (next_ ? next_->EstimatedSize() : static_cast<size_t>(-static_cast<int64_t>(kBlockTrailerSize)));
  }
  int NumLevels() const { return 1 + (next_ ? next_->NumLevels() : 0); }
 private:
  struct Info { bool is_ready = false; std::string last_key; bool has_next = false; std::string next_first; };
  void FlushCurrent(std::string* contents) {
    *contents = cur_->Finish().str();
    cur_.reset(); policy_.reset();
    size_ += contents->size() + kBlockTrailerSize;
  }
  TableOptions o_;
  std::unique_ptr<BlockBuilder> cur_;
  std::unique_ptr<FlushBySize> policy_;
  Info ready_, next_add_;
  std::unique_ptr<MultiLevelIndexBuilder> next_;
  BlockHandle next_last_flushed_;
  bool next_just_flushed_ = false, flushing_ = false, next_last_flushed_set_ = false;
  size_t size_ = 0;
};

// ---------------------------------------------------------------------------------------------
// util/hash.cc:32-75. The tail bytes are added as SIGNED chars (a disk-format quirk kept on purpose).
uint32_t LevelDbHash(const uint8_t* data, size_t n, uint32_t seed) {
  const uint32_t m = 0xc6a4a793u, r = 24;
  const uint8_t* limit = data + n;
  uint32_t h = static_cast<uint32_t>(seed ^ (n * m));
  while (data + 4 <= limit) {
    h += DecodeFixed32(data); data += 4;
    h *= m; h ^= (h >> 16);
  }
  switch (limit - data) {
    case 3: h += static_cast<uint32_t>(static_cast<int32_t>(static_cast<int8_t>(data[2])) << 16);  // fallthrough
    case 2: h += static_cast<uint32_t>(static_cast<int32_t>(static_cast<int8_t>(data[1])) << 8);   // fallthrough
    case 1: h += static_cast<uint32_t>(static_cast<int32_t>(static_cast<int8_t>(data[0])));
            h *= m; h ^= (h >> r);
            break;
  }
  return h;
}

static constexpr size_t kCacheLine = 64;      // port/port_posix.h:179

FixedSizeFilterBits::FixedSizeFilterBits(size_t total_bits, double error_rate) {
  const double kLog2 = std::log(2.0);
  num_lines_ = (total_bits + kCacheLine * 8 - 1) / (kCacheLine * 8);
  if (num_lines_ % 2 == 0) {                       // bloom.cc:395-403
    if (num_lines_ * kCacheLine < 4096) num_lines_++; else num_lines_--;
  }
  total_bits_ = num_lines_ * kCacheLine * 8;
  const double minus_log_error_rate = -std::log(error_rate);
  num_probes_ = static_cast<size_t>(minus_log_error_rate / kLog2);
  num_probes_ = std::max<size_t>(num_probes_, 1);
  num_probes_ = std::min<size_t>(num_probes_, 255);
  const double max_keys = total_bits_ * kLog2 * kLog2 / minus_log_error_rate;
  max_keys_ = static_cast<size_t>(max_keys);
  data_.assign(total_bits_ / 8 + 5, '\0');
}

void FixedSizeFilterBits::AddKey(Slice key) {
  ++keys_added_;
  uint32_t h = BloomHash(key);
  const uint32_t delta = (h >> 17) | (h << 15);
  const size_t b = (h % num_lines_) * (kCacheLine * 8);
  for (size_t i = 0; i < num_probes_; ++i) {
    const size_t bitpos = b + (h % (kCacheLine * 8));
    data_[bitpos / 8] = static_cast<char>(data_[bitpos / 8] | (1 << (bitpos % 8)));
    h += delta;
  }
}

std::string FixedSizeFilterBits::Finish() {
  data_[total_bits_ / 8] = static_cast<char>(num_probes_);
  uint32_t nl = static_cast<uint32_t>(num_lines_);
  memcpy(&data_[total_bits_ / 8 + 1], &nl, 4);
  return data_;
}

bool FixedSizeFilterBits::MayMatch(Slice filter, Slice key) {      // bloom.cc:159-197,~330-360
  if (filter.n <= 5) return false;
  const size_t len = filter.n - 5;
  const size_t num_probes = filter.p[len];
  const uint32_t num_lines = DecodeFixed32(filter.p + len + 1);
  if (num_lines == 0 || len % num_lines != 0) return true;
  uint32_t h = BloomHash(key);
  const uint32_t delta = (h >> 17) | (h << 15);
  const size_t b = (h % num_lines) * (kCacheLine * 8);
  for (size_t i = 0; i < num_probes; ++i) {
    const size_t bitpos = b + (h % (kCacheLine * 8));
    if ((filter.p[bitpos / 8] & (1 << (bitpos % 8))) == 0) return false;
    h += delta;
  }
  return true;
}

size_t DocKeyV3FilterPrefix(Slice user_key) {
  try { return DocKeyEncodedSize(user_key, 2); } catch (const std::exception&) { return 0; }
}

// ---------------------------------------------------------------------------------------------
TableBuilder::TableBuilder(const TableOptions& o)
    : o_(o), data_block_(o.block_restart_interval, o.key_encoding, o.use_delta_encoding),
      policy_{o.block_size, static_cast<uint64_t>(o.block_size_deviation), 1, &data_block_},
      index_(new MultiLevelIndexBuilder(o)) {
  if (o.filter_policy) {
    // block_based_table_builder.cc:398-415,465-467: the filter builder is started right away; the
    // filter index is a plain binary-search index over bytewise-compared filter keys.
    filter_.reset(new FixedSizeFilterBits(static_cast<size_t>(o.filter_block_size) * 8, 0.01));
    filter_index_.reset(new BlockBuilder(o.index_block_restart_interval, kSharedPrefix));
  }
}
TableBuilder::~TableBuilder() {}

void TableBuilder::WriteRawBlock(Slice c, std::string* file, BlockHandle* h, bool compressible) {
  uint8_t type = 0;   // kNoCompression
  std::string compressed;
  // CompressBlock (block_based_table_builder.cc:115-131) for what goes through WriteBlock: data blocks, index blocks and
  // the filter index (:550,586,790,823,869); filter blocks, properties and the metaindex are written raw (:600,849,864)
  if (o_.compression == 1 && (file == &data_ || compressible)) {
    SnappyCompress(c, &compressed);
    if (compressed.size() < c.n - (c.n / 8u)) { c = Slice(compressed); type = 1; }     // GoodCompressionRatio :109-112
  }
  h->offset = file->size(); h->size = c.n;
  file->append(reinterpret_cast<const char*>(c.p), c.n);
  uint32_t crc = Crc32cExtend(Crc32cValue(c.p, c.n), &type, 1);
  file->push_back(static_cast<char>(type));
  PutFixed32(file, Crc32cMask(crc));
}

// block_based_table_builder.cc:594-620.
void TableBuilder::FlushFilterBlock(const Slice* next_block_first_filter_key) {
  std::string contents = filter_->Finish();
  WriteRawBlock(Slice(contents), &meta_, &filter_pending_);
  props_.filter_size += contents.size() + kBlockTrailerSize;
  ++props_.num_filter_blocks;
  if (next_block_first_filter_key) filter_.reset(new FixedSizeFilterBits(static_cast<size_t>(o_.filter_block_size) * 8, 0.01));
  // ShortenedIndexBuilder::AddIndexEntry with BytewiseComparator (index_builder.cc:60-86)
  if (!next_block_first_filter_key) BytewiseFindShortSuccessor(&last_filter_key_);
  else BytewiseFindShortestSeparator(&last_filter_key_, *next_block_first_filter_key);
  std::string enc; PutVarint64(&enc, filter_pending_.offset); PutVarint64(&enc, filter_pending_.size);
  filter_index_->Add(Slice(last_filter_key_), Slice(enc));
}

void TableBuilder::Add(Slice key, Slice value) {
  if (policy_.Update(key, value)) FlushDataBlock(key, true);
  if (filter_) {                                                   // block_based_table_builder.cc:514-528
    Slice user_key(key.p, key.n - 8);
    Slice filter_key(user_key.p, DocKeyV3FilterPrefix(user_key));
    if (!filter_key.empty() && (props_.num_entries == 0 || Slice(last_filter_key_).compare(filter_key) != 0)) {
      if (filter_->IsFull()) FlushFilterBlock(&filter_key);
      filter_->AddKey(filter_key);
      last_filter_key_ = filter_key.str();
    }
  }
  last_key_.assign(reinterpret_cast<const char*>(key.p), key.n);
  data_block_.Add(key, value);
  props_.num_entries++;
  props_.raw_key_size += key.n;
  props_.raw_value_size += value.n;
  uint8_t t = key.p[key.n - 8];
  if (t == kTypeDeletion || t == kTypeSingleDeletion) deleted_keys_++;
}

void TableBuilder::FlushDataBlock(Slice next_first_key, bool has_next) {
  if (!data_block_.empty()) {
    Slice c = data_block_.Finish();
    WriteRawBlock(c, &data_, &pending_);
    data_block_.Reset();
    data_handles_.push_back(pending_);
    props_.data_size += pending_.size + kBlockTrailerSize;
  }
  ++props_.num_data_blocks;
  index_->AddIndexEntry(&last_key_, has_next ? &next_first_key : nullptr, pending_, true);
  while (index_->ShouldFlush()) {
    std::string contents;
    bool r = index_->FlushNextBlock(&contents, last_index_handle_, last_index_handle_set_);
    if (!r) throw std::runtime_error("index flush returned false");
    WriteRawBlock(Slice(contents), &meta_, &last_index_handle_, true);
    last_index_handle_set_ = true;
    ++props_.num_data_index_blocks;
  }
}

void TableBuilder::Finish() {
  if (!data_block_.empty()) FlushDataBlock(Slice(), false);
  if (filter_) FlushFilterBlock(nullptr);
  closed_ = true;
  std::string top_index;
  bool have_top = index_->FlushNextBlock(&top_index, last_index_handle_, last_index_handle_set_);
  if (have_top) ++props_.num_data_index_blocks;

  // Properties block (meta_blocks.cc:67-135): BlockBuilder restart interval 1, keys sorted.
  std::map<std::string, std::string> p;
  auto addu = [&](const char* name, uint64_t v) { std::string s; PutVarint64(&s, v); p[name] = s; };
  props_.data_index_size = index_->EstimatedSize() + kBlockTrailerSize;
  addu("rocksdb.raw.key.size", props_.raw_key_size);
  addu("rocksdb.raw.value.size", props_.raw_value_size);
  addu("rocksdb.data.size", props_.data_size);
  addu("rocksdb.data.index.size", props_.data_index_size);
  addu("rocksdb.filter.index.size", props_.filter_index_size);
  addu("rocksdb.num.entries", props_.num_entries);
  addu("rocksdb.num.data.blocks", props_.num_data_blocks);
  addu("rocksdb.num.filter.blocks", props_.num_filter_blocks);
  addu("rocksdb.num.data.index.blocks", props_.num_data_index_blocks);
  addu("rocksdb.filter.size", props_.filter_size);
  addu("rocksdb.format.version", props_.format_version);
  addu("rocksdb.fixed.key.length", props_.fixed_key_len);
  // User-collected: InternalKeyPropertiesCollector (db/table_properties_collector.cc:46-57) and
  // BlockBasedTablePropertiesCollector (block_based_table_builder.cc:360-380).
  addu("rocksdb.deleted.keys", deleted_keys_);
  { std::string v; PutFixed32(&v, o_.multi_level_index ? 2 : 0); p["rocksdb.block.based.table.index.type"] = v; }
  p["rocksdb.block.based.table.whole.key.filtering"] = "1";
  p["rocksdb.block.based.table.prefix.filtering"] = "0";
  { std::string v; PutFixed32(&v, index_->NumLevels()); p["rocksdb.block.based.table.index.num.levels"] = v; }
  { std::string v(1, static_cast<char>(o_.key_encoding));
    p["rocksdb.block.based.table.data.block.key.value.encoding.format"] = v; }
  // Filter index block: written before the properties block (block_based_table_builder.cc:795-830).
  BlockHandle filter_index_handle;
  if (filter_) {
    Slice fi = filter_index_->Finish();
    WriteRawBlock(fi, &meta_, &filter_index_handle, true);
    props_.filter_index_size = filter_index_->CurrentSizeEstimate() + kBlockTrailerSize;
    p["rocksdb.filter.index.size"].clear(); PutVarint64(&p["rocksdb.filter.index.size"], props_.filter_index_size);
    p["rocksdb.filter.policy"] = "DocKeyV3Filter";
  }
  BlockBuilder pb(1, kSharedPrefix);
  for (auto& kv : p) pb.Add(Slice(kv.first), Slice(kv.second));
  BlockHandle props_handle;
  WriteRawBlock(pb.Finish(), &meta_, &props_handle);

  // Metaindex block.
  BlockBuilder mb(1, kSharedPrefix);
  { std::string enc; PutVarint64(&enc, props_handle.offset); PutVarint64(&enc, props_handle.size);
    // MetaIndexBuilder keeps a sorted map (meta_blocks.cc:71-83): "fixedsizefilter." < "rocksdb."
    if (filter_) {
      std::string fe; PutVarint64(&fe, filter_index_handle.offset); PutVarint64(&fe, filter_index_handle.size);
      mb.Add(Slice(std::string("fixedsizefilter.DocKeyV3Filter")), Slice(fe));
    }
    mb.Add(Slice(std::string("rocksdb.properties")), Slice(enc)); }
  BlockHandle metaindex_handle;
  WriteRawBlock(mb.Finish(), &meta_, &metaindex_handle);
  if (have_top) { WriteRawBlock(Slice(top_index), &meta_, &last_index_handle_, true); last_index_handle_set_ = true; }

  // Footer (format.cc:118-153), version 2, checksum type kCRC32c = 1.
  std::string f;
  f.push_back(1);
  PutVarint64(&f, metaindex_handle.offset); PutVarint64(&f, metaindex_handle.size);
  PutVarint64(&f, last_index_handle_.offset); PutVarint64(&f, last_index_handle_.size);
  f.resize(kFooterSize - 12);
  PutFixed32(&f, 2);
  PutFixed32(&f, static_cast<uint32_t>(kBlockBasedTableMagicNumber & 0xffffffffu));
  PutFixed32(&f, static_cast<uint32_t>(kBlockBasedTableMagicNumber >> 32));
  meta_.append(f);
}

// ---------------------------------------------------------------------------------------------
Slice TableReader::ReadBlock(Slice file, BlockHandle h, bool verify, std::string* scratch) {
  if (h.offset + h.size + kBlockTrailerSize > file.n) throw Corruption("truncated block read");
  const uint8_t* p = file.p + h.offset;
  if (p[h.size] > 1 || (p[h.size] == 1 && !scratch)) throw NotSupported("compressed block: only Snappy data blocks are restated in the oracle");
  if (verify) {      // the checksum covers the stored (compressed) bytes + the type byte (format.cc:352-395)
    uint32_t stored = Crc32cUnmask(DecodeFixed32(p + h.size + 1));
    uint32_t actual = Crc32cValue(p, h.size + 1);
    if (stored != actual) throw Corruption("block checksum mismatch");
  }
  if (p[h.size] == 1) { SnappyUncompress(Slice(p, h.size), scratch); return Slice(*scratch); }
  return Slice(p, h.size);
}

// ---- Snappy raw format ------------------------------------------------------------------------
void SnappyUncompress(Slice c, std::string* out) {
  const uint8_t* p = c.p; const uint8_t* e = c.p + c.n;
  uint32_t ulen = 0;
  p = GetVarint32Ptr(p, e, &ulen);
  if (!p) throw Corruption("snappy: bad length preamble");
  out->clear(); out->reserve(ulen);
  while (p < e) {
    const uint8_t tag = *p++;
    const uint32_t kind = tag & 3;
    if (kind == 0) {
      uint32_t len = tag >> 2;
      if (len >= 60) {
        const uint32_t nb = len - 59;
        if (static_cast<size_t>(e - p) < nb) throw Corruption("snappy: truncated literal length");
        len = 0;
        for (uint32_t i = 0; i < nb; i++) len |= static_cast<uint32_t>(p[i]) << (8 * i);
        p += nb;
      }
      len += 1;
      if (static_cast<size_t>(e - p) < len) throw Corruption("snappy: truncated literal");
      out->append(reinterpret_cast<const char*>(p), len);
      p += len;
    } else {
      uint32_t len, off;
      if (kind == 1) { if (p >= e) throw Corruption("snappy: truncated copy"); len = 4 + ((tag >> 2) & 7); off = (static_cast<uint32_t>(tag >> 5) << 8) | *p++; }
      else if (kind == 2) { if (e - p < 2) throw Corruption("snappy: truncated copy"); len = 1 + (tag >> 2); off = p[0] | (p[1] << 8); p += 2; }
      else { if (e - p < 4) throw Corruption("snappy: truncated copy"); len = 1 + (tag >> 2); off = DecodeFixed32(p); p += 4; }
      if (off == 0 || off > out->size()) throw Corruption("snappy: copy offset out of range");
      const size_t start = out->size() - off;
      for (uint32_t i = 0; i < len; i++) out->push_back((*out)[start + i]);       // may overlap its own output
    }
    if (out->size() > ulen) throw Corruption("snappy: output longer than announced");
  }
  if (out->size() != ulen) throw Corruption("snappy: output shorter than announced");
}

// The encoder every writer of this repository shares (this oracle, the engine's host writer and its GPU kernel), so
// that whole files can be compared byte for byte: the input is cut into 64 KB fragments like the snappy library does
// (a match never crosses a fragment); inside a fragment the classic greedy scan — hash the four bytes at every visited
// position into a table of 2^12 fragment-relative positions (zero-initialised; a stale or empty slot only costs a
// failed comparison), take the slot's previous occupant as the candidate, and on a four-byte match extend it as far
// as it goes; positions covered by a match are not visited. It is NOT the library's encoder (different table size and
// no skipping heuristic): compressed bytes are unpinned, the format is pinned by SnappyUncompress.
void SnappyCompress(Slice raw, std::string* out) {
  out->clear();
  PutVarint32(out, static_cast<uint32_t>(raw.n));
  constexpr size_t kFragment = 65536;
  constexpr uint32_t kHashBits = 12;
  std::vector<uint16_t> table(1u << kHashBits);
  for (size_t fs = 0; fs < raw.n; fs += kFragment) {
    const uint8_t* b = raw.p + fs; const size_t n = std::min(kFragment, raw.n - fs);
    auto emit_literal = [&](size_t from, size_t to) {
      if (from >= to) return;
      const size_t len = to - from, l1 = len - 1;                       // <= 65536
      if (l1 < 60) out->push_back(static_cast<char>(l1 << 2));
      else if (l1 < 256) { out->push_back(static_cast<char>(60 << 2)); out->push_back(static_cast<char>(l1)); }
      else { out->push_back(static_cast<char>(61 << 2)); out->push_back(static_cast<char>(l1 & 0xff)); out->push_back(static_cast<char>(l1 >> 8)); }
      out->append(reinterpret_cast<const char*>(b + from), len);
    };
    std::fill(table.begin(), table.end(), 0);
    size_t lit = 0, i = 0;
    while (i + 4 <= n) {
      uint32_t w; memcpy(&w, b + i, 4);
      const uint32_t hsh = (w * 0x1e35a7bdu) >> (32 - kHashBits);
      const size_t cand = table[hsh];
      table[hsh] = static_cast<uint16_t>(i);
      uint32_t cw = 0;
      if (cand < i) memcpy(&cw, b + cand, 4);
      if (cand < i && cw == w) {
        size_t len = 4;
        while (i + len < n && b[cand + len] == b[i + len]) len++;
        emit_literal(lit, i);
        const uint32_t off = static_cast<uint32_t>(i - cand);
        size_t left = len;
        while (left) {
          size_t l = std::min<size_t>(left, 64);
          if (left - l > 0 && left - l < 4) l = left - 4;                 // every piece is at least 4 bytes long
          if (l >= 4 && l <= 11 && off < 2048) { out->push_back(static_cast<char>(1 | ((l - 4) << 2) | ((off >> 8) << 5))); out->push_back(static_cast<char>(off & 0xff)); }
          else { out->push_back(static_cast<char>(2 | ((l - 1) << 2))); out->push_back(static_cast<char>(off & 0xff)); out->push_back(static_cast<char>(off >> 8)); }
          left -= l;
        }
        i += len; lit = i;
      } else {
        i++;
      }
    }
    emit_literal(lit, n);
  }
}

static BlockHandle DecodeHandle(Slice* s) {
  BlockHandle h;
  const uint8_t* p = GetVarint64Ptr(s->p, s->p + s->n, &h.offset);
  if (p) p = GetVarint64Ptr(p, s->p + s->n, &h.size);
  if (!p) throw Corruption("bad block handle");
  s->remove_prefix(p - s->p);
  return h;
}

void TableReader::Open(Slice meta_file, Slice data_file, bool verify) {
  meta = meta_file; data = data_file;
  if (meta.n < kFooterSize) throw Corruption("file is too short to be an sstable");
  const uint8_t* f = meta.p + meta.n - kFooterSize;
  uint64_t magic = static_cast<uint64_t>(DecodeFixed32(f + kFooterSize - 8)) |
                   (static_cast<uint64_t>(DecodeFixed32(f + kFooterSize - 4)) << 32);
  if (magic != kBlockBasedTableMagicNumber) throw Corruption("bad table magic number");
  Slice s(f + 1, 40);
  BlockHandle metaindex = DecodeHandle(&s);
  BlockHandle index = DecodeHandle(&s);
  // metaindex -> properties
  {
    BlockIter it(ReadBlock(meta, metaindex, verify), kSharedPrefix);
    for (it.SeekToFirst(); it.Valid(); it.Next()) {
      if (it.key().str() == "rocksdb.properties") {
        Slice v = it.value();
        BlockHandle ph = DecodeHandle(&v);
        BlockIter pit(ReadBlock(meta, ph, verify), kSharedPrefix);
        for (pit.SeekToFirst(); pit.Valid(); pit.Next()) properties[pit.key().str()] = pit.value().str();
      } else if (it.key().str().rfind("fixedsizefilter.", 0) == 0) {
        Slice v = it.value();
        BlockHandle fh = DecodeHandle(&v);
        std::string fscratch;
        BlockIter fit(ReadBlock(meta, fh, verify, &fscratch), kSharedPrefix);
        for (fit.SeekToFirst(); fit.Valid(); fit.Next()) { Slice hv = fit.value(); filter_blocks.emplace_back(fit.key().str(), DecodeHandle(&hv)); }
      }
    }
  }
  auto enc = properties.find("rocksdb.block.based.table.data.block.key.value.encoding.format");
  key_encoding = (enc == properties.end() || enc->second.empty()) ? static_cast<int>(kSharedPrefix) : static_cast<int>(static_cast<uint8_t>(enc->second[0]));
  auto lv = properties.find("rocksdb.block.based.table.index.num.levels");
  num_index_levels = (lv == properties.end() || lv->second.size() < 4)
                         ? 1 : DecodeFixed32(reinterpret_cast<const uint8_t*>(lv->second.data()));
  // Walk the multi-level index top-down.
  std::vector<BlockHandle> level{index};
  for (int l = 0; l < num_index_levels; l++) {
    std::vector<BlockHandle> next;
    for (auto& h : level) {
      std::string iscratch;                  // index blocks are stored compressed when the table's blocks are
      BlockIter it(ReadBlock(meta, h, verify, &iscratch), kSharedPrefix);
      for (it.SeekToFirst(); it.Valid(); it.Next()) { Slice v = it.value(); next.push_back(DecodeHandle(&v)); }
    }
    level.swap(next);
  }
  data_blocks = level;
}

}  // namespace orc
