// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_codec.h).
//
// CPU restatement of the RocksDB-fork SST block format on the compaction path: BlockBuilder /
// BlockIter (both key-delta encodings), flush policy, BlockBasedTableBuilder (split data / meta
// files, multi-level index, properties, metaindex, footer) and the reader used by compaction.
#pragma once
#include "oracle_codec.h"
#include <map>
#include <memory>

namespace orc {

constexpr size_t kBlockTrailerSize = 5;                     // table/format.h:208
constexpr uint64_t kBlockBasedTableMagicNumber = 0x88e241b785f4cff7ull;  // block_based_table_builder.cc:195
constexpr size_t kFooterSize = 1 + 40 + 4 + 8;              // table/format.h:170
constexpr size_t kLastInternalComponentSize = 8;

enum KeyEncoding { kSharedPrefix = 1, kThreeSharedParts = 2 };   // rocksdb/types.h:50-56
enum ValueType : uint8_t { kTypeDeletion = 0, kTypeValue = 1, kTypeMerge = 2, kTypeSingleDeletion = 7 };
constexpr uint64_t kMaxSequenceNumber = (1ull << 56) - 1;
constexpr uint8_t kValueTypeForSeek = kTypeSingleDeletion;

inline uint64_t PackSeqAndType(uint64_t seq, uint8_t t) { return (seq << 8) | t; }   // dbformat.cc:42-46

// db/dbformat.cc:92-114 InternalKeyComparator::Compare.
inline int CompareInternalKey(Slice a, Slice b) {
  Slice ua(a.p, a.n - 8), ub(b.p, b.n - 8);
  int r = ua.compare(ub);
  if (r == 0) {
    uint64_t an = DecodeFixed64(a.p + a.n - 8), bn = DecodeFixed64(b.p + b.n - 8);
    if (an > bn) r = -1; else if (an < bn) r = 1;
  }
  return r;
}

struct BlockHandle { uint64_t offset = 0, size = 0; };

struct TableOptions {
  uint32_t block_size = 32 * 1024;          // db_block_size_bytes (dockv/packed_row.cc:39)
  int block_restart_interval = 16;          // docdb_rocksdb_util.cc:74,188
  int index_block_restart_interval = 1;     // table.h:153
  int block_size_deviation = 10;            // table.h:144
  uint32_t index_block_size = 32 * 1024;    // docdb_rocksdb_util.cc:132-136
  uint32_t min_keys_per_index_block = 100;
  int key_encoding = kSharedPrefix;
  bool use_delta_encoding = true;
  bool multi_level_index = true;            // docdb_rocksdb_util.cc:772-773
  // 0 = no filter policy; 1 = DocDbAwareV3FilterPolicy ("DocKeyV3Filter", docdb_filter_policy.h:71-80):
  // fixed-size bloom blocks of filter_block_size bytes, error rate 0.01 (filter_policy.h:179-180),
  // keyed by the DocKey up to its hashed components / first range component
  // (docdb_rocksdb_util.cc:761-763, docdb_filter_policy.cc:103-108).
  int filter_policy = 0;
  uint32_t filter_block_size = 64 * 1024;   // db_filter_block_size_bytes (docdb_rocksdb_util.cc:132)
  // rocksdb::CompressionType of the DATA blocks: 0 = kNoCompression, 1 = kSnappyCompression (the production default,
  // docdb_rocksdb_util.cc:184). A block is stored compressed only when that saves at least 12.5 %
  // (block_based_table_builder.cc:109-131). Test inputs only: the compressed BYTES are this file's own encoder's, not
  // the snappy library's (parity unpinned for them); what is pinned is the format, i.e. that they decompress back.
  int compression = 0;
};

// Snappy raw format (format_description.txt of the snappy library, third_party: yugabyte-db-thirdparty, not vendored):
// varint32 uncompressed length, then literal / copy elements. Compress emits a valid (greedy, hash-based) encoding;
// Uncompress accepts any valid stream and throws Corruption otherwise.
void SnappyCompress(Slice raw, std::string* out);
void SnappyUncompress(Slice compressed, std::string* out);

// util/hash.cc:32-75 (the LevelDB hash) and util/hash.h:40-42.
uint32_t LevelDbHash(const uint8_t* data, size_t n, uint32_t seed);
inline uint32_t BloomHash(Slice key) { return LevelDbHash(key.p, key.n, 0xbc9f1d34u); }

// util/bloom.cc:384-455 FixedSizeFilterBitsBuilder + :43-61 AddHash; reader side :159-197,
// FullFilterBitsReader::HashMayMatch.
class FixedSizeFilterBits {
 public:
  FixedSizeFilterBits(size_t total_bits, double error_rate);
  void AddKey(Slice key);
  bool IsFull() const { return keys_added_ >= max_keys_; }
  std::string Finish();                       // filter data + num_probes byte + fixed32 num_lines
  size_t max_keys() const { return max_keys_; }
  size_t num_lines() const { return num_lines_; }
  size_t num_probes() const { return num_probes_; }
  static bool MayMatch(Slice filter, Slice key);
 private:
  std::string data_;
  size_t max_keys_, keys_added_ = 0, total_bits_, num_lines_, num_probes_;
};

// FilterPolicy::KeyTransformer of DocDbAwareV3FilterPolicy (docdb_filter_policy.cc:27-47,103-108):
// length of the filter key inside `user_key`, 0 when the key does not decode as a DocKey.
size_t DocKeyV3FilterPrefix(Slice user_key);

// table/block_builder.cc:63-412.
class BlockBuilder {
 public:
  BlockBuilder(int restart_interval, int key_encoding, bool use_delta = true)
      : restart_interval_(restart_interval), enc_(key_encoding), use_delta_(use_delta) { restarts_.push_back(0); }
  void Reset() { buf_.clear(); restarts_.assign(1, 0); counter_ = 0; finished_ = false; last_key_.clear(); }
  void Add(Slice key, Slice value);
  Slice Finish();
  size_t CurrentSizeEstimate() const {
    return buf_.size() + (finished_ ? 0 : restarts_.size() * 4 + 4);
  }
  size_t EstimateSizeAfterKV(Slice key, Slice value) const {
    size_t e = CurrentSizeEstimate() + key.n + value.n;
    if (counter_ >= restart_interval_) e += 4;
    e += 4 + VarintLength(key.n) + VarintLength(value.n);
    return e;
  }
  bool empty() const { return buf_.empty(); }
  size_t NumKeys() const { return restarts_.size() * restart_interval_ + counter_; }  // sic (:112-114)
 private:
  int restart_interval_, enc_;
  bool use_delta_;
  std::string buf_;
  std::vector<uint32_t> restarts_;
  int counter_ = 0;
  bool finished_ = false;
  std::string last_key_;
};

// table/block.cc:348-447 BlockIter (forward only) for both encodings.
class BlockIter {
 public:
  BlockIter(Slice block, int key_encoding);
  bool Valid() const { return valid_; }
  void SeekToFirst();
  void Next();
  Slice key() const { return Slice(key_); }
  Slice value() const { return value_; }
  uint32_t num_restarts() const { return num_restarts_; }
 private:
  bool Parse();
  const uint8_t* data_;
  uint32_t restarts_off_, num_restarts_;
  int enc_;
  uint32_t next_ = 0;
  std::string key_;
  Slice value_;
  bool valid_ = false;
};

// table/flush_block_policy.cc:29-92.
struct FlushBySize {
  uint64_t block_size, deviation; size_t min_keys; const BlockBuilder* b;
  bool Update(Slice key, Slice value) const {
    if (b->empty()) return false;
    size_t cur = b->CurrentSizeEstimate();
    bool almost = b->EstimateSizeAfterKV(key, value) > block_size && deviation > 0 &&
                  cur * 100 > block_size * (100 - deviation);
    return (cur >= block_size || almost) && b->NumKeys() >= min_keys;
  }
};

// util/comparator.cc:53-93 + db/dbformat.cc:139-172.
void BytewiseFindShortestSeparator(std::string* start, Slice limit);
void BytewiseFindShortSuccessor(std::string* key);
void InternalFindShortestSeparator(std::string* start, Slice limit);
void InternalFindShortSuccessor(std::string* key);

struct TableProps {
  uint64_t raw_key_size = 0, raw_value_size = 0, data_size = 0, data_index_size = 0,
           filter_index_size = 0, num_entries = 0, num_data_blocks = 0, num_filter_blocks = 0,
           num_data_index_blocks = 0, filter_size = 0, format_version = 0, fixed_key_len = 0;
};

// table/index_builder.cc:143-289 MultiLevelIndexBuilder.
class MultiLevelIndexBuilder;

// table/block_based_table_builder.cc:498-903, split-SST mode (data blocks -> data file,
// everything else -> meta file), kNoCompression, CRC32c, no filter policy.
class TableBuilder {
 public:
  explicit TableBuilder(const TableOptions& o);
  ~TableBuilder();
  void Add(Slice ikey, Slice value);
  void Finish();
  const std::string& data_file() const { return data_; }
  const std::string& meta_file() const { return meta_; }
  const TableProps& props() const { return props_; }
  uint64_t NumEntries() const { return props_.num_entries; }
  uint64_t TotalFileSize() const { return data_.size() + meta_.size(); }
  const std::vector<BlockHandle>& data_block_handles() const { return data_handles_; }
 private:
  void FlushDataBlock(Slice next_first_key, bool has_next);
  void WriteRawBlock(Slice contents, std::string* file, BlockHandle* h, bool compressible = false);
  TableOptions o_;
  BlockBuilder data_block_;
  FlushBySize policy_;
  std::unique_ptr<MultiLevelIndexBuilder> index_;
  std::string last_key_;
  std::string data_, meta_;
  TableProps props_;
  void FlushFilterBlock(const Slice* next_block_first_filter_key);
  std::unique_ptr<FixedSizeFilterBits> filter_;
  std::unique_ptr<BlockBuilder> filter_index_;
  std::string last_filter_key_;
  BlockHandle filter_pending_;
  BlockHandle pending_, last_index_handle_;
  bool last_index_handle_set_ = false;
  std::vector<BlockHandle> data_handles_;
  uint64_t deleted_keys_ = 0;
  bool closed_ = false;
};

// Reader side (table/block_based_table_reader.cc:1494-1673, format.cc:340-500, index_reader.h).
struct TableReader {
  Slice meta, data;
  int key_encoding = kSharedPrefix;
  int num_index_levels = 1;
  std::map<std::string, std::string> properties;
  std::vector<BlockHandle> data_blocks;      // in file order, from walking the index
  // fixed-size bloom filter: (index key, block handle) of every filter block, from the filter index
  // found under "fixedsizefilter.<policy>" in the metaindex (block_based_table_reader.cc:805-830)
  std::vector<std::pair<std::string, BlockHandle>> filter_blocks;
  void Open(Slice meta_file, Slice data_file, bool verify_checksums = true);
  // Returns block contents (without trailer); verifies CRC if asked.
  // scratch: receives the contents of a compressed block (format.cc:441-500 UncompressBlockContents); without it a
  // compressed block is an error
  static Slice ReadBlock(Slice file, BlockHandle h, bool verify, std::string* scratch = nullptr);
};

}  // namespace orc
