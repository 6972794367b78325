// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_codec.h).
//
// CPU restatement of the compaction hot loop: SST iterators -> MergingIterator (binary heap) ->
// CompactionIterator (rule A + seqno zeroing) -> DocDBCompactionFeed (MVCC retention) ->
// TableBuilder. One thread per compaction, like the reference (rocksdb/util/options.cc:258).
#pragma once
#include <utility>
#include <vector>
#include "oracle_sst.h"
#include <functional>

namespace orc {

// docdb/docdb_compaction_context.h:57-111,178-196 flattened.
struct RetentionParams {
  bool enabled = true;                       // false = no compaction_context_factory (plain RocksDB)
  uint64_t primary_cutoff_ht = kHtMin;       // HistoryCutoff::primary_cutoff_ht (HybridTime repr)
  uint64_t cotables_cutoff_ht = kHtInvalid;  // invalid = not set
  int64_t table_ttl_ns = kMaxTtlNs;          // MonoDelta; kMaxTtl = no table TTL
  bool retain_delete_markers_in_major_compaction = false;
  uint64_t other_min_ht = kHtMax;            // CompactionHybridTimeConstraints::other_min; kMax = "major"
  std::string lower_bound, upper_bound;      // KeyBounds (empty = unbounded)
};

struct CompactionParams {
  bool bottommost_level = true;              // Compaction::bottommost_level()
  uint64_t last_sequence = kMaxSequenceNumber;  // VersionSet::LastSequence() => earliest_snapshot_
  std::string largest_user_key;              // Compaction::GetLargestUserKey(); computed if !has_largest
  bool has_largest_user_key = false;
  RetentionParams retention;
};

struct CompactionStats {
  uint64_t num_input_records = 0, num_output_records = 0;
  uint64_t num_dropped_hidden = 0;      // rule A (compaction_iterator.cc:388-400)
  uint64_t num_dropped_obsolete = 0;    // kTypeDeletion at bottommost
  uint64_t num_dropped_feed = 0;        // dropped by DocDBCompactionFeed
  uint64_t total_input_raw_key_bytes = 0, total_input_raw_value_bytes = 0;
  uint64_t total_output_raw_key_bytes = 0, total_output_raw_value_bytes = 0;
  // FileMetaData::smallest / largest .user_values as DocDBCompactionFeed accumulates them
  // (docdb_compaction_context.cc:754-773 + doc_boundary_values_extractor.cc:40-64): per range-group component
  // index of the surviving DocKeys, the bytewise smallest / largest encoded component (tag = 10 + index).
  std::vector<std::pair<uint32_t, std::string>> smallest_user_values, largest_user_values;
};

// rocksdb/db/compaction_context.h:25-35.
struct CompactionFeed {
  virtual ~CompactionFeed() {}
  virtual void Feed(Slice internal_key, Slice value) = 0;
  virtual void Flush() = 0;
};

// hybrid_time_filter: the file's global filter; cotable_filters: (database oid, hybrid time) sorted by oid — the tail of
// FdWithBoundaries::user_filter_data (docdb/docdb_rocksdb_util.cc:503-509).
struct SstInput {
  Slice meta, data;
  uint64_t hybrid_time_filter = kHtInvalid;
  std::vector<std::pair<uint32_t, uint64_t>> cotable_filters;
};

// Runs the whole loop (rocksdb/db/compaction_job.cc:664-895). `sink` receives every surviving
// (internal key, value) in output order.
void RunCompaction(const std::vector<SstInput>& inputs, const CompactionParams& params,
                   CompactionFeed* sink, CompactionStats* stats, bool verify_checksums = true);

// Same, but over already-decoded sorted runs (used by the retention golden tests, which the
// reference drives through mock tables / tiny SSTs).
struct KvRun { std::vector<std::pair<std::string, std::string>> kv; };
void RunCompactionOnRuns(const std::vector<KvRun>& runs, const CompactionParams& params,
                         CompactionFeed* sink, CompactionStats* stats);

// ---- docdb/compaction_file_filter.{h,cc}: whole-file expiration by TTL (CompactionFileFilterFactory; the picker marks
// the files it discards delete_after_compaction, compaction_picker.cc:476-492, and MakeInputIterator leaves them out,
// db/version_set.cc:3812-3820). HybridTimes are reprs; kNoExpiration = kMax, kUseDefaultTTL = kInitial
// (dockv/doc_ttl_util.h:65-73).
enum ExpiryMode { EXP_NORMAL = 0, EXP_TABLE_ONLY = 1, EXP_TRUST_VALUE = 2 };     // compaction_file_filter.h:26-30
constexpr uint64_t kNoExpiration = kHtMax, kUseDefaultTTL = kHtMin + 1;
struct ExpirationTime {                                                         // compaction_file_filter.h:32-43
  uint64_t ttl_expiration_ht = kNoExpiration;   // the largest frontier's max_value_level_ttl_expiration_time
  uint64_t created_ht = kHtMax;                 // the largest frontier's hybrid_time
};
// ExtractExpirationTime (:70-85) from the two frontier fields; `has_frontier` false = no largest user frontier.
ExpirationTime ExtractExpirationTime(bool has_frontier, uint64_t frontier_ht, uint64_t max_value_level_ttl_expiration_ht);
uint64_t ComputeExpiration(uint64_t ht, int64_t ttl_ns);                                        // doc_ttl_util.cc:81-87
uint64_t MaxExpirationFromValueAndTableTTL(uint64_t key_ht, int64_t table_ttl_ns, uint64_t value_expiry);   // :107-129
bool HasExpiredTTL(uint64_t expiration_ht, uint64_t read_ht);                                   // :42-47
bool TtlIsExpired(ExpirationTime expiry, int64_t table_ttl_ns, uint64_t now, ExpiryMode mode);  // compaction_file_filter.cc:126-144
// DocDBCompactionFileFilterFactory::CreateCompactionFileFilter over all input files, then ::Filter of each (:150-243):
// true = kDiscard.
std::vector<bool> FileFilterDecisions(const std::vector<ExpirationTime>& files, int64_t table_ttl_ns, uint64_t primary_cutoff_ht,
                                      uint64_t cotables_cutoff_ht, uint64_t now, ExpiryMode mode);

}  // namespace orc
