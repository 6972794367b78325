// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_codec.h).
#include "oracle_compaction.h"

namespace orc {

// ---------------------------------------------------------------------------------------------
// Input iterators. table/two_level_iterator.cc + BlockIter: walk data blocks in index order.
struct InputIter {
  virtual ~InputIter() {}
  virtual bool Valid() const = 0;
  virtual void Next() = 0;
  virtual Slice key() const = 0;
  virtual Slice value() const = 0;
};

struct SstIter : InputIter {
  TableReader r;
  size_t blk = 0;
  std::unique_ptr<BlockIter> it;
  std::string scratch;                 // contents of the current block when it is stored compressed
  bool verify;
  uint64_t ht_filter;
  std::vector<std::pair<uint32_t, uint64_t>> cotable_filters;
  SstIter(const SstInput& in, bool verify_) : verify(verify_), ht_filter(in.hybrid_time_filter), cotable_filters(in.cotable_filters) {
    r.Open(in.meta, in.data, verify);
    blk = 0;
    Load();
    SkipFiltered();
  }
  void Load() {
    it.reset();
    while (blk < r.data_blocks.size()) {
      it.reset(new BlockIter(TableReader::ReadBlock(r.data, r.data_blocks[blk], verify, &scratch), r.key_encoding));
      it->SeekToFirst();
      if (it->Valid()) return;
      blk++;
    }
    it.reset();
  }
  void Advance() {
    it->Next();
    if (!it->Valid()) { blk++; Load(); }
  }
  // docdb/docdb_rocksdb_util.cc:525-565 HybridTimeFilteringIterator::Satisfied: the logical AND of the file's global
  // filter and, for keys of a cotable, the filter of the cotable's database (the database oid is bytes 12..15 of the
  // table uuid, little endian, :548-551; Uuid::FromComparable leaves that half of the uuid in place, util/uuid.cc:162-178).
  bool Satisfied(Slice uk) const {
    uint64_t ht; uint32_t wid;
    size_t sz;
    try {
      sz = DocHtEncodedSizeFromEnd(uk);
      DecodeDocHt(Slice(uk.p + uk.n - sz, sz), &ht, &wid);
    } catch (const Corruption&) { return true; }
    if (ht_filter != kHtInvalid && ht > ht_filter) return false;
    if (cotable_filters.empty()) return true;
    const size_t key_len = uk.n - sz;                    // user key without the DocHybridTime
    if (key_len < 1 || uk.p[0] != 'y') return true;      // kTableId
    if (key_len < 17) return true;
    const uint8_t* u = reinterpret_cast<const uint8_t*>(uk.p) + 1;
    const uint32_t db_oid = static_cast<uint32_t>(u[12]) | (static_cast<uint32_t>(u[13]) << 8) | (static_cast<uint32_t>(u[14]) << 16) |
                            (static_cast<uint32_t>(u[15]) << 24);
    for (const auto& f : cotable_filters)
      if (f.first == db_oid) return ht <= f.second;
    return true;
  }
  void SkipFiltered() {
    if (ht_filter == kHtInvalid && cotable_filters.empty()) return;
    while (it) {
      Slice k = it->key();
      if (Satisfied(Slice(k.p, k.n - 8))) return;
      Advance();
    }
  }
  bool Valid() const override { return it != nullptr; }
  void Next() override { Advance(); SkipFiltered(); }
  Slice key() const override { return it->key(); }
  Slice value() const override { return it->value(); }
};

struct RunIter : InputIter {
  const KvRun* run; size_t i = 0;
  explicit RunIter(const KvRun* r) : run(r) {}
  bool Valid() const override { return i < run->kv.size(); }
  void Next() override { i++; }
  Slice key() const override { return Slice(run->kv[i].first); }
  Slice value() const override { return Slice(run->kv[i].second); }
};

// ---------------------------------------------------------------------------------------------
// table/merger.cc:251-698 + util/heap.h:54-192: binary min-heap of child iterators ordered by
// InternalKeyComparator; Next() advances the top child and sifts it down ("replace top"), with
// the reference's fast path of first comparing against the cached smaller root child.
class MergingIterator {
 public:
  explicit MergingIterator(std::vector<std::unique_ptr<InputIter>>* children) {
    for (auto& c : *children) if (c->Valid()) heap_.push_back(c.get());
    for (size_t i = heap_.size(); i-- > 0;) SiftDown(i, 0);
  }
  bool Valid() const { return !heap_.empty(); }
  Slice key() const { return heap_[0]->key(); }
  Slice value() const { return heap_[0]->value(); }
  void Next() {
    InputIter* cur = heap_[0];
    cur->Next();
    const size_t n = heap_.size();
    if (!cur->Valid()) {                      // merger.cc:691-696: pop and re-establish the top
      heap_[0] = heap_.back();
      heap_.pop_back();
      best_child_ = 0;
      if (heap_.size() > 1) SiftDown(0, 0);
      return;
    }
    if (n == 1) return;
    if (!best_child_) best_child_ = (n > 2 && Less(2, 1)) ? 2 : 1;
    // merger.cc:669-677 fast path: still smaller than the cached best root child.
    if (CompareInternalKey(cur->key(), heap_[best_child_]->key()) < 0) return;
    SiftDown(0, best_child_);
    best_child_ = 0;
  }
 private:
  bool Less(size_t a, size_t b) const { return CompareInternalKey(heap_[a]->key(), heap_[b]->key()) < 0; }
  // util/heap.h:135-180 down_root. `first_child` (if non-zero) is the already-known smaller child
  // of node i, known to be <= the sifted value.
  void SiftDown(size_t i, size_t first_child) {
    const size_t n = heap_.size();
    InputIter* v = heap_[i];
    if (first_child) { heap_[i] = heap_[first_child]; i = first_child; }
    for (;;) {
      size_t l = 2 * i + 1, r = l + 1;
      if (l >= n) break;
      size_t m = (r < n && Less(r, l)) ? r : l;
      if (CompareInternalKey(heap_[m]->key(), v->key()) >= 0) break;
      heap_[i] = heap_[m];
      i = m;
    }
    heap_[i] = v;
  }
  std::vector<InputIter*> heap_;
  size_t best_child_ = 0;
};

// ---------------------------------------------------------------------------------------------
// docdb/docdb_compaction_context.cc:643-1317 DocDBCompactionFeed, restricted to what a schema-
// less caller can reach: no packed rows, no deleted-column list, no vector-index metadata
// filter (those need SchemaPackingProvider / tablet callbacks and throw NotSupported if met).
struct Expiration {          // dockv/expiration.h
  int64_t ttl_ns = kMaxTtlNs;
  uint64_t write_ht = kHtMin;
};
struct OverwriteData { EncodedDocHt ht; Expiration exp; };

// common/hybrid_time.cc:172-195
static int CompareHybridTimesToDelta(uint64_t begin, uint64_t end, int64_t delta_ns) {
  if (end < begin) return -1;
  uint64_t bn = HtMicros(begin) * 1000, en = HtMicros(end) * 1000, dn = static_cast<uint64_t>(delta_ns);
  if (en - bn > dn) return 1;
  if (en - bn == dn) {
    uint64_t bl = HtLogical(begin), el = HtLogical(end);
    return el > bl ? 1 : (el < bl ? -1 : 0);
  }
  return -1;
}
// dockv/doc_ttl_util.cc:25-31,63-75
static bool HasExpiredTTL(uint64_t key_ht, int64_t ttl_ns, uint64_t read_ht) {
  if (ttl_ns == kMaxTtlNs || ttl_ns == 0) return false;
  return CompareHybridTimesToDelta(key_ht, read_ht, ttl_ns) > 0;
}
static int64_t ComputeTTL(int64_t value_ttl, int64_t default_ttl) {
  if (value_ttl != kMaxTtlNs) return (value_ttl / 1000000 == 0) ? kMaxTtlNs : value_ttl;
  return default_ttl;
}

class DocDBFeed : public CompactionFeed {
 public:
  DocDBFeed(CompactionFeed* next, const RetentionParams& r, uint64_t* dropped)
      : next_(next), r_(r), dropped_(dropped),
        cutoff_primary_(r.primary_cutoff_ht, kMaxWriteId),
        min_other_(r.retain_delete_markers_in_major_compaction ? kHtMin : r.other_min_ht, kMinWriteId),
        ht_min_(kHtMin, 0) {
    if (r.cotables_cutoff_ht != kHtInvalid) { has_cotables_cutoff_ = true; cutoff_cotables_ = EncodedDocHt(r.cotables_cutoff_ht, kMaxWriteId); }
  }
  void Flush() override { next_->Flush(); }

  void Feed(Slice internal_key, Slice value) override {
    bool kept = FeedImpl(internal_key, value);
    if (!kept) ++*dropped_;
  }

 private:
  bool CanHaveOtherDataBefore(const EncodedDocHt& ht) const { return CompareEncHt(ht, min_other_) >= 0; }
  const Expiration& LastExpiration() const {
    static const Expiration kDefault;
    return ow_.empty() ? kDefault : ow_.back().exp;
  }
  // PassToNextFeed (docdb_compaction_context.cc:754-762) + UpdateBoundaryValues (:764-773): the first entry passed on
  // for every DocKey (doc_key_serial_) contributes its range components to the per-tag minima / maxima.
  bool Forward(Slice ikey, Slice value) {
    if (last_passed_serial_ != doc_key_serial_) {
      UpdateBoundaryValues(Slice(ikey.p, ikey.n - 8));
      last_passed_serial_ = doc_key_serial_;
    }
    next_->Feed(ikey, value);
    return true;
  }
  void UpdateBoundaryValues(Slice user_key) {
    const uint8_t t = user_key.empty() ? kt::kInvalid : user_key[0];
    // IsMetaKeyType (dockv/value_type.h:252-273): internal DocDB records are skipped
    if (t == kt::kVectorIndexMetadata || t == kt::kTransactionApplyState || t == kt::kExternalTransactionId || t == kt::kTransactionId) return;
    std::vector<Slice> comps;
    DocKeyRangeComponents(user_key, &comps);
    for (size_t i = 0; i < comps.size(); i++) {
      const uint32_t tag = 10 + static_cast<uint32_t>(i);            // TagForRangeComponent
      auto upd = [&](std::vector<std::pair<uint32_t, std::string>>* dst, int sign) {   // rocksdb/db/metadata.cc:44-57
        for (auto& v : *dst)
          if (v.first == tag) { if (Slice(v.second).compare(comps[i]) * sign > 0) v.second = comps[i].str(); return; }
        dst->emplace_back(tag, comps[i].str());
      };
      upd(&smallest_, 1); upd(&largest_, -1);
    }
  }
 public:
  std::vector<std::pair<uint32_t, std::string>> smallest_, largest_;
 private:

  bool FeedImpl(Slice internal_key, Slice value) {
    Slice key(internal_key.p, internal_key.n - 8);
    const uint8_t key_type = key.empty() ? kt::kInvalid : key[0];
    const bool is_meta = key_type == kt::kVectorIndexMetadata || key_type == kt::kTransactionApplyState;
    const bool is_sub_doc_key = !is_meta;
    if (key_type == kt::kObsoleteIntentPrefix) return false;                       // :951
    if (is_sub_doc_key) {                                                          // :955
      bool within = (r_.lower_bound.empty() || key.compare(Slice(r_.lower_bound)) >= 0) &&
                    (r_.upper_bound.empty() || key.compare(Slice(r_.upper_bound)) < 0);
      if (!within) return false;
    }
    if (key_type == kt::kVectorIndexMetadata)
      throw NotSupported("vector index metadata keys need the tablet's VectorMetadataFilter");

    // :972 MemoryDifferencePos over min(len(key), len(prev_key_))
    size_t same = 0;
    { size_t m = std::min(key.n, prev_key_.size());
      while (same < m && key.p[same] == static_cast<uint8_t>(prev_key_[same])) same++; }
    size_t shared;                                                                 // :977-989
    if (!same) shared = 0;
    else { shared = ends_.size(); while (shared > 0 && ends_[shared - 1] > same) --shared; }
    if (shared < (is_sub_doc_key ? 2u : 1u)) ++doc_key_serial_;                    // :999-1003
    ends_.resize(shared);
    if (is_sub_doc_key) {
      DecodeDocKeyAndSubKeyEnds(key, &ends_);                                      // :1008
    } else {
      // DecodeMetaSubKeyEnds :921-937 (kTransactionApplyState => DocKey::EncodedSize whole)
      if (ends_.empty()) ends_.push_back(DocKeyEncodedSize(key, 1));
    }
    const size_t new_stack = ends_.size();
    if (shared < ow_.size()) ow_.resize(shared);                                   // :1021
    size_t ht_size = DocHtEncodedSizeFromEnd(key);                                 // :1026
    EncodedDocHt ht(Slice(key.p + key.n - ht_size, ht_size));
    EncodedDocHt prev_ow = ow_.empty() ? ht_min_ : ow_.back().ht;                  // :1048
    const bool is_ttl_row = !value.empty() && value[0] == kt::kMergeFlags;         // :1066
    if (CompareEncHt(ht, prev_ow) < 0 && !is_ttl_row) return false;                // :1067-1074
    if (ow_.size() < new_stack - 1) ow_.resize(new_stack - 1, OverwriteData{prev_ow, LastExpiration()});  // :1078
    Expiration popped = ow_.empty() ? Expiration() : ow_.back().exp;               // :1083
    if (ow_.size() == new_stack) ow_.pop_back();                                   // :1087
    if (same != ends_.back()) within_merge_block_ = false;                         // :1092

    uint64_t chosen_ht = r_.primary_cutoff_ht;                                     // :1103-1114
    const EncodedDocHt* chosen = &cutoff_primary_;
    if (key_type == kt::kTableId && has_cotables_cutoff_) { chosen = &cutoff_cotables_; chosen_ht = r_.cotables_cutoff_ht; }

    if (CompareEncHt(ht, *chosen) > 0) {                                           // :1117-1130
      AssignPrevKey(key, same);
      ow_.push_back(OverwriteData{prev_ow, LastExpiration()});
      Slice vs = value;
      DecodeControlFields(&vs, nullptr);
      if (!vs.empty() && (vs[0] == vt::kPackedRowV1 || vs[0] == vt::kPackedRowV2))
        throw NotSupported("packed rows need SchemaPackingProvider (SURVEY 8f-3)");
      return Forward(internal_key, value);
    }

    Slice vs = value;
    Slice intent_doc_ht;
    ControlFields cf = DecodeControlFields(&vs, &intent_doc_ht);                   // :1141
    // :1150-1210 deleted columns / packed-row start need a schema provider: none here, so
    // ColumnDeleted() is false and can_start_packing() is false => falls through.

    const EncodedDocHt& overwrite_ht = (is_ttl_row || CompareEncHt(prev_ow, ht) > 0) ? prev_ow : ht;   // :1212
    const uint8_t value_type = vs.empty() ? 0 : vs[0];

    // CalcExpiration :779-801
    Expiration expiration;
    {
      if (within_merge_block_) expiration = popped;
      else {
        const Expiration last = LastExpiration();
        if (cf.ttl_ns == kMaxTtlNs && !is_ttl_row) expiration = last;
        else {
          uint64_t h; uint32_t w; DecodeDocHt(ht.slice(), &h, &w);
          if (h < last.write_ht) expiration = last;
          else { expiration.write_ht = h; expiration.ttl_ns = cf.ttl_ns; }
        }
      }
    }
    ow_.push_back(OverwriteData{overwrite_ht, expiration});                        // :1226
    if (ow_.size() != new_stack) throw Corruption("Overwrite size does not match new_stack_size");
    AssignPrevKey(key, same);                                                      // :1233

    if (value_type == vt::kTombstone && !CanHaveOtherDataBefore(ht)) return false; // :1246
    if (is_ttl_row) { within_merge_block_ = true; return false; }                  // :1252

    int64_t true_ttl = ComputeTTL(expiration.ttl_ns, r_.table_ttl_ns);             // :1259
    uint64_t key_ht;
    if (true_ttl == expiration.ttl_ns) key_ht = expiration.write_ht;
    else { uint32_t w; DecodeDocHt(ht.slice(), &key_ht, &w); }
    const bool has_expired = HasExpiredTTL(key_ht, true_ttl, chosen_ht);
    if (has_expired) {                                                             // :1268-1277
      if (!CanHaveOtherDataBefore(ht)) return false;
      static const std::string kTomb(1, static_cast<char>(vt::kTombstone));
      return Forward(internal_key, Slice(kTomb));
    } else if (within_merge_block_) {                                              // :1278-1293
      if (expiration.ttl_ns != kMaxTtlNs) {
        uint64_t h; uint32_t w; DecodeDocHt(ht.slice(), &h, &w);
        int64_t diff_us = static_cast<int64_t>(HtMicros(ow_.back().exp.write_ht) - HtMicros(h));
        expiration.ttl_ns += diff_us * 1000;
        ow_.back().exp.ttl_ns = expiration.ttl_ns;
      }
      cf.ttl_ns = expiration.ttl_ns;
      new_value_.clear();
      AppendControlFields(cf, &new_value_);
      new_value_.append(reinterpret_cast<const char*>(vs.p), vs.n);
      within_merge_block_ = false;
      return Forward(internal_key, Slice(new_value_));
    } else if (value_type == vt::kPackedRowV1 || value_type == vt::kPackedRowV2) {
      throw NotSupported("packed rows need SchemaPackingProvider (SURVEY 8f-3)");
    } else if (!intent_doc_ht.empty()) {                                           // :1298-1307
      new_value_.clear();
      AppendControlFields(cf, &new_value_);
      new_value_.append(reinterpret_cast<const char*>(vs.p), vs.n);
      return Forward(internal_key, Slice(new_value_));
    }
    return Forward(internal_key, value);
  }

  void AssignPrevKey(Slice key, size_t same) {   // :1313-1317
    size_t size = ends_.back();
    prev_key_.resize(size);
    memcpy(&prev_key_[0] + same, key.p + same, size - same);
  }

  CompactionFeed* next_;
  RetentionParams r_;
  uint64_t* dropped_;
  EncodedDocHt cutoff_primary_, cutoff_cotables_, min_other_, ht_min_;
  bool has_cotables_cutoff_ = false;
  std::string prev_key_;
  std::vector<size_t> ends_;
  std::vector<OverwriteData> ow_;
  bool within_merge_block_ = false;
  size_t doc_key_serial_ = 0;
  size_t last_passed_serial_ = 0;
  std::string new_value_;
};

// ---------------------------------------------------------------------------------------------
// db/compaction_iterator.cc:139-483 for the no-snapshot case DocDB runs in (snapshots_ empty =>
// visible_at_tip_ = last_sequence): first occurrence of a user key is output, later ones are
// hidden (rule A); kTypeDeletion is dropped at the bottommost level; Merge / SingleDelete entries
// never occur in a regular DocDB and are rejected. Live ranges (:170-196) are applied as a key
// filter (equivalent to the seek for sorted input).
static void CompactionLoop(MergingIterator* input, const CompactionParams& p, CompactionFeed* sink,
                           CompactionStats* st) {
  uint64_t feed_dropped = 0;
  std::unique_ptr<DocDBFeed> docdb;
  CompactionFeed* feed = sink;
  std::vector<std::pair<std::string, std::string>> live;   // GetLiveRanges :1393-1409
  if (p.retention.enabled) {
    docdb.reset(new DocDBFeed(sink, p.retention, &feed_dropped));
    feed = docdb.get();
    const auto& lo = p.retention.lower_bound; const auto& up = p.retention.upper_bound;
    if (!lo.empty() || !up.empty()) {
      std::string meta_end(1, static_cast<char>(kt::kTransactionApplyState + 1));
      live.push_back({std::string(), meta_end});
      live.push_back({Slice(lo).compare(Slice(meta_end)) < 0 ? meta_end : lo, up});
    }
  }
  size_t live_idx = 0;
  std::string current_user_key;
  bool has_current = false;
  std::string out_key;
  while (input->Valid()) {
    Slice key = input->key(), value = input->value();
    st->num_input_records++;
    if (key.n < 8) throw Corruption("Corrupted internal key not expected.");
    Slice user_key(key.p, key.n - 8);
    uint64_t packed = DecodeFixed64(key.p + key.n - 8);
    uint8_t type = packed & 0xff;
    uint64_t seq = packed >> 8;
    if (type > kTypeSingleDeletion || (type > kTypeMerge && type < kTypeSingleDeletion))
      throw Corruption("Corrupted internal key not expected.");
    if (!live.empty()) {
      while (live_idx < live.size() && !live[live_idx].second.empty() &&
             Slice(live[live_idx].second).compare(user_key) < 0) live_idx++;
      if (live_idx >= live.size()) break;
      if (user_key.compare(Slice(live[live_idx].first)) < 0) {
        // Seek(next_range_start): the skipped records are never counted as input by the
        // reference except the one that triggered the seek.
        input->Next();
        while (input->Valid()) {
          Slice k2 = input->key();
          if (Slice(k2.p, k2.n - 8).compare(Slice(live[live_idx].first)) >= 0) break;
          input->Next();
        }
        continue;
      }
    }
    st->total_input_raw_key_bytes += key.n;
    st->total_input_raw_value_bytes += value.n;
    bool first_occurrence = !has_current || !(user_key == Slice(current_user_key));
    if (first_occurrence) { current_user_key = user_key.str(); has_current = true; }
    if (type == kTypeSingleDeletion || type == kTypeMerge)
      throw NotSupported("SingleDelete/Merge records are not produced by regular DocDB");
    if (!first_occurrence) { st->num_dropped_hidden++; input->Next(); continue; }            // rule A
    if (type == kTypeDeletion && seq <= p.last_sequence && p.bottommost_level) {             // :401-420
      st->num_dropped_obsolete++; input->Next(); continue;
    }
    // PrepareOutput :467-483
    Slice out = key;
    if (p.bottommost_level && seq < p.last_sequence && type != kTypeMerge &&
        !(user_key == Slice(p.largest_user_key))) {
      out_key.assign(reinterpret_cast<const char*>(key.p), key.n - 8);
      PutFixed64(&out_key, PackSeqAndType(0, type));
      out = Slice(out_key);
    }
    feed->Feed(out, value);
    input->Next();
  }
  feed->Flush();
  st->num_dropped_feed = feed_dropped;
  if (docdb) { st->smallest_user_values = docdb->smallest_; st->largest_user_values = docdb->largest_; }
}

struct CountingSink : CompactionFeed {
  CompactionFeed* next; CompactionStats* st;
  void Feed(Slice k, Slice v) override {
    st->num_output_records++; st->total_output_raw_key_bytes += k.n; st->total_output_raw_value_bytes += v.n;
    next->Feed(k, v);
  }
  void Flush() override { next->Flush(); }
};

void RunCompaction(const std::vector<SstInput>& inputs, const CompactionParams& params,
                   CompactionFeed* sink, CompactionStats* stats, bool verify) {
  CompactionParams p = params;
  if (!p.has_largest_user_key) {
    // Compaction::GetLargestUserKey (db/compaction.cc:318): max over input files' largest keys.
    std::string best; bool any = false;
    for (auto& in : inputs) {
      TableReader r; r.Open(in.meta, in.data, false);
      if (r.data_blocks.empty()) continue;
      std::string scratch;
      BlockIter it(TableReader::ReadBlock(r.data, r.data_blocks.back(), false, &scratch), r.key_encoding);
      std::string last;
      for (it.SeekToFirst(); it.Valid(); it.Next()) last = it.key().str();
      if (last.size() < 8) continue;
      last.resize(last.size() - 8);
      if (!any || Slice(last).compare(Slice(best)) > 0) { best = last; any = true; }
    }
    p.largest_user_key = best; p.has_largest_user_key = true;
  }
  std::vector<std::unique_ptr<InputIter>> children;
  for (auto& in : inputs) children.emplace_back(new SstIter(in, verify));
  MergingIterator merged(&children);
  CountingSink cs; cs.next = sink; cs.st = stats;
  CompactionLoop(&merged, p, &cs, stats);
}

void RunCompactionOnRuns(const std::vector<KvRun>& runs, const CompactionParams& params,
                         CompactionFeed* sink, CompactionStats* stats) {
  CompactionParams p = params;
  if (!p.has_largest_user_key) {
    std::string best; bool any = false;
    for (auto& r : runs) {
      if (r.kv.empty()) continue;
      std::string last = r.kv.back().first; last.resize(last.size() - 8);
      if (!any || Slice(last).compare(Slice(best)) > 0) { best = last; any = true; }
    }
    p.largest_user_key = best; p.has_largest_user_key = true;
  }
  std::vector<std::unique_ptr<InputIter>> children;
  for (auto& r : runs) children.emplace_back(new RunIter(&r));
  MergingIterator merged(&children);
  CountingSink cs; cs.next = sink; cs.st = stats;
  CompactionLoop(&merged, p, &cs, stats);
}

// ---------------------------------------------------------------------------------------------
// docdb/compaction_file_filter.cc, dockv/doc_ttl_util.cc:42-47,81-129, common/hybrid_time.h:150-165,185-194.
static bool HtIsSpecial(uint64_t v) { return v == kHtMin || v == kHtMax || v == kHtInvalid; }
static uint64_t HtAddDelta(uint64_t ht, int64_t delta_ns) {            // AddDelta -> AddMicroseconds(delta.ToMicroseconds())
  if (HtIsSpecial(ht)) return ht;
  const uint64_t micros = static_cast<uint64_t>(delta_ns / 1000);
  return ht + (micros << 12);
}
ExpirationTime ExtractExpirationTime(bool has_frontier, uint64_t frontier_ht, uint64_t max_value_level_ttl_expiration_ht) {
  if (!has_frontier) return ExpirationTime{};
  ExpirationTime e;
  e.ttl_expiration_ht = max_value_level_ttl_expiration_ht != kHtInvalid ? max_value_level_ttl_expiration_ht : kNoExpiration;   // GetValueOr
  e.created_ht = frontier_ht;
  return e;
}
uint64_t ComputeExpiration(uint64_t ht, int64_t ttl_ns) {
  const uint64_t expiry = HtAddDelta(ht, ttl_ns);
  return CompareHybridTimesToDelta(ht, expiry, ttl_ns) == 0 ? expiry : kNoExpiration;     // overflow check
}
uint64_t MaxExpirationFromValueAndTableTTL(uint64_t key_ht, int64_t table_ttl_ns, uint64_t value_expiry) {
  if (value_expiry == kNoExpiration || HtIsSpecial(key_ht)) return kNoExpiration;
  if (table_ttl_ns == kMaxTtlNs) return value_expiry == kUseDefaultTTL ? kNoExpiration : value_expiry;
  const uint64_t table_expiry = ComputeExpiration(key_ht, table_ttl_ns);
  if (table_expiry == kNoExpiration) return kNoExpiration;
  return value_expiry >= table_expiry ? value_expiry : table_expiry;
}
bool HasExpiredTTL(uint64_t expiration_ht, uint64_t read_ht) {
  if (expiration_ht == kNoExpiration || expiration_ht == kUseDefaultTTL) return false;
  return expiration_ht < read_ht;
}
bool TtlIsExpired(ExpirationTime expiry, int64_t table_ttl_ns, uint64_t now, ExpiryMode mode) {
  const uint64_t ttl_expiry_ht = mode == EXP_TABLE_ONLY ? kUseDefaultTTL : expiry.ttl_expiration_ht;
  if (mode == EXP_TRUST_VALUE && ttl_expiry_ht != kHtInvalid && ttl_expiry_ht != kUseDefaultTTL) return HasExpiredTTL(ttl_expiry_ht, now);
  return HasExpiredTTL(MaxExpirationFromValueAndTableTTL(expiry.created_ht, table_ttl_ns, ttl_expiry_ht), now);
}
std::vector<bool> FileFilterDecisions(const std::vector<ExpirationTime>& files, int64_t table_ttl_ns, uint64_t primary_cutoff_ht,
                                      uint64_t cotables_cutoff_ht, uint64_t now, ExpiryMode mode) {
  // the factory (:196-243): the history cutoff is the smaller valid one; min_kept_ht = the smallest creation time among the
  // files that are NOT expired or still inside the history retention window
  uint64_t history_cutoff = kHtMax;
  if (cotables_cutoff_ht != kHtInvalid) history_cutoff = std::min(history_cutoff, cotables_cutoff_ht);
  if (primary_cutoff_ht != kHtInvalid) history_cutoff = std::min(history_cutoff, primary_cutoff_ht);
  uint64_t min_kept_ht = kHtMax;
  for (const ExpirationTime& e : files)
    if (!TtlIsExpired(e, table_ttl_ns, now, mode) || !(e.created_ht < history_cutoff)) min_kept_ht = std::min(min_kept_ht, e.created_ht);
  // the filter (:150-188): files created before min_kept_ht go, with both sanity checks repeated
  std::vector<bool> discard;
  for (const ExpirationTime& e : files) {
    bool d = false;
    if (e.created_ht < min_kept_ht) d = e.created_ht < history_cutoff && TtlIsExpired(e, table_ttl_ns, now, mode);
    discard.push_back(d);
  }
  return discard;
}

}  // namespace orc
