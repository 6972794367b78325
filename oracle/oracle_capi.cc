// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_codec.h).
//
// C ABI over the oracle for tests/ (ctypes) and bench.py's cpu_baseline / --impl reference legs,
// plus the synthetic SST generator for the BASELINE.json config shapes.
#include "oracle_compaction.h"
#include <chrono>
#include <thread>
#include <atomic>

using namespace orc;

extern "C" {

struct orc_sst { std::string data, meta; uint64_t num_entries = 0, raw_key = 0, raw_val = 0; std::vector<BlockHandle> handles; };

struct orc_table_options {
  uint32_t block_size; int32_t block_restart_interval; int32_t key_encoding; int32_t block_size_deviation;
  uint32_t index_block_size; uint32_t min_keys_per_index_block;
  int32_t filter_policy; uint32_t filter_block_size;
  int32_t compression;
};

struct orc_compaction_params {
  int32_t bottommost_level;
  uint64_t last_sequence;
  const uint8_t* largest_user_key; uint64_t largest_user_key_len; int32_t has_largest_user_key;
  int32_t retention_enabled;
  uint64_t primary_cutoff_ht, cotables_cutoff_ht;
  int64_t table_ttl_ns;
  int32_t retain_delete_markers;
  uint64_t other_min_ht;
  const uint8_t* lower_bound; uint64_t lower_len;
  const uint8_t* upper_bound; uint64_t upper_len;
};

struct orc_stats {
  uint64_t num_input_records, num_output_records, num_dropped_hidden, num_dropped_obsolete, num_dropped_feed;
  uint64_t in_key_bytes, in_val_bytes, out_key_bytes, out_val_bytes;
  uint64_t kv_hash;       // order-sensitive FNV-1a over (klen,key,vlen,value) of the output stream
  double seconds;         // wall time of the compaction loop (decode..encode), no file IO
};

struct orc_result {
  orc_sst* out = nullptr;               // built output SST (if requested)
  std::string keys, vals;               // collected KV stream (if requested)
  std::vector<uint64_t> koff, voff;
  orc_stats stats{};
  std::vector<std::pair<uint32_t, std::string>> user_values[2];   // [0] smallest, [1] largest (tag, encoded component)
  std::string error;
};

static thread_local std::string g_err;
const char* orc_last_error() { return g_err.c_str(); }

static TableOptions ToOpts(const orc_table_options* o) {
  TableOptions t;
  if (o) {
    if (o->block_size) t.block_size = o->block_size;
    if (o->block_restart_interval) t.block_restart_interval = o->block_restart_interval;
    if (o->key_encoding) t.key_encoding = o->key_encoding;
    if (o->block_size_deviation >= 0) t.block_size_deviation = o->block_size_deviation;
    if (o->index_block_size) t.index_block_size = o->index_block_size;
    t.compression = o->compression;
    if (o->min_keys_per_index_block) t.min_keys_per_index_block = o->min_keys_per_index_block;
    t.filter_policy = o->filter_policy;
    if (o->filter_block_size) t.filter_block_size = o->filter_block_size;
  }
  return t;
}

static CompactionParams ToParams(const orc_compaction_params* p) {
  CompactionParams c;
  c.bottommost_level = p->bottommost_level != 0;
  c.last_sequence = p->last_sequence;
  c.has_largest_user_key = p->has_largest_user_key != 0;
  if (c.has_largest_user_key) c.largest_user_key.assign(reinterpret_cast<const char*>(p->largest_user_key), p->largest_user_key_len);
  c.retention.enabled = p->retention_enabled != 0;
  c.retention.primary_cutoff_ht = p->primary_cutoff_ht;
  c.retention.cotables_cutoff_ht = p->cotables_cutoff_ht;
  c.retention.table_ttl_ns = p->table_ttl_ns;
  c.retention.retain_delete_markers_in_major_compaction = p->retain_delete_markers != 0;
  c.retention.other_min_ht = p->other_min_ht;
  if (p->lower_len) c.retention.lower_bound.assign(reinterpret_cast<const char*>(p->lower_bound), p->lower_len);
  if (p->upper_len) c.retention.upper_bound.assign(reinterpret_cast<const char*>(p->upper_bound), p->upper_len);
  return c;
}

// ---- codecs (pinned by golden vectors in tests/test_oracle_codec.py) -------------------------
int orc_encode_doc_ht(uint64_t ht_repr, uint32_t write_id, uint8_t* out) { return EncodeDocHt(ht_repr, write_id, out); }
int orc_decode_doc_ht(const uint8_t* p, uint64_t n, uint64_t* ht, uint32_t* wid) {
  try { DecodeDocHt(Slice(p, n), ht, wid); return 0; } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
int orc_signed_varint(int64_t v, uint8_t* out) { return FastEncodeSignedVarInt(v, out); }
int orc_unsigned_varint(uint64_t v, uint8_t* out) { return FastEncodeUnsignedVarInt(v, out); }
int orc_decode_signed_varint(const uint8_t* p, uint64_t n, int64_t* v) {
  try { return static_cast<int>(FastDecodeSignedVarInt(p, n, v)); } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
uint32_t orc_crc32c(const uint8_t* p, uint64_t n) { return Crc32cValue(p, n); }
uint32_t orc_crc32c_mask(uint32_t c) { return Crc32cMask(c); }
int orc_subdockey_ends(const uint8_t* key, uint64_t n, uint64_t* ends, int cap) {
  try {
    std::vector<size_t> e; DecodeDocKeyAndSubKeyEnds(Slice(key, n), &e);
    int k = 0; for (size_t x : e) { if (k < cap) ends[k] = x; k++; }
    return k;
  } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
int orc_shortest_separator(const uint8_t* start, uint64_t sn, const uint8_t* limit, uint64_t ln, uint8_t* out, uint64_t cap) {
  std::string s(reinterpret_cast<const char*>(start), sn);
  InternalFindShortestSeparator(&s, Slice(limit, ln));
  if (s.size() > cap) return -1;
  memcpy(out, s.data(), s.size());
  return static_cast<int>(s.size());
}

// Snappy raw format (oracle_sst.h): sizes first (out == NULL), then the bytes. -1 = malformed stream.
int64_t orc_snappy_compress(const uint8_t* p, uint64_t n, uint8_t* out, uint64_t cap) {
  std::string c; SnappyCompress(Slice(p, n), &c);
  if (out) { if (c.size() > cap) return -1; memcpy(out, c.data(), c.size()); }
  return static_cast<int64_t>(c.size());
}
int64_t orc_snappy_uncompress(const uint8_t* p, uint64_t n, uint8_t* out, uint64_t cap) {
  try {
    std::string u; SnappyUncompress(Slice(p, n), &u);
    if (out) { if (u.size() > cap) return -1; memcpy(out, u.data(), u.size()); }
    return static_cast<int64_t>(u.size());
  } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

// ---- whole-file TTL expiration (docdb/compaction_file_filter.cc) ------------------------------
int orc_ttl_is_expired(uint64_t ttl_expiration_ht, uint64_t created_ht, int64_t table_ttl_ns, uint64_t now, int mode) {
  return TtlIsExpired(ExpirationTime{ttl_expiration_ht, created_ht}, table_ttl_ns, now, static_cast<ExpiryMode>(mode)) ? 1 : 0;
}
// frontier fields per file: created[i] = ConsensusFrontier::hybrid_time, value_ttl[i] = max_value_level_ttl_expiration_time
// (kHtInvalid = not set); has_frontier[i] = 0: the file has no largest user frontier. discard[i] = 1: kDiscard.
void orc_file_filter(uint64_t n, const uint8_t* has_frontier, const uint64_t* created, const uint64_t* value_ttl, int64_t table_ttl_ns,
                     uint64_t primary_cutoff_ht, uint64_t cotables_cutoff_ht, uint64_t now, int mode, uint8_t* discard) {
  std::vector<ExpirationTime> files;
  for (uint64_t i = 0; i < n; i++) files.push_back(ExtractExpirationTime(has_frontier[i] != 0, created[i], value_ttl[i]));
  std::vector<bool> d = FileFilterDecisions(files, table_ttl_ns, primary_cutoff_ht, cotables_cutoff_ht, now, static_cast<ExpiryMode>(mode));
  for (uint64_t i = 0; i < n; i++) discard[i] = d[i] ? 1 : 0;
}

// ---- SST build / read ------------------------------------------------------------------------
orc_sst* orc_sst_build(uint64_t n, const uint8_t* keys, const uint64_t* koff, const uint8_t* vals,
                       const uint64_t* voff, const orc_table_options* o) {
  try {
    TableBuilder b(ToOpts(o));
    for (uint64_t i = 0; i < n; i++)
      b.Add(Slice(keys + koff[i], koff[i + 1] - koff[i]), Slice(vals + voff[i], voff[i + 1] - voff[i]));
    b.Finish();
    auto* s = new orc_sst;
    s->data = b.data_file(); s->meta = b.meta_file();
    s->num_entries = b.props().num_entries; s->raw_key = b.props().raw_key_size; s->raw_val = b.props().raw_value_size;
    s->handles = b.data_block_handles();
    return s;
  } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
orc_sst* orc_sst_from_bytes(const uint8_t* meta, uint64_t mn, const uint8_t* data, uint64_t dn) {
  auto* s = new orc_sst;
  s->meta.assign(reinterpret_cast<const char*>(meta), mn);
  s->data.assign(reinterpret_cast<const char*>(data), dn);
  try {
    TableReader r; r.Open(Slice(s->meta), Slice(s->data), false);
    s->handles = r.data_blocks;
  } catch (const std::exception& e) { g_err = e.what(); delete s; return nullptr; }
  return s;
}
void orc_sst_free(orc_sst* s) { delete s; }
uint64_t orc_sst_data_size(const orc_sst* s) { return s->data.size(); }
uint64_t orc_sst_meta_size(const orc_sst* s) { return s->meta.size(); }
const uint8_t* orc_sst_data(const orc_sst* s) { return reinterpret_cast<const uint8_t*>(s->data.data()); }
const uint8_t* orc_sst_meta(const orc_sst* s) { return reinterpret_cast<const uint8_t*>(s->meta.data()); }
uint64_t orc_sst_num_entries(const orc_sst* s) { return s->num_entries; }
uint64_t orc_sst_raw_key_bytes(const orc_sst* s) { return s->raw_key; }
uint64_t orc_sst_raw_val_bytes(const orc_sst* s) { return s->raw_val; }
uint64_t orc_sst_num_blocks(const orc_sst* s) { return s->handles.size(); }
void orc_sst_block_handles(const orc_sst* s, uint64_t* offsets, uint64_t* sizes) {
  for (size_t i = 0; i < s->handles.size(); i++) { offsets[i] = s->handles[i].offset; sizes[i] = s->handles[i].size; }
}
int orc_sst_key_encoding(const orc_sst* s) {
  try { TableReader r; r.Open(Slice(s->meta), Slice(s->data), false); return r.key_encoding; }
  catch (const std::exception& e) { g_err = e.what(); return -1; }
}

// Fully decode an SST into a flat KV list (result->keys/koff/vals/voff).
orc_result* orc_sst_read_all(const orc_sst* s, int verify) {
  auto* res = new orc_result;
  try {
    TableReader r; r.Open(Slice(s->meta), Slice(s->data), verify != 0);
    res->koff.push_back(0); res->voff.push_back(0);
    for (auto& h : r.data_blocks) {
      std::string scratch;
      BlockIter it(TableReader::ReadBlock(r.data, h, verify != 0, &scratch), r.key_encoding);
      for (it.SeekToFirst(); it.Valid(); it.Next()) {
        res->keys.append(reinterpret_cast<const char*>(it.key().p), it.key().n);
        res->vals.append(reinterpret_cast<const char*>(it.value().p), it.value().n);
        res->koff.push_back(res->keys.size()); res->voff.push_back(res->vals.size());
      }
    }
    res->stats.num_output_records = res->koff.size() - 1;
  } catch (const std::exception& e) { res->error = e.what(); }
  return res;
}

// ---- compaction ------------------------------------------------------------------------------
struct ResultSink : CompactionFeed {
  // kv_hash: order-sensitive digest matching ybgpu_job_kv_stream_digest: per entry FNV-1a-64 over
  // (klen u32 LE, key, vlen u32 LE, value), finalised together with the entry index, summed.
  orc_result* res; TableBuilder* builder; bool collect; uint64_t h = 0; uint64_t index = 0;
  static uint64_t Fnv(uint64_t x, const uint8_t* p, size_t n) { for (size_t i = 0; i < n; i++) { x ^= p[i]; x *= 1099511628211ull; } return x; }
  static uint64_t Fin(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; }
  void Feed(Slice k, Slice v) override {
    uint32_t kl = static_cast<uint32_t>(k.n), vl = static_cast<uint32_t>(v.n);
    uint64_t e = 1469598103934665603ull;
    e = Fnv(e, reinterpret_cast<uint8_t*>(&kl), 4); e = Fnv(e, k.p, k.n);
    e = Fnv(e, reinterpret_cast<uint8_t*>(&vl), 4); e = Fnv(e, v.p, v.n);
    h += Fin(e + index * 0x9e3779b97f4a7c15ull);
    index++;
    if (collect) {
      res->keys.append(reinterpret_cast<const char*>(k.p), k.n); res->vals.append(reinterpret_cast<const char*>(v.p), v.n);
      res->koff.push_back(res->keys.size()); res->voff.push_back(res->vals.size());
    }
    if (builder) builder->Add(k, v);
  }
  void Flush() override {}
};

static void FillStats(orc_result* res, const CompactionStats& st, uint64_t hash, double secs) {
  auto& o = res->stats;
  o.num_input_records = st.num_input_records; o.num_output_records = st.num_output_records;
  o.num_dropped_hidden = st.num_dropped_hidden; o.num_dropped_obsolete = st.num_dropped_obsolete;
  o.num_dropped_feed = st.num_dropped_feed;
  o.in_key_bytes = st.total_input_raw_key_bytes; o.in_val_bytes = st.total_input_raw_value_bytes;
  o.out_key_bytes = st.total_output_raw_key_bytes; o.out_val_bytes = st.total_output_raw_value_bytes;
  o.kv_hash = hash; o.seconds = secs;
  res->user_values[0] = st.smallest_user_values; res->user_values[1] = st.largest_user_values;
}

// mode bit0: collect KV stream; bit1: build output SST; bit2: hash_kv==0 -> skip hashing (baseline
// timing should not pay for the test hash): when set, the FNV hash is skipped.
struct orc_cotable_filters { const uint32_t* db_oids; const uint64_t* hybrid_times; uint64_t n; };
orc_result* orc_compact2(int n_inputs, const orc_sst* const* inputs, const uint64_t* ht_filters, const orc_cotable_filters* cotable_filters,
                         const orc_compaction_params* params, const orc_table_options* topts, int mode, int verify_checksums);
orc_result* orc_compact(int n_inputs, const orc_sst* const* inputs, const uint64_t* ht_filters,
                        const orc_compaction_params* params, const orc_table_options* topts, int mode,
                        int verify_checksums) {
  return orc_compact2(n_inputs, inputs, ht_filters, nullptr, params, topts, mode, verify_checksums);
}
orc_result* orc_compact2(int n_inputs, const orc_sst* const* inputs, const uint64_t* ht_filters, const orc_cotable_filters* cotable_filters,
                         const orc_compaction_params* params, const orc_table_options* topts, int mode, int verify_checksums) {
  auto* res = new orc_result;
  try {
    std::vector<SstInput> in;
    for (int i = 0; i < n_inputs; i++) {
      SstInput s; s.meta = Slice(inputs[i]->meta); s.data = Slice(inputs[i]->data);
      s.hybrid_time_filter = ht_filters ? ht_filters[i] : kHtInvalid;
      if (cotable_filters)
        for (uint64_t q = 0; q < cotable_filters[i].n; q++) s.cotable_filters.emplace_back(cotable_filters[i].db_oids[q], cotable_filters[i].hybrid_times[q]);
      in.push_back(s);
    }
    CompactionParams p = ToParams(params);
    std::unique_ptr<TableBuilder> tb;
    if (mode & 2) tb.reset(new TableBuilder(ToOpts(topts)));
    res->koff.push_back(0); res->voff.push_back(0);
    CompactionStats st;
    auto t0 = std::chrono::steady_clock::now();
    uint64_t hash = 0;
    if (mode & 4) {
      struct Fast : CompactionFeed { TableBuilder* b; void Feed(Slice k, Slice v) override { if (b) b->Add(k, v); } void Flush() override {} } fs;
      fs.b = tb.get();
      RunCompaction(in, p, &fs, &st, verify_checksums != 0);
    } else {
      ResultSink sink; sink.res = res; sink.builder = tb.get(); sink.collect = mode & 1;
      RunCompaction(in, p, &sink, &st, verify_checksums != 0);
      hash = sink.h;
    }
    if (tb) {
      if (tb->NumEntries() > 0) tb->Finish();
    }
    auto t1 = std::chrono::steady_clock::now();
    if (tb && tb->NumEntries() > 0) {
      res->out = new orc_sst;
      res->out->data = tb->data_file(); res->out->meta = tb->meta_file();
      res->out->num_entries = tb->props().num_entries; res->out->raw_key = tb->props().raw_key_size;
      res->out->raw_val = tb->props().raw_value_size; res->out->handles = tb->data_block_handles();
    }
    FillStats(res, st, hash, std::chrono::duration<double>(t1 - t0).count());
  } catch (const std::exception& e) { res->error = e.what(); }
  return res;
}

// Compaction over flat sorted runs (no SST decode): run r has entries [run_start[r], run_start[r+1]).
orc_result* orc_compact_runs(int n_runs, const uint64_t* run_start, const uint8_t* keys, const uint64_t* koff,
                             const uint8_t* vals, const uint64_t* voff, const orc_compaction_params* params) {
  auto* res = new orc_result;
  try {
    std::vector<KvRun> runs(n_runs);
    for (int r = 0; r < n_runs; r++)
      for (uint64_t i = run_start[r]; i < run_start[r + 1]; i++)
        runs[r].kv.emplace_back(std::string(reinterpret_cast<const char*>(keys + koff[i]), koff[i + 1] - koff[i]),
                                std::string(reinterpret_cast<const char*>(vals + voff[i]), voff[i + 1] - voff[i]));
    res->koff.push_back(0); res->voff.push_back(0);
    ResultSink sink; sink.res = res; sink.builder = nullptr; sink.collect = true;
    CompactionStats st;
    RunCompactionOnRuns(runs, ToParams(params), &sink, &st);
    FillStats(res, st, sink.h, 0);
  } catch (const std::exception& e) { res->error = e.what(); }
  return res;
}

void orc_result_free(orc_result* r) { if (r) { delete r->out; delete r; } }
const char* orc_result_error(const orc_result* r) { return r->error.empty() ? nullptr : r->error.c_str(); }
const orc_stats* orc_result_stats(const orc_result* r) { return &r->stats; }
const orc_sst* orc_result_sst(const orc_result* r) { return r->out; }
uint64_t orc_result_num_kv(const orc_result* r) { return r->koff.empty() ? 0 : r->koff.size() - 1; }
const uint8_t* orc_result_keys(const orc_result* r) { return reinterpret_cast<const uint8_t*>(r->keys.data()); }
const uint8_t* orc_result_vals(const orc_result* r) { return reinterpret_cast<const uint8_t*>(r->vals.data()); }
const uint64_t* orc_result_koff(const orc_result* r) { return r->koff.data(); }
const uint64_t* orc_result_voff(const orc_result* r) { return r->voff.data(); }
uint64_t orc_result_keys_size(const orc_result* r) { return r->keys.size(); }
uint32_t orc_result_num_user_values(const orc_result* r, int which) { return static_cast<uint32_t>(r->user_values[which & 1].size()); }
uint32_t orc_result_user_value(const orc_result* r, int which, uint32_t i, const uint8_t** p, uint64_t* n) {
  const auto& v = r->user_values[which & 1][i];
  *p = reinterpret_cast<const uint8_t*>(v.second.data()); *n = v.second.size();
  return v.first;
}
uint64_t orc_result_vals_size(const orc_result* r) { return r->vals.size(); }

// ---- synthetic SST generator (SURVEY.md 8d "Synthetic inputs") -------------------------------
// Row i (0 <= i < num_rows) has DocKey 'G' hash16 'S' <24 non-zero bytes> 00 00 '!' '!' (32 B);
// hash16 and the first 8 string bytes are monotone in i so files come out sorted without a sort;
// the remaining 16 string bytes are pseudo-random. Each row has `cols` columns ('K' + column id)
// and each (row, column) has `versions` versions at HT base_micros + v*1000 (v = 0 oldest).
// Version (i, c, v) lives in file mix(seed, i, c, v) % num_files, so runs interleave randomly.
// Values: 'S' + (value_len-1) pseudo-random bytes, or the 1-byte tombstone 'X' with probability
// tombstone_per_1024/1024 (never for v == versions-1 unless tombstone_newest).
struct orc_gen_config {
  uint64_t seed, num_rows; uint32_t cols, versions, num_files, value_len;
  uint64_t base_micros; uint32_t tombstone_per_1024; uint32_t tombstone_newest;
  uint64_t row_offset;      // first row id (tablet sharding: distinct key ranges per tablet)
  uint64_t hash_rows_total; // rows used to scale hash16 (0 => num_rows)
};

static inline uint64_t Mix64(uint64_t x) { x += 0x9e3779b97f4a7c15ull; x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull; x = (x ^ (x >> 27)) * 0x94d049bb133111ebull; return x ^ (x >> 31); }

static inline void GenDocKey(const orc_gen_config& c, uint64_t row, uint8_t* out /*32*/) {
  uint64_t gid = c.row_offset + row;
  uint64_t total = c.hash_rows_total ? c.hash_rows_total : c.num_rows;
  uint32_t h16 = static_cast<uint32_t>((static_cast<unsigned __int128>(gid) << 16) / total);
  if (h16 > 0xffff) h16 = 0xffff;
  out[0] = 'G'; out[1] = static_cast<uint8_t>(h16 >> 8); out[2] = static_cast<uint8_t>(h16); out[3] = 'S';
  // 8 base-255 digits (+1 => non-zero), big-endian: strictly increasing with gid.
  uint64_t t = gid;
  for (int d = 7; d >= 0; d--) { out[4 + d] = static_cast<uint8_t>(1 + t % 255); t /= 255; }
  uint64_t r = Mix64(c.seed ^ (gid * 0x100000001b3ull));
  for (int j = 0; j < 16; j++) {
    if ((j & 7) == 0 && j) r = Mix64(r);
    out[12 + j] = static_cast<uint8_t>(1 + ((r >> (8 * (j & 7))) & 0xff) % 255);
  }
  out[28] = 0; out[29] = 0; out[30] = '!'; out[31] = '!';
}

orc_sst* orc_gen_sst(const orc_gen_config* cfg, uint32_t file_index, const orc_table_options* o) {
  try {
    const orc_gen_config& c = *cfg;
    TableBuilder b(ToOpts(o));
    std::string key, val;
    uint64_t ordinal = 0;
    const uint64_t seq_base = (1ull << 50) + (static_cast<uint64_t>(file_index) << 34);
    uint8_t tmp[40];
    for (uint64_t row = 0; row < c.num_rows; row++) {
      uint8_t dk[32];
      bool have_dk = false;
      for (uint32_t col = 0; col < c.cols; col++) {
        for (uint32_t vv = c.versions; vv-- > 0;) {     // newest first = ascending key order
          uint64_t m = Mix64(c.seed * 0x9e3779b1ull + (c.row_offset + row) * 1315423911ull + col * 2654435761ull + vv * 40503ull);
          if (m % c.num_files != file_index) continue;
          if (!have_dk) { GenDocKey(c, row, dk); have_dk = true; }
          key.assign(reinterpret_cast<char*>(dk), 32);
          key.push_back('K');
          int n = FastEncodeSignedVarInt(static_cast<int64_t>(col) + 1, tmp);
          key.append(reinterpret_cast<char*>(tmp), n);
          key.push_back('#');
          n = EncodeDocHt(HtFromMicros(c.base_micros + static_cast<uint64_t>(vv) * 1000), 0, tmp);
          key.append(reinterpret_cast<char*>(tmp), n);
          PutFixed64(&key, PackSeqAndType(seq_base + ordinal, kTypeValue));
          ordinal++;
          uint64_t r = Mix64(m ^ 0xabcdef);
          bool tomb = c.tombstone_per_1024 && (r & 1023) < c.tombstone_per_1024 &&
                      (vv + 1 != c.versions || c.tombstone_newest);
          if (tomb) {
            val.assign(1, 'X');
          } else {
            val.resize(c.value_len);
            val[0] = 'S';
            uint64_t x = r | 1;
            size_t j = 1;
            while (j + 8 <= val.size()) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; memcpy(&val[j], &x, 8); j += 8; }
            while (j < val.size()) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; val[j++] = static_cast<char>(x); }
          }
          b.Add(Slice(key), Slice(val));
        }
      }
    }
    b.Finish();
    auto* s = new orc_sst;
    s->data = b.data_file(); s->meta = b.meta_file();
    s->num_entries = b.props().num_entries; s->raw_key = b.props().raw_key_size; s->raw_val = b.props().raw_value_size;
    s->handles = b.data_block_handles();
    return s;
  } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}

// Generate all files of a config concurrently (one thread per file, up to max_threads at once).
int orc_gen_ssts(const orc_gen_config* cfg, const orc_table_options* o, orc_sst** out, int max_threads) {
  std::atomic<uint32_t> next{0};
  std::atomic<int> failed{0};
  int nt = std::max(1, std::min<int>(max_threads, cfg->num_files));
  std::vector<std::thread> th;
  for (int t = 0; t < nt; t++)
    th.emplace_back([&] {
      for (;;) {
        uint32_t f = next.fetch_add(1);
        if (f >= cfg->num_files) break;
        out[f] = orc_gen_sst(cfg, f, o);
        if (!out[f]) failed = 1;
      }
    });
  for (auto& t : th) t.join();
  return failed ? -1 : 0;
}

// Metadata dump for tests: [u32 n_props]{[u32 klen][key][u32 vlen][val]} [u32 n_filters]{[u32 klen][index key][u32 flen][filter block]}.
// Returns the size needed; fills out when cap suffices.
uint64_t orc_sst_meta_dump(const orc_sst* s, uint8_t* out, uint64_t cap) {
  try {
    TableReader r; r.Open(Slice(s->meta), Slice(s->data), true);
    std::string b;
    auto put = [&](const std::string& x) { PutFixed32(&b, static_cast<uint32_t>(x.size())); b.append(x); };
    PutFixed32(&b, static_cast<uint32_t>(r.properties.size()));
    for (auto& kv : r.properties) { put(kv.first); put(kv.second); }
    PutFixed32(&b, static_cast<uint32_t>(r.filter_blocks.size()));
    for (auto& f : r.filter_blocks) { put(f.first); put(TableReader::ReadBlock(r.meta, f.second, true).str()); }
    if (cap >= b.size()) memcpy(out, b.data(), b.size());
    return b.size();
  } catch (const std::exception& e) { g_err = e.what(); return 0; }
}

// ---- bloom filter helpers (tests) -------------------------------------------------------------
uint32_t orc_bloom_hash(const uint8_t* key, uint64_t n) { return BloomHash(Slice(key, n)); }
uint64_t orc_docdb_filter_prefix(const uint8_t* key, uint64_t n) { return DocKeyV3FilterPrefix(Slice(key, n)); }
// Builds one fixed-size filter block from `count` keys ([len u32][bytes] records); returns its size
// (contents without trailer) and copies at most cap bytes to out. params[0..2] = max_keys, num_lines, num_probes.
uint64_t orc_fixed_size_filter(uint64_t total_bits, const uint8_t* keys, uint64_t count, uint8_t* out, uint64_t cap, uint64_t* params) {
  FixedSizeFilterBits f(total_bits, 0.01);
  const uint8_t* p = keys;
  for (uint64_t i = 0; i < count; i++) { uint32_t n; memcpy(&n, p, 4); f.AddKey(Slice(p + 4, n)); p += 4 + n; }
  if (params) { params[0] = f.max_keys(); params[1] = f.num_lines(); params[2] = f.num_probes(); }
  std::string s = f.Finish();
  memcpy(out, s.data(), std::min<uint64_t>(cap, s.size()));
  return s.size();
}
int orc_filter_may_match(const uint8_t* filter, uint64_t flen, const uint8_t* key, uint64_t klen) {
  return FixedSizeFilterBits::MayMatch(Slice(filter, flen), Slice(key, klen)) ? 1 : 0;
}

}  // extern "C"
