// TEST-ONLY host harness: runs the __host__ __device__ per-entry logic of
// yugabyte-db_b200/csrc/dev_logic.cuh on the CPU, in the same order the GPU tile kernel applies
// it (records -> merged order -> rule A -> row groups -> feed_step), so the logic can be checked
// against the oracle in a container without a GPU. It is not part of the product and is never
// loaded by it.
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../../yugabyte-db_b200/csrc/dev_logic.cuh"

using namespace ybgpu;

namespace {
struct Out { std::string keys, vals; std::vector<uint64_t> koff{0}, voff{0}; int error = 0; };
Out* g_out = nullptr;
int g_cut_every = 0;     // > 0: emulate a merge-tile boundary before every g_cut_every-th sorted record
}

extern "C" {

// Emulates tiles that start in the middle of a row group: before every n-th record (in merged order) that is not
// the first of its group, the feed state is thrown away and rebuilt the way the tile kernel does it — reset,
// cotable seed, replay_ancestors over the records that precede it.
void hh_set_cut_every(int n) { g_cut_every = n; }

// returns 0 or a positive DevError
int hh_compact(int n_runs, const uint64_t* run_start, const uint8_t* keys, const uint64_t* koff, const uint8_t* vals,
               const uint64_t* voff, int retention, uint64_t cutoff_ht, int64_t table_ttl_ns, int retain_markers,
               uint64_t other_min_ht, int bottommost, uint64_t last_sequence, const uint8_t* largest, uint64_t largest_len,
               const uint8_t* lower, uint64_t lower_len, const uint8_t* upper, uint64_t upper_len, uint64_t cotables_cutoff_ht) {
  delete g_out; g_out = new Out;
  const uint64_t n = run_start[n_runs];
  size_t max_ulen = 0;
  for (uint64_t i = 0; i < n; i++) max_ulen = std::max<size_t>(max_ulen, koff[i + 1] - koff[i] - 8);
  const int S = std::max<int>(32, static_cast<int>(((max_ulen + 16) + 15) & ~15ull));
  std::vector<uint8_t> recs(static_cast<size_t>(n) * S + 16, 0);
  // align to 16
  uint8_t* base = recs.data();
  std::vector<uint8_t> aligned_store(static_cast<size_t>(n) * S + 64, 0);
  base = aligned_store.data();
  base += (16 - (reinterpret_cast<uintptr_t>(base) & 15)) & 15;
  for (uint64_t i = 0; i < n; i++) {
    uint8_t* r = base + i * S;
    const uint32_t klen = static_cast<uint32_t>(koff[i + 1] - koff[i]);
    const uint32_t ulen = klen - 8, vlen = static_cast<uint32_t>(voff[i + 1] - voff[i]);
    memcpy(r, keys + koff[i], ulen);
    memcpy(r + S - 16, keys + koff[i] + ulen, 8);
    uint16_t ul16 = static_cast<uint16_t>(ulen); memcpy(r + S - 8, &ul16, 2);
    r[S - 6] = vlen ? vals[voff[i]] : 0; r[S - 5] = 0;
    memcpy(r + S - 4, &vlen, 4);
  }
  std::vector<uint32_t> order(n);
  for (uint64_t i = 0; i < n; i++) order[i] = static_cast<uint32_t>(i);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return cmp_records(base + size_t(a) * S, base + size_t(b) * S, S) < 0; });

  RetentionDev R{};
  R.enabled = retention; R.cutoff_ht = cutoff_ht; R.table_ttl_ns = table_ttl_ns;
  R.cutoff_enc.n = static_cast<uint8_t>(doc_ht_encode(cutoff_ht, 0xffffffffu, R.cutoff_enc.b));
  R.min_other_enc.n = static_cast<uint8_t>(doc_ht_encode(retain_markers ? 0 : other_min_ht, 0, R.min_other_enc.b));
  R.ht_min_enc.n = static_cast<uint8_t>(doc_ht_encode(0, 0, R.ht_min_enc.b));
  R.has_cotables_cutoff = cotables_cutoff_ht != 0xfffffffffffffffeull; R.cotables_cutoff_ht = cotables_cutoff_ht;
  if (R.has_cotables_cutoff) R.cotables_cutoff_enc.n = static_cast<uint8_t>(doc_ht_encode(cotables_cutoff_ht, 0xffffffffu, R.cotables_cutoff_enc.b));
  R.lower_len = static_cast<uint32_t>(lower_len); memcpy(R.lower, lower, lower_len);
  R.upper_len = static_cast<uint32_t>(upper_len); memcpy(R.upper, upper, upper_len);

  // merged-order copy of the records (one "run") for replay_ancestors
  std::vector<uint8_t> sorted_store(static_cast<size_t>(n) * S + 64, 0);
  uint8_t* sbase = sorted_store.data();
  sbase += (16 - (reinterpret_cast<uintptr_t>(sbase) & 15)) & 15;
  std::vector<uint64_t> sorted_voff(n + 1, 0);
  if (g_cut_every > 0)
    for (uint64_t i = 0; i < n; i++) { memcpy(sbase + i * S, base + size_t(order[i]) * S, S); sorted_voff[i] = voff[order[i]]; }

  FeedState st; feed_state_reset(&st);
  const uint8_t* prev_group = nullptr; int prev_g = -1;
  auto seed_cotable = [&](const uint8_t* e, uint32_t ulen) -> int {
      // cotable / colocated rows: seed slot 0 from the table's tombstone entries, as the tile kernel does
      if (retention && ulen && (e[0] == 'y' || e[0] == '0')) {
        const int id = dockey_id_size(e, ulen);
        if (id > 0 && static_cast<uint32_t>(id) < ulen && e[id] != '!') {
          FeedState ts; feed_state_reset(&ts);
          const uint8_t* prev_t = nullptr; bool any = false;
          for (uint64_t x = 0; x < n; x++) {
            const uint8_t* c = base + size_t(order[x]) * S;
            const uint32_t cl = rec_ulen(c, S);
            if (cl < static_cast<uint32_t>(id) + 1 || memcmp(c, e, id) != 0 || c[id] != '!') continue;
            if (prev_t && cmp_user_keys(prev_t, rec_ulen(prev_t, S), c, cl) == 0) { continue; }
            prev_t = c;
            const uint64_t sfx = rec_suffix(c, S);
            if ((sfx & 0xff) == 0 && bottommost && (sfx >> 8) <= last_sequence) continue;
            ValueRewrite rw2{};
            const uint32_t vl = rec_vlen(c, S);
            int d2 = feed_step(&ts, R, c, cl, rec_vfirst(c, S), has_control_fields(rec_vfirst(c, S)) ? vals + voff[order[x]] : nullptr, vl, &rw2);
            if (d2 < 0) return -d2;
            any = true;
          }
          if (any && ts.n_ow >= 1 && ts.n_ends == 1) feed_state_seed(&st, e, id, ts.ow[0]);
        }
      }
    return 0;
  };
  const uint8_t* prev_rec = nullptr;
  for (uint64_t i = 0; i < n; i++) {
    const uint32_t id = order[i];
    const uint8_t* e = base + size_t(id) * S;
    const uint32_t ulen = rec_ulen(e, S);
    const int g = group_prefix_len(e, ulen, retention != 0);
    if (g < 0) return -g;
    const bool new_group = !prev_group || g != prev_g || common_prefix_len(e, g, prev_group, g) < static_cast<uint32_t>(g);
    if (new_group) {
      feed_state_reset(&st); prev_group = e; prev_g = g;
      int rc = seed_cotable(e, ulen);
      if (rc) return rc;
    } else if (retention && g_cut_every > 0 && i % g_cut_every == 0) {
      // a tile boundary inside the group: rebuild the state from nothing
      feed_state_reset(&st);
      int rc = seed_cotable(e, ulen);
      if (rc) return rc;
      ReplayRun rr{sbase, static_cast<uint32_t>(i), vals, sorted_voff.data()};
      const int d3 = replay_ancestors(&st, R, &rr, 1, S, sbase + i * S, ulen, bottommost, last_sequence);
      if (d3 < 0) return -d3;
    }
    const bool first_occ = !prev_rec || cmp_user_keys(prev_rec, rec_ulen(prev_rec, S), e, ulen) != 0;
    prev_rec = e;
    if (!first_occ) continue;
    uint64_t suffix = rec_suffix(e, S);
    const uint32_t type = suffix & 0xff; const uint64_t seq = suffix >> 8;
    if (type == 0 && bottommost && seq <= last_sequence) continue;
    if (bottommost && seq < last_sequence) {
      bool is_largest = ulen == largest_len && memcmp(e, largest, ulen) == 0;
      if (!is_largest) suffix &= 0xff;
    }
    const uint8_t* val = vals + voff[id];
    const uint32_t vlen = rec_vlen(e, S);
    int d = ENT_KEEP;
    ValueRewrite rw{};
    if (retention) {
      // the device only dereferences the value when control fields are announced
      d = feed_step(&st, R, e, ulen, rec_vfirst(e, S), has_control_fields(rec_vfirst(e, S)) ? val : nullptr, vlen, &rw);
      if (d < 0) return -d;
      if (d == 0) continue;
    }
    g_out->keys.append(reinterpret_cast<const char*>(e), ulen);
    g_out->keys.append(reinterpret_cast<const char*>(&suffix), 8);
    if (d & ENT_VAL_TOMBSTONE) g_out->vals.push_back('X');
    else if (d & ENT_VAL_REENCODE) {
      g_out->vals.append(reinterpret_cast<const char*>(rw.prefix), rw.prefix_len);
      g_out->vals.append(reinterpret_cast<const char*>(val + rw.skip), vlen - rw.skip);
    } else g_out->vals.append(reinterpret_cast<const char*>(val), vlen);
    g_out->koff.push_back(g_out->keys.size()); g_out->voff.push_back(g_out->vals.size());
  }
  return 0;
}

uint64_t hh_num() { return g_out->koff.size() - 1; }
const uint8_t* hh_keys() { return reinterpret_cast<const uint8_t*>(g_out->keys.data()); }
const uint8_t* hh_vals() { return reinterpret_cast<const uint8_t*>(g_out->vals.data()); }
const uint64_t* hh_koff() { return g_out->koff.data(); }
const uint64_t* hh_voff() { return g_out->voff.data(); }

// hidden_by_ht_filters (HybridTimeFilteringIterator::Satisfied, negated)
int hh_hidden_by_ht_filters(const uint8_t* key, uint32_t ulen, uint64_t global, const uint32_t* oids, const uint64_t* hts, uint32_t n) {
  return hidden_by_ht_filters(key, ulen, global, oids, hts, n) ? 1 : 0;
}

int hh_group_prefix_len(const uint8_t* key, int ulen, int retention) {
  std::vector<uint8_t> buf(ulen + 32, 0); uint8_t* p = buf.data(); p += (16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15;
  memcpy(p, key, ulen);
  return group_prefix_len(p, ulen, retention != 0);
}
// Bloom filter key length two ways: out[0] = by-product of the row-group walk (what the merge kernel
// stores), out[1] = docdb_filter_prefix_len (what the host writer uses). Returns the group prefix length.
int hh_filter_len_pair(const uint8_t* key, int ulen, int retention, int* out) {
  std::vector<uint8_t> buf(ulen + 32, 0); uint8_t* p = buf.data(); p += (16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15;
  memcpy(p, key, ulen);
  int f = -1;
  const int g = group_prefix_len(p, ulen, retention != 0, &f);
  out[0] = f; out[1] = docdb_filter_prefix_len(p, ulen);
  return g;
}
uint32_t hh_bloom_hash(const uint8_t* key, uint32_t n) { return leveldb_hash(key, n, kBloomSeed); }
int hh_doc_ht_encode(uint64_t ht, uint32_t wid, uint8_t* out) { return doc_ht_encode(ht, wid, out); }
int hh_parse_entry_header(const uint8_t* p, uint32_t avail, uint32_t* a, uint32_t* b, uint32_t* c) { return parse_entry_header(p, avail, a, b, c); }

}  // extern "C"
