// host_gen.cc — synthetic split-SST generator for the BASELINE.json config shapes (SURVEY.md 8d
// "Synthetic inputs"), written through the product's own SplitSstWriter. Benchmark tooling: this
// is how bench.py makes its inputs without touching oracle/. The byte layout it produces is
// specified in DESIGN.md ("Synthetic workload") and cross-checked against the oracle's
// independent generator in tests/test_abi_cpu.py.
#include <atomic>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ybgpu_compaction.h"
#include "dev_logic.cuh"
#include "host_sst.h"

namespace {

inline uint64_t Mix64(uint64_t x) {
  x += 0x9e3779b97f4a7c15ull; x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull; x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}

// Row -> 32-byte DocKey: 'G' hash16 'S' <8 base-255 digits of the row id> <16 pseudo-random
// non-zero bytes> 00 00 '!' '!'. hash16 and the digits are monotone in the row id.
void MakeDocKey(const ybgpu_gen_config& c, uint64_t row, uint8_t* out) {
  const uint64_t gid = c.row_offset + row;
  const uint64_t total = c.hash_rows_total ? c.hash_rows_total : c.num_rows;
  uint32_t h16 = static_cast<uint32_t>((static_cast<unsigned __int128>(gid) << 16) / total);
  if (h16 > 0xffff) h16 = 0xffff;
  out[0] = 'G'; out[1] = static_cast<uint8_t>(h16 >> 8); out[2] = static_cast<uint8_t>(h16); out[3] = 'S';
  uint64_t t = gid;
  for (int d = 7; d >= 0; d--) { out[4 + d] = static_cast<uint8_t>(1 + t % 255); t /= 255; }
  uint64_t r = Mix64(c.seed ^ (gid * 0x100000001b3ull));
  for (int j = 0; j < 16; j++) {
    if ((j & 7) == 0 && j) r = Mix64(r);
    out[12 + j] = static_cast<uint8_t>(1 + ((r >> (8 * (j & 7))) & 0xff) % 255);
  }
  out[28] = 0; out[29] = 0; out[30] = '!'; out[31] = '!';
}

}  // namespace

struct ybgpu_sst { std::string data, meta; uint64_t num_entries = 0, raw_bytes = 0; };

extern "C" {

ybgpu_status ybgpu_gen_sst(const ybgpu_gen_config* cfg, uint32_t file_index, const ybgpu_job_options* topts, ybgpu_sst** out) {
  if (!cfg || !topts || !out || file_index >= cfg->num_files) return YBGPU_INVALID_ARGUMENT;
  try {
    const ybgpu_gen_config& c = *cfg;
    ybgpu::host::TableOptions t;
    t.block_size = topts->block_size; t.block_restart_interval = topts->block_restart_interval;
    t.block_size_deviation = topts->block_size_deviation; t.index_block_size = topts->index_block_size;
    t.min_keys_per_index_block = topts->min_keys_per_index_block; t.key_encoding = topts->output_key_encoding;
    ybgpu::host::SplitSstWriter w(t);
    std::vector<uint8_t> key(96), val(c.value_len ? c.value_len : 1);
    uint64_t ordinal = 0, raw = 0;
    const uint64_t seq_base = (1ull << 50) + (static_cast<uint64_t>(file_index) << 34);
    for (uint64_t row = 0; row < c.num_rows; row++) {
      uint8_t dk[32];
      bool have = false;
      for (uint32_t col = 0; col < c.cols; col++) {
        for (uint32_t v = c.versions; v-- > 0;) {
          const uint64_t m = Mix64(c.seed * 0x9e3779b1ull + (c.row_offset + row) * 1315423911ull + col * 2654435761ull + v * 40503ull);
          if (m % c.num_files != file_index) continue;
          if (!have) { MakeDocKey(c, row, dk); have = true; }
          size_t n = 0;
          memcpy(key.data(), dk, 32); n = 32;
          key[n++] = 'K';
          n += ybgpu::fast_varint_encode(static_cast<int64_t>(col) + 1, key.data() + n);
          key[n++] = '#';
          n += ybgpu::doc_ht_encode(((c.base_micros + static_cast<uint64_t>(v) * 1000) << 12), 0, key.data() + n);
          const uint64_t suffix = ((seq_base + ordinal) << 8) | 1;
          memcpy(key.data() + n, &suffix, 8); n += 8;
          ordinal++;
          const uint64_t r = Mix64(m ^ 0xabcdef);
          const bool tomb = c.tombstone_per_1024 && (r & 1023) < c.tombstone_per_1024 && (v + 1 != c.versions || c.tombstone_newest);
          size_t vl;
          if (tomb) { val[0] = 'X'; vl = 1; }
          else {
            vl = c.value_len;
            val[0] = 'S';
            uint64_t x = r | 1;
            size_t j = 1;
            while (j + 8 <= vl) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; memcpy(&val[j], &x, 8); j += 8; }
            while (j < vl) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; val[j++] = static_cast<uint8_t>(x); }
          }
          w.Add(key.data(), n, val.data(), vl);
          raw += n + vl;
        }
      }
    }
    w.Finish();
    ybgpu_sst* s = new ybgpu_sst;
    s->data = w.data_file(); s->meta = w.meta_file(); s->num_entries = w.NumEntries(); s->raw_bytes = raw;
    *out = s;
    return YBGPU_OK;
  } catch (const std::exception&) {
    return YBGPU_NOT_SUPPORTED;
  }
}

ybgpu_status ybgpu_gen_ssts(const ybgpu_gen_config* cfg, const ybgpu_job_options* topts, ybgpu_sst** out, int32_t max_threads) {
  if (!cfg || !out) return YBGPU_INVALID_ARGUMENT;
  std::atomic<uint32_t> next{0};
  std::atomic<int> failed{0};
  const int nt = std::max(1, std::min<int>(max_threads, cfg->num_files));
  std::vector<std::thread> th;
  for (int t = 0; t < nt; t++)
    th.emplace_back([&] {
      for (;;) {
        uint32_t f = next.fetch_add(1);
        if (f >= cfg->num_files) break;
        if (ybgpu_gen_sst(cfg, f, topts, &out[f]) != YBGPU_OK) failed = 1;
      }
    });
  for (auto& t : th) t.join();
  return failed ? YBGPU_RUNTIME_ERROR : YBGPU_OK;
}

void ybgpu_sst_free(ybgpu_sst* s) { delete s; }
const uint8_t* ybgpu_sst_data(const ybgpu_sst* s, uint64_t* len) { *len = s->data.size(); return reinterpret_cast<const uint8_t*>(s->data.data()); }
const uint8_t* ybgpu_sst_meta(const ybgpu_sst* s, uint64_t* len) { *len = s->meta.size(); return reinterpret_cast<const uint8_t*>(s->meta.data()); }
uint64_t ybgpu_sst_num_entries(const ybgpu_sst* s) { return s->num_entries; }
uint64_t ybgpu_sst_raw_bytes(const ybgpu_sst* s) { return s->raw_bytes; }

}  // extern "C"
