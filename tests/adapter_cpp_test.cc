// Compiled by tests/test_adapter_cpp.py: exercises the C++ adapter above the C ABI.
// argv[1] = "cpu": only checks that creation fails loudly without a GPU and the host TableBuilder works.
// argv[1] = "gpu": runs a small compaction read from files given as argv[2..] pairs (base, data).
// argv[1] = "gpusub": the same with max_subcompactions = 4 (one output file per key range).
// argv[1] = "gpusnappy": the same as "gpu" with kSnappyCompression output.
#include <cstdlib>
#include <cstdio>
#include <fstream>
#include <iterator>
#include "../yugabyte-db_b200/csrc/adapter/gpu_compaction_job.h"

using namespace ybgpu_adapter;

static std::string ReadFile(const char* p) { std::ifstream f(p, std::ios::binary); return std::string(std::istreambuf_iterator<char>(f), {}); }

int main(int argc, char** argv) {
  std::string mode = argc > 1 ? argv[1] : "cpu";
  ybgpu_job_options o; ybgpu_job_options_init(&o);
  GpuSideTableBuilder tb(o);
  std::string k1 = std::string("a") + std::string(8, '\1'), k2 = std::string("b") + std::string(8, '\1');
  tb.Add(Slice(k1), Slice(std::string("v1")));
  tb.Add(Slice(k2), Slice(std::string("v2")));
  if (!tb.Finish().ok() || tb.NumEntries() != 2) { printf("table builder failed\n"); return 1; }
  {   // compaction_job_test.cc:389-436 SeqNoTrackingWithFewDeletes: inputs 3..4 and 1..2, the one survivor zeroed => 0..4
    InputFile f1, f2; f1.smallest_seqno = 3; f1.largest_seqno = 4; f2.smallest_seqno = 1; f2.largest_seqno = 2;
    uint64_t lo = 99, hi = 99;
    GpuCompactionJob::SeqnoBounds({f1, f2}, 0, 0, 1, &lo, &hi);
    if (lo != 0 || hi != 4) { printf("seqno bounds %llu..%llu\n", (unsigned long long)lo, (unsigned long long)hi); return 1; }
    GpuCompactionJob::SeqnoBounds({InputFile(), InputFile()}, 7, 9, 2, &lo, &hi);      // inputs' bounds unknown: survivors only
    if (lo != 7 || hi != 9) { printf("seqno bounds (untracked inputs)\n"); return 1; }
  }
  {   // ConcatFiles on host-built tables with disjoint key ranges: one index over all blocks, offsets rebased
    ybgpu_job_options to; ybgpu_job_options_init(&to); to.block_size = 256;
    std::vector<GpuCompactionJob::OutputFile> files;
    uint64_t blocks = 0, total = 0;
    for (int p = 0; p < 3; p++) {
      GpuSideTableBuilder b(to);
      GpuCompactionJob::OutputFile f;
      for (int i = 0; i < 200; i++) {
        char k[32]; snprintf(k, sizeof(k), "key%d%05d", p, i);
        std::string ik = std::string(k) + std::string("\1\0\0\0\0\0\0\0", 8);
        b.Add(Slice(ik), Slice(std::string(40, 'v')));
        if (i == 0) f.smallest_key = ik;
        f.largest_key = ik;
      }
      if (!b.Finish().ok()) { printf("piece build failed\n"); return 1; }
      Slice d, m; b.Files(&d, &m);
      f.data_file.assign(reinterpret_cast<const char*>(d.data()), d.size());
      f.base_file.assign(reinterpret_cast<const char*>(m.data()), m.size());
      uint64_t n = 0; int32_t enc = 0;
      ybgpu_sst_meta_handles(m.data(), m.size(), nullptr, 0, &n, &enc);
      blocks += n; total += d.size();
      files.push_back(f);
    }
    std::string data, base;
    Status cs = GpuCompactionJob::ConcatFiles(to, files, &data, &base);
    if (!cs.ok()) { printf("concat: %s\n", cs.ToString().c_str()); return 1; }
    uint64_t n = 0; int32_t enc = 0;
    ybgpu_sst_meta_handles(reinterpret_cast<const uint8_t*>(base.data()), base.size(), nullptr, 0, &n, &enc);
    std::vector<ybgpu_block_handle> h(n);
    ybgpu_sst_meta_handles(reinterpret_cast<const uint8_t*>(base.data()), base.size(), h.data(), n, &n, &enc);
    if (n != blocks || data.size() != total || h.back().offset + h.back().size + 5 != total) { printf("concat layout\n"); return 1; }
    for (uint64_t i = 1; i < n; i++) if (h[i].offset != h[i - 1].offset + h[i - 1].size + 5) { printf("concat handles\n"); return 1; }
  }
  if (mode == "filefilter") {
    // argv: filefilter <table_ttl_ns> <primary_cutoff> <cotables_cutoff> <now> <expiry mode> then per file:
    // <has frontier 0/1> <frontier hybrid time> <max value-level TTL expiration time>. Prints one K / D per file
    // (DocDBCompactionFileFilterFactory + Filter as the adapter restates them), then the TtlIsExpired verdict per file.
    if (argc < 7 || (argc - 7) % 3) { printf("usage\n"); return 2; }
    DocDBRetention r;
    r.table_ttl_ns = strtoll(argv[2], nullptr, 10);
    r.primary_cutoff_ht = strtoull(argv[3], nullptr, 10); r.cotables_cutoff_ht = strtoull(argv[4], nullptr, 10);
    const uint64_t now = strtoull(argv[5], nullptr, 10);
    const ExpiryMode em = static_cast<ExpiryMode>(atoi(argv[6]));
    std::vector<InputFile> files;
    for (int i = 7; i + 2 < argc; i += 3) {
      InputFile f;
      f.has_largest_frontier = atoi(argv[i]) != 0;
      f.frontier_hybrid_time = strtoull(argv[i + 1], nullptr, 10);
      f.max_value_level_ttl_expiration_time = strtoull(argv[i + 2], nullptr, 10);
      files.push_back(f);
    }
    std::vector<InputFile> marked = files;
    const size_t n = MarkExpiredFiles(&marked, r, now, em);
    size_t seen = 0;
    for (const InputFile& f : marked) { printf("%c", f.delete_after_compaction ? 'D' : 'K'); seen += f.delete_after_compaction; }
    if (seen != n) { printf(" count\n"); return 1; }
    printf(" ");
    for (const InputFile& f : files) printf("%c", TtlIsExpired(ExtractExpirationTime(&f), r.table_ttl_ns, now, em) ? 'E' : 'L');
    printf("\n");
    return 0;
  }
  if (mode == "checkfile") {
    // argv: checkfile <base file> <data file> <num entries> <paranoid 0/1> <tail zeros>: prints the status code of
    // GpuCompactionJob::CheckOutputFile.
    if (argc != 7) { printf("usage\n"); return 2; }
    const std::string base = ReadFile(argv[2]), data = ReadFile(argv[3]);
    Status s = GpuCompactionJob::CheckOutputFile(Slice(data), Slice(base), strtoull(argv[4], nullptr, 10), atoi(argv[5]) != 0, strtoull(argv[6], nullptr, 10));
    printf("%d\n", static_cast<int>(s.code()));
    return 0;
  }
  if (mode == "cpu") {
    if (ybgpu_device_count() == 0) {
      GpuCompactionJob job(GpuCompactionJob::Params{});
      Status s = job.Prepare({});
      if (s.ok()) { printf("expected failure without a GPU\n"); return 1; }
      printf("no-gpu status: %s\n", s.ToString().c_str());
    }
    {
      // CompactionJobStats mapping (compaction_job.cc:851-861,897-920,1337-1371) on synthetic engine stats
      ybgpu_job_stats st{};
      st.num_input_records = 100; st.num_output_records = 60; st.num_record_drop_hidden = 25; st.num_record_drop_obsolete = 5;
      st.total_input_raw_key_bytes = 4000; st.total_input_raw_value_bytes = 9000;
      std::string b1(300, 'b'), d1(5000, 'd'), b2(200, 'b'), d2(7000, 'd');
      std::vector<InputFile> in(2);
      in[0].base_file = Slice(b1); in[0].data_file = Slice(d1); in[1].base_file = Slice(b2); in[1].data_file = Slice(d2); in[1].delete_after_compaction = true;
      GpuCompactionJob::CompactionJobStats js;
      GpuCompactionJob::FillCompactionJobStats(st, in, 8123, 1, std::string("abcdefghijkl") + std::string(8, '\1'), std::string("zz") + std::string(8, '\0'), &js);
      if (js.num_input_records != 100 || js.num_output_records != 60 || js.num_records_replaced != 25 || js.num_expired_deletion_records != 5 ||
          js.num_input_files != 2 || js.total_input_bytes != 12500 || js.total_output_bytes != 8123 || js.num_output_files != 1 ||
          js.total_input_raw_key_bytes != 4000 || js.total_input_raw_value_bytes != 9000 || js.smallest_output_key_prefix != "abcdefgh" ||
          js.largest_output_key_prefix != "zz") { printf("job stats mapping\n"); return 1; }
      GpuCompactionJob::FillCompactionJobStats(st, in, 0, 0, "", "", &js);
      if (js.num_output_files != 0) { printf("job stats (no output)\n"); return 1; }
    }
    printf("OK\n");
    return 0;
  }
  std::vector<std::string> bufs;
  for (int i = 2; i + 1 < argc; i += 2) { bufs.push_back(ReadFile(argv[i])); bufs.push_back(ReadFile(argv[i + 1])); }
  std::vector<InputFile> in;
  for (size_t i = 0; i + 1 < bufs.size(); i += 2) { InputFile f; f.base_file = Slice(bufs[i]); f.data_file = Slice(bufs[i + 1]); in.push_back(f); }
  GpuCompactionJob::Params p;
  if (mode == "gpusub") p.max_subcompactions = 4;
  if (mode == "gpusnappy") p.output_compression = YBGPU_COMPRESSION_SNAPPY;   // Options::compression as DocDB sets it (docdb_rocksdb_util.cc:184)
  if (mode == "gpufeed") {
    // RunIntoFeed: the surviving stream goes entry by entry through a host CompactionFeed (compaction_context.h:25-35)
    // into the host TableBuilder; a second job's feed fails at entry `abort_at` and the status must come back
    // unchanged (compaction_job.cc:797-800: the first non-OK Feed aborts the loop).
    struct BuilderFeed : CompactionFeed {
      GpuSideTableBuilder* tb; uint64_t n = 0, abort_at = ~0ull; bool flushed = false;
      Status Feed(const Slice& k, const Slice& v) override {
        if (n == abort_at) return Status(Status::kIOError, "injected feed failure");
        n++; tb->Add(k, v); return tb->status();
      }
      Status Flush() override { flushed = true; return Status::OK(); }
    };
    ybgpu_job_options to; ybgpu_job_options_init(&to);
    GpuSideTableBuilder tb(to);
    BuilderFeed feed; feed.tb = &tb;
    GpuCompactionJob job(p);
    Status s = job.Prepare(in);
    if (!s.ok()) { printf("prepare: %s\n", s.ToString().c_str()); return 1; }
    s = job.RunIntoFeed(&feed);
    if (!s.ok() || !feed.flushed) { printf("RunIntoFeed: %s\n", s.ToString().c_str()); return 1; }
    if (!tb.Finish().ok() || tb.NumEntries() != feed.n || feed.n != job.stats().num_output_records) { printf("feed count\n"); return 1; }
    Slice d, m; tb.Files(&d, &m);
    std::ofstream(std::string(argv[2]) + ".feed.data", std::ios::binary).write(reinterpret_cast<const char*>(d.data()), d.size());
    std::ofstream(std::string(argv[2]) + ".feed.base", std::ios::binary).write(reinterpret_cast<const char*>(m.data()), m.size());
    GpuSideTableBuilder tb2(to);
    BuilderFeed bad; bad.tb = &tb2; bad.abort_at = feed.n / 2;
    GpuCompactionJob job2(p);
    s = job2.Prepare(in);
    if (!s.ok()) { printf("prepare2: %s\n", s.ToString().c_str()); return 1; }
    s = job2.RunIntoFeed(&bad);
    if (s.ok() || s.code() != Status::kIOError || bad.n != feed.n / 2 || bad.flushed) { printf("abort semantics: %s n=%llu\n", s.ToString().c_str(), (unsigned long long)bad.n); return 1; }
    GpuCompactionJob::Params ps = p; ps.max_subcompactions = 4;
    GpuCompactionJob job3(ps);
    s = job3.Prepare(in);
    if (!s.ok()) { printf("prepare3: %s\n", s.ToString().c_str()); return 1; }
    s = job3.RunIntoFeed(&bad);
    if (!s.IsNotSupported()) { printf("RunIntoFeed with subcompactions must be NotSupported\n"); return 1; }
    printf("OK in=%llu out=%llu\n", (unsigned long long)job.stats().num_input_records, (unsigned long long)feed.n);
    return 0;
  }
  GpuCompactionJob job(p);
  Status s = job.Prepare(in);
  if (!s.ok()) { printf("prepare: %s\n", s.ToString().c_str()); return 1; }
  s = job.Run();
  if (!s.ok()) { printf("run: %s\n", s.ToString().c_str()); return 1; }
  GpuCompactionJob::OutputMeta m;
  s = job.Install(&m);
  if (!s.ok()) { printf("install: %s\n", s.ToString().c_str()); return 1; }
  if (mode == "gpusub") {
    for (size_t i = 0; i < job.outputs().size(); i++) {
      std::ofstream(std::string(argv[2]) + ".sub" + std::to_string(i) + ".base", std::ios::binary) << job.outputs()[i].base_file;
      std::ofstream(std::string(argv[2]) + ".sub" + std::to_string(i) + ".data", std::ios::binary) << job.outputs()[i].data_file;
    }
    std::string one_data, one_base;
    Status cs = job.ConcatenatedOutput(&one_data, &one_base);
    if (!cs.ok()) { printf("concat: %s\n", cs.ToString().c_str()); return 1; }
    std::ofstream(std::string(argv[2]) + ".one.base", std::ios::binary) << one_base;
    std::ofstream(std::string(argv[2]) + ".one.data", std::ios::binary) << one_data;
    if (!job.outputs().empty() && (m.smallest_key != job.outputs().front().smallest_key || m.largest_key != job.outputs().back().largest_key)) { printf("boundaries\n"); return 1; }
    printf("OK in=%llu out=%llu files=%zu\n", (unsigned long long)job.stats().num_input_records, (unsigned long long)m.num_entries, job.outputs().size());
    return 0;
  }
  std::ofstream(std::string(argv[2]) + ".out.base", std::ios::binary) << job.output_base_file();
  std::ofstream(std::string(argv[2]) + ".out.data", std::ios::binary) << job.output_data_file();
  printf("OK in=%llu out=%llu\n", (unsigned long long)job.stats().num_input_records, (unsigned long long)m.num_entries);
  return 0;
}
