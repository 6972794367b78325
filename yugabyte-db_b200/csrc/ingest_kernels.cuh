// ingest_kernels.cuh — K0/K1 fused: one pass over the input files.
//
// Replaces, for shared-prefix inputs with internal keys of at most 64 bytes, the three passes
//   k_crc_blocks (ReadBlock checksum verification, table/format.cc:352-395),
//   k_prepass    (entry counts, validation; BlockIter walk, table/block.cc:348-447),
//   k_decode_*   (key reconstruction into fixed-stride records)
// by ONE kernel that reads every input byte from HBM exactly once:
//
//   * a data block (contents + 5-byte trailer) is staged into shared memory by one 1-D TMA bulk copy
//     (cp.async.bulk global->shared, completion on an mbarrier) — no register staging, no per-thread loads;
//   * lanes own restart intervals and walk the entry headers in shared memory (count + validation);
//   * one thread per entry computes the RAW CRC32C (zero initial register, no final complement) of the
//     entry's value and of the bytes in front of it (header + key delta). The block checksum is the GF(2)-
//     linear combination of the segment CRCs (crc(A || B) = crc(A) * x^(8|B|) + crc(B)), compared with the
//     stored trailer; the per-entry VALUE CRCs are kept (4 B per entry): the block encoder derives the
//     output blocks' checksums from them without ever looking at value bytes again (values are copied
//     verbatim from input to output, so their CRC contribution only needs shifting);
//   * the entry base of the block inside its file comes from a decoupled look-back over per-block counts
//     (blocks are claimed through an atomic ticket, so every predecessor of a claimed block is running);
//   * lanes walk their intervals a second time, rebuild the internal keys in registers and write records.
//
// CRC table look-ups dominate the shared-memory pipe; the four 256-entry slicing tables are replicated
// ING_REP times, lane l using copy l % ING_REP, which cuts the expected bank conflict degree of a look-up
// from ~3.5 to ~2.1.
//
// Anything this kernel does not take — other key encodings, keys longer than 64 bytes, blocks larger than the
// staging buffer, more than ING_MAXE entries in a block — raises J->ingest_fallback (not an error) and the host
// runs the general kernels (k_prepass, k_decode_all, k_crc_blocks, k_value_crc) instead.
//
// Included by engine.cu only (after encode_kernels.cuh: CRC tables and helpers).
#pragma once

namespace ybgpu {

constexpr int ING_THREADS = 128;
constexpr uint32_t ING_BUF = 36 * 1024;       // staged bytes per block: contents + trailer + alignment slack
constexpr int ING_MAXE = 512;                 // entries per block
constexpr int ING_REP = 8;                    // replication of the CRC tables
constexpr int ING_NVI = 4;                    // internal keys up to 64 bytes
constexpr int ING_FALLBACK_WIDER = 1;         // a key does not fit the guessed record stride: once more with the widest
constexpr int ING_FALLBACK_GENERAL = 2;       // not for this kernel: general path
constexpr size_t ING_SMEM = ING_BUF + 32 + 4 * 256 * ING_REP * 4 + ING_MAXE * 8 + ING_MAXE * 4;

struct IngestView {
  const RunView* runs;
  const uint32_t* blk_base;        // [k+1] first global block index of every run
  unsigned long long* status;      // [total blocks] look-back word: flag << 32 | entries
  uint32_t* ticket;
  uint32_t* totals;                // [k] exact entries per run
  const uint32_t* cap;             // [k] record capacity per run (upper bound the arrays were sized by)
  const RangeDev* range;           // nullptr = no key range
  int k, S, verify;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D TMA bulk copy global -> shared; dst, src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
// Per run: sum of the blocks' restart counts (-> an upper bound of the entries: every interval holds at most
// `restart interval` entries) and the restart interval itself, read off the first block that has two intervals.
__global__ void __launch_bounds__(256) k_restart_probe(const RunView* runs, const uint32_t* blk_base, int k, unsigned long long* restart_sum, JobDev* J) {
  const uint32_t total = blk_base[k];
  for (uint32_t gb = blockIdx.x * blockDim.x + threadIdx.x; gb < total; gb += gridDim.x * blockDim.x) {
    int ri_ = 0;
    while (blk_base[ri_ + 1] <= gb) ri_++;
    const RunView& run = runs[ri_];
    const uint32_t b = gb - blk_base[ri_];
    const uint8_t* blk = run.data + run.blk_off[b];
    const uint32_t size = run.blk_size[b];
    uint32_t nres = size >= 4 ? ldg_u32_unaligned(blk + size - 4) : 0;
    const uint8_t type = blk[size];                        // trailer: compression type of the stored block
    if (type != 0) {
      // Snappy blocks are uncompressed by the host's next stage (snappy_kernels.cuh), then the probe runs again
      if (type == 1) atomicAdd(&J->n_compressed, 1u); else dev_fail(J, DEV_ERR_COMPRESSED, b);
      nres = 0;
    } else if (nres == 0 || static_cast<uint64_t>(nres) * 4 + 4 > size) { dev_fail(J, DEV_ERR_BAD_BLOCK, b); nres = 0; }
    {
      // one atomic per (warp, run): consecutive blocks almost always belong to the same file
      const uint32_t active = __activemask();
      const uint32_t peers = __match_any_sync(active, ri_);
      const uint32_t sum = __reduce_add_sync(peers, nres);
      if (lane_id() == __ffs(peers) - 1 && sum) atomicAdd(&restart_sum[ri_], static_cast<unsigned long long>(sum));
    }
    if (nres == 0) continue;
    // The first interval of every 16th block is walked as well: the longest key met is the host's guess for the
    // record stride (a longer key inside k_ingest only costs a second attempt with the widest stride).
    if ((nres >= 2 && __ldcg(&J->restart_interval[ri_]) == 0) || (nres >= 2 && (b & 15) == 0)) {
      // walk the first interval's headers (either encoding is delimited by the second restart offset)
      const uint32_t restarts_off = size - 4 - 4 * nres;
      uint32_t p = ldg_u32_unaligned(blk + restarts_off);
      const uint32_t end = ldg_u32_unaligned(blk + restarts_off + 4);
      if (p > end || end > restarts_off) { dev_fail(J, DEV_ERR_BAD_BLOCK, b); continue; }
      uint32_t n = 0, klen = 0, maxk = 0;
      bool ok = true;
      while (p < end && ok) {
        if (run.key_encoding == 2) {
          TspHeader th; uint32_t nk, ms, ml;
          const int h = parse_entry_header_tsp(blk + p, end - p, &th);
          ok = h && tsp_key_layout(th, klen, &nk, &ms, &ml);
          if (ok) { klen = nk; p += h + th.ns1 + th.ns2 + th.vlen; n++; }
        } else {
          uint32_t shared, non_shared, vlen;
          const int h = parse_entry_header(blk + p, end - p, &shared, &non_shared, &vlen);
          ok = h != 0;
          if (ok) { klen = shared + non_shared; p += h + non_shared + vlen; n++; }
        }
        maxk = max(maxk, klen);
      }
      if (!ok || p != end || n == 0) { dev_fail(J, DEV_ERR_BAD_ENTRY, b); continue; }
      atomicCAS(&J->restart_interval[ri_], 0u, n);
      atomicMax(&J->max_ikey_len, maxk);
    }
  }
}

struct IngEntry { uint16_t estart, vstart; uint32_t vlen; };    // offsets inside the block

// raw CRC32C step functions over the lane's private copy of the tables: tabs[(t * 256 + e) * ING_REP + copy]
#define ING_TAB(t, e) tabs[(((t) << 8) + (e)) * ING_REP + copy]
__device__ __forceinline__ uint32_t ing_crc_byte(const uint32_t* tabs, uint32_t copy, uint32_t c, uint32_t byte) {
  return ING_TAB(0, (c ^ byte) & 0xff) ^ (c >> 8);
}
__device__ __forceinline__ uint32_t ing_crc_word(const uint32_t* tabs, uint32_t copy, uint32_t c, uint32_t w) {
  c ^= w;
  return ING_TAB(3, c & 0xff) ^ ING_TAB(2, (c >> 8) & 0xff) ^ ING_TAB(1, (c >> 16) & 0xff) ^ ING_TAB(0, c >> 24);
}
// raw CRC (register starts at `c`) of smem bytes [p, p + n)
__device__ __forceinline__ uint32_t ing_crc_span(const uint32_t* tabs, uint32_t copy, uint32_t c, const uint8_t* p, uint32_t n) {
  while (n && (reinterpret_cast<uintptr_t>(p) & 3)) { c = ing_crc_byte(tabs, copy, c, *p++); n--; }
  const uint32_t* w = reinterpret_cast<const uint32_t*>(p);
  uint32_t nw = n >> 2;
  // two words in flight: the loads do not depend on the register
  while (nw >= 2) {
    const uint32_t a = w[0], b = w[1];
    c = ing_crc_word(tabs, copy, c, a);
    c = ing_crc_word(tabs, copy, c, b);
    w += 2; nw -= 2;
  }
  if (nw) { c = ing_crc_word(tabs, copy, c, *w++); }
  p = reinterpret_cast<const uint8_t*>(w);
  for (uint32_t i = 0; i < (n & 3); i++) c = ing_crc_byte(tabs, copy, c, p[i]);
  return c;
}
#undef ING_TAB

__global__ void __launch_bounds__(ING_THREADS, 3) k_ingest(IngestView V, JobDev* J) {
  extern __shared__ __align__(16) uint8_t ing_smem[];
  uint8_t* buf = ing_smem + 16;                                               // 16 B front pad (walk 2 reads up to 12 B in front of the block), ING_BUF, 16 B back pad
  uint32_t* tabs = reinterpret_cast<uint32_t*>(ing_smem + ING_BUF + 32);      // 4 * 256 * ING_REP words
  IngEntry* etab = reinterpret_cast<IngEntry*>(tabs + 4 * 256 * ING_REP);     // ING_MAXE
  uint32_t* ecrc = reinterpret_cast<uint32_t*>(etab + ING_MAXE);              // ING_MAXE value CRCs
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t sh_gb, sh_n, sh_base, sh_bad, sh_crc_acc;
  __shared__ int sh_fb;
  __shared__ uint32_t sh_wsum[ING_THREADS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t copy = threadIdx.x & (ING_REP - 1);
  for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) {
    const uint32_t v = (&g_crc_tab[0][0])[i];
#pragma unroll
    for (int c = 0; c < ING_REP; c++) tabs[i * ING_REP + c] = v;
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint32_t total_blocks = V.blk_base[V.k];
  const int S = V.S;
  uint32_t parity = 0;

  for (;;) {
    if (threadIdx.x == 0) {
      sh_gb = atomicAdd(V.ticket, 1u);
      sh_bad = 0; sh_crc_acc = 0; sh_fb = 0;
    }
    __syncthreads();
    const uint32_t gb = sh_gb;
    if (gb >= total_blocks) break;
    int run_idx = 0;
    while (V.blk_base[run_idx + 1] <= gb) run_idx++;
    const RunView& run = V.runs[run_idx];
    const uint32_t run_first = V.blk_base[run_idx];
    const uint32_t b = gb - run_first;
    const uint64_t boff = run.blk_off[b];
    const uint32_t size = run.blk_size[b];
    const uint8_t* gsrc = run.data + boff;
    const uint32_t mis = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(gsrc) & 15);
    const uint32_t span = (mis + size + 5 + 15) & ~15u;
    // not for this kernel: the host switches to the general path
    const bool unfit = span > ING_BUF || run.key_encoding != 1;
    if (unfit) {
      // results are discarded once the fallback flag is up; the look-back chain only has to stay alive
      if (threadIdx.x == 0) {
        atomicMax(&J->ingest_fallback, ING_FALLBACK_GENERAL);
        *reinterpret_cast<volatile unsigned long long*>(&V.status[gb]) = 2ull << 32;
      }
      __syncthreads();
      continue;
    }
    if (threadIdx.x == 0) {
      mbar_expect_tx(&bar, span);
      bulk_g2s(buf, gsrc - mis, span, &bar);
    }
    mbar_wait(&bar, parity);
    parity ^= 1;
    const uint8_t* blk = buf + mis;

    // ---- walk 1: lanes own restart intervals; entry table, counts, validation
    const uint32_t nres = ld_u32_unaligned(blk + size - 4);
    const uint32_t ri = J->restart_interval[run_idx] ? J->restart_interval[run_idx] : 0xffffffffu;   // 0: every block has one interval
    uint32_t bad = 0;
    if (nres == 0 || static_cast<uint64_t>(nres) * 4 + 4 > size) bad = DEV_ERR_BAD_BLOCK;
    else if (blk[size] != 0) bad = DEV_ERR_COMPRESSED;
    const uint32_t restarts_off = bad ? 0 : size - 4 - 4 * nres;
    uint32_t my_n = 0, my_maxk = 0;
    int fallback = 0;
    if (!bad) {
      for (uint32_t r = threadIdx.x; r < nres; r += blockDim.x) {
        uint32_t p = ld_u32_unaligned(blk + restarts_off + 4 * r);
        const uint32_t end = (r + 1 < nres) ? ld_u32_unaligned(blk + restarts_off + 4 * (r + 1)) : restarts_off;
        if (p > end || end > restarts_off) { bad = DEV_ERR_BAD_BLOCK; break; }
        uint32_t n = 0, klen = 0;
        const uint64_t slot0 = static_cast<uint64_t>(r) * (ri == 0xffffffffu ? 0u : ri);
        while (p < end) {
          uint32_t shared, non_shared, vlen;
          const int h = parse_entry_header(blk + p, end - p, &shared, &non_shared, &vlen);
          if (!h || shared > klen || (n == 0 && shared != 0) || static_cast<uint64_t>(p) + h + non_shared + vlen > end) { bad = DEV_ERR_BAD_ENTRY; break; }
          klen = shared + non_shared;
          if (klen < 8) { bad = DEV_ERR_SHORT_KEY; break; }
          my_maxk = max(my_maxk, klen);
          const uint64_t slot = slot0 + n;
          if (slot >= ING_MAXE) { fallback = ING_FALLBACK_GENERAL; break; }
          IngEntry e; e.estart = static_cast<uint16_t>(p); e.vstart = static_cast<uint16_t>(p + h + non_shared); e.vlen = vlen;
          etab[slot] = e;
          p += h + non_shared + vlen;
          n++;
        }
        if (bad || fallback) break;
        // every interval but the last of a block is full (BlockBuilder restarts every block_restart_interval entries)
        if (r + 1 < nres) { if (n != ri) { bad = DEV_ERR_IRREGULAR_RESTARTS; break; } }
        else if (n > ri || n == 0) { bad = n ? DEV_ERR_IRREGULAR_RESTARTS : DEV_ERR_BAD_ENTRY; break; }
        my_n += n;
      }
    }
    if (my_maxk > 16 * ING_NVI) fallback = ING_FALLBACK_GENERAL;
    else if (my_maxk > static_cast<uint32_t>(S) - 8) fallback = max(fallback, ING_FALLBACK_WIDER);      // user key longer than S - 16
    if (bad) atomicMax(&sh_bad, bad);
    if (fallback) { atomicMax(&J->ingest_fallback, fallback); atomicMax(&sh_fb, fallback); }
    // a flag raised by another CTA: this block's results are discarded anyway (one uniform decision per CTA)
    if (threadIdx.x == 0) { const int g = *reinterpret_cast<volatile int*>(&J->ingest_fallback); if (g) atomicMax(&sh_fb, g); }
    uint32_t wn = __reduce_add_sync(0xffffffffu, my_n);
    const uint32_t wk = __reduce_max_sync(0xffffffffu, my_maxk);
    if (lane == 0) { sh_wsum[wid] = wn; if (wk > __ldcg(&J->max_ikey_len)) atomicMax(&J->max_ikey_len, wk); }
    __syncthreads();
    if (sh_bad) { if (threadIdx.x == 0) dev_fail(J, sh_bad, b); }
    uint32_t n_ent = 0;
#pragma unroll
    for (int q = 0; q < ING_THREADS / 32; q++) n_ent += sh_wsum[q];
    if (sh_bad || sh_fb) n_ent = 0;
    // ---- publish this block's count, so that successors can look back while the CRCs are computed
    if (threadIdx.x == 0) {
      const unsigned long long flag = gb == run_first ? 2ull : 1ull;
      *reinterpret_cast<volatile unsigned long long*>(&V.status[gb]) = (flag << 32) | n_ent;
    }

    // ---- CRCs: one thread per entry
    const uint32_t L = size + 1;                       // contents + type byte
    uint32_t acc = 0;
    for (uint32_t e = threadIdx.x; e < n_ent; e += blockDim.x) {
      const IngEntry en = etab[e];
      const uint32_t vc = ing_crc_span(tabs, copy, 0u, blk + en.vstart, en.vlen);
      ecrc[e] = vc;
      if (V.verify) {
        const uint32_t gc = ing_crc_span(tabs, copy, 0u, blk + en.estart, en.vstart - en.estart);
        const uint32_t vend = en.vstart + en.vlen;
        uint32_t E = gc ? crc_mulmod(__ldg(&g_crc_xpow8[en.vlen]), gc) : 0u;      // en.vlen < 64 K: inside the table
        E ^= vc;
        if (E) acc ^= crc_mulmod(__ldg(&g_crc_xpow8[L - vend]), E);
      }
    }
    if (V.verify && n_ent) {
      acc = __reduce_xor_sync(0xffffffffu, acc);
      if (lane == 0 && acc) atomicXor(&sh_crc_acc, acc);
      __syncthreads();
      if (threadIdx.x == 0) {
        // tail: restart array, restart count, type byte; then the 0xffffffff initial register and the final complement
        const IngEntry last = etab[n_ent - 1];
        const uint32_t tail0 = last.vstart + last.vlen;
        uint32_t r = sh_crc_acc ^ ing_crc_span(tabs, copy, 0u, blk + tail0, L - tail0);
        r ^= crc_mulmod(L <= CRC_XPOW_TABLE ? g_crc_xpow8[L] : crc_xpow_bytes(L, g_crc_x2n), 0xffffffffu);
        const uint32_t crc = crc_mask(~r);
        if (crc != ld_u32_unaligned(blk + size + 1)) dev_fail(J, DEV_ERR_BAD_CRC, b);
      }
    }

    // ---- look-back: entries of this file in front of the block
    if (wid == 0) {
      unsigned long long excl = 0;
      if (gb != run_first) {
        long long hi = static_cast<long long>(gb) - 1;
        for (;;) {
          const long long idx = hi - lane;
          unsigned long long s = (2ull << 32);                                   // in front of the run: prefix 0
          if (idx >= static_cast<long long>(run_first)) s = *reinterpret_cast<volatile unsigned long long*>(&V.status[idx]);
          const uint32_t flag = static_cast<uint32_t>(s >> 32);
          const uint32_t ballot_x = __ballot_sync(0xffffffffu, flag == 0);
          const uint32_t ballot_p = __ballot_sync(0xffffffffu, flag == 2);
          const int first_p = ballot_p ? __ffs(ballot_p) - 1 : 32;
          const uint32_t upto = first_p >= 31 ? 0xffffffffu : ((2u << first_p) - 1u);
          if (ballot_x & upto) continue;                                        // a needed predecessor has not published yet
          const uint32_t contrib = lane <= first_p ? static_cast<uint32_t>(s) : 0u;
          excl += __reduce_add_sync(0xffffffffu, contrib);
          if (first_p < 32) break;
          hi -= 32;
        }
        if (lane == 0) *reinterpret_cast<volatile unsigned long long*>(&V.status[gb]) = (2ull << 32) | ((excl + n_ent) & 0xffffffffull);
      }
      if (lane == 0) {
        sh_base = static_cast<uint32_t>(excl);
        if (gb + 1 == V.blk_base[run_idx + 1]) V.totals[run_idx] = static_cast<uint32_t>(excl + n_ent);
        if (excl + n_ent > V.cap[run_idx]) { dev_fail(J, DEV_ERR_BAD_BLOCK, b); sh_n = 0; } else sh_n = n_ent;
      }
    }
    __syncthreads();
    const uint32_t base = sh_base;
    const uint32_t n_out = sh_n;
    // value offsets and value CRCs: coalesced
    for (uint32_t e = threadIdx.x; e < n_out; e += blockDim.x) {
      run.val_off[base + e] = boff + etab[e].vstart;
      run.val_crc[base + e] = ecrc[e];
    }
    // ---- walk 2: rebuild the internal keys in registers, write the records
    if (n_out) {
      for (uint32_t r = threadIdx.x; r < nres; r += blockDim.x) {
        uint32_t p = ld_u32_unaligned(blk + restarts_off + 4 * r);
        const uint32_t end = (r + 1 < nres) ? ld_u32_unaligned(blk + restarts_off + 4 * (r + 1)) : restarts_off;
        uint32_t idx = base + r * (ri == 0xffffffffu ? 0u : ri);
        uint4 kv[ING_NVI];
#pragma unroll
        for (int w = 0; w < ING_NVI; w++) kv[w] = make_uint4(0, 0, 0, 0);
        while (p < end) {
          uint32_t shared, non_shared, vlen;
          const int h = parse_entry_header(blk + p, end - p, &shared, &non_shared, &vlen);
          p += h;
          const uint32_t klen = shared + non_shared;
          const uint32_t ulen = klen - 8;
#pragma unroll
          for (int w = 0; w < ING_NVI; w++) {
            const int lo = 16 * w;
            if (lo + 16 <= static_cast<int>(shared)) continue;
            if (lo >= static_cast<int>(klen)) { kv[w] = make_uint4(0, 0, 0, 0); continue; }
            // 16 key-delta bytes starting at blk + p + lo - shared (any alignment, shared memory)
            const uint8_t* src = blk + p + lo - static_cast<int>(shared);
            const uint32_t sh = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(src) & 3);
            const uint32_t* sa = reinterpret_cast<const uint32_t*>(src - sh);
            const uint32_t w0 = sa[0], w1 = sa[1], w2 = sa[2], w3 = sa[3], w4 = sa[4];
            const uint32_t bits = sh * 8;
            const uint4 nw = make_uint4(__funnelshift_r(w0, w1, bits), __funnelshift_r(w1, w2, bits), __funnelshift_r(w2, w3, bits), __funnelshift_r(w3, w4, bits));
            const uint4 keep = low_bytes_mask16(static_cast<int>(shared) - lo);
            const uint4 valid = low_bytes_mask16(static_cast<int>(klen) - lo);
            kv[w].x = (kv[w].x & keep.x) | (nw.x & ~keep.x & valid.x);
            kv[w].y = (kv[w].y & keep.y) | (nw.y & ~keep.y & valid.y);
            kv[w].z = (kv[w].z & keep.z) | (nw.z & ~keep.z & valid.z);
            kv[w].w = (kv[w].w & keep.w) | (nw.w & ~keep.w & valid.w);
          }
          p += non_shared;
          // suffix = internal-key bytes [ulen, ulen + 8)
          uint4 va = kv[0], vb = kv[1];
#pragma unroll
          for (int w = 1; w < ING_NVI; w++) if (static_cast<int>(ulen >> 4) == w) { va = kv[w]; vb = (w + 1 < ING_NVI) ? kv[w + 1] : make_uint4(0, 0, 0, 0); }
          uint32_t s0, s1;
          {
            uint32_t w0 = va.x, w1 = va.y, w2 = va.z, w3 = va.w, w4 = vb.x, w5 = vb.y;
            const uint32_t sh = ulen & 15, qq = sh >> 2, bits = (sh & 3) * 8;
            if (qq & 1) { w0 = w1; w1 = w2; w2 = w3; w3 = w4; w4 = w5; }
            if (qq & 2) { w0 = w2; w1 = w3; w2 = w4; }
            s0 = __funnelshift_r(w0, w1, bits); s1 = __funnelshift_r(w1, w2, bits);
          }
          uint8_t flags = 0;
          if (run.ht_filter != 0xfffffffffffffffeull || V.range) {
            // rare: per-file HybridTime filter / key range of a subcompaction — the user key as bytes
            __align__(16) uint8_t kb[16 * ING_NVI];
#pragma unroll
            for (int w = 0; w < ING_NVI; w++) reinterpret_cast<uint4*>(kb)[w] = kv[w];
            if (run.ht_filter != 0xfffffffffffffffeull) {
              const uint32_t htl = doc_ht_len_from_end(kb, ulen);
              uint64_t ht;
              if (htl && doc_ht_decode(kb + ulen - htl, htl, &ht) && ht > run.ht_filter) flags |= REC_F_HT_FILTERED;
            }
            if (V.range) {
              if (V.range->lower_len && cmp_raw(kb, ulen, V.range->lower, V.range->lower_len) < 0) flags |= REC_F_OUT_OF_RANGE;
              if (V.range->upper_len && cmp_raw(kb, ulen, V.range->upper, V.range->upper_len) >= 0) flags |= REC_F_OUT_OF_RANGE;
            }
          }
          uint8_t* rec = run.rec + static_cast<size_t>(idx) * S;
          const int key_vecs = (S - 16) >> 4;
#pragma unroll
          for (int w = 0; w < ING_NVI; w++) {
            if (w < key_vecs) {
              const uint4 m = low_bytes_mask16(static_cast<int>(ulen) - 16 * w);
              reinterpret_cast<uint4*>(rec)[w] = make_uint4(kv[w].x & m.x, kv[w].y & m.y, kv[w].z & m.z, kv[w].w & m.w);
            }
          }
          const uint8_t vfirst = vlen ? blk[p] : 0;
          uint4 tr;
          tr.x = s0; tr.y = s1;
          tr.z = ulen | (static_cast<uint32_t>(vfirst) << 16) | (static_cast<uint32_t>(flags) << 24);
          tr.w = vlen;
          *reinterpret_cast<uint4*>(rec + S - 16) = tr;
          p += vlen;
          idx++;
        }
      }
    }
    __syncthreads();        // every read of the staging buffer is done before the next bulk copy lands in it
  }
}

// General path: RAW CRC32C of every value, one thread per entry, straight from the data file in HBM.
__global__ void __launch_bounds__(256) k_value_crc(const RunView* runs, int k, int S) {
  __shared__ uint32_t tabs[4 * 256 * ING_REP];
  for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) {
    const uint32_t v = (&g_crc_tab[0][0])[i];
#pragma unroll
    for (int c = 0; c < ING_REP; c++) tabs[i * ING_REP + c] = v;
  }
  __syncthreads();
  const uint32_t copy = threadIdx.x & (ING_REP - 1);
  for (int r = 0; r < k; r++) {
    const RunView& run = runs[r];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < run.n_entries; i += gridDim.x * blockDim.x) {
      const uint8_t* rec = run.rec + static_cast<size_t>(i) * S;
      uint32_t n = rec_vlen(rec, S);
      const uint8_t* p = run.data + run.val_off[i];
      uint32_t c = 0;
      while (n && (reinterpret_cast<uintptr_t>(p) & 3)) { c = tabs[(((c ^ __ldg(p)) & 0xff)) * ING_REP + copy] ^ (c >> 8); p++; n--; }
      const uint32_t* w = reinterpret_cast<const uint32_t*>(p);
      for (uint32_t q = 0; q < (n >> 2); q++) {
        c ^= __ldg(w + q);
        c = tabs[((3 << 8) + (c & 0xff)) * ING_REP + copy] ^ tabs[((2 << 8) + ((c >> 8) & 0xff)) * ING_REP + copy] ^
            tabs[((1 << 8) + ((c >> 16) & 0xff)) * ING_REP + copy] ^ tabs[(c >> 24) * ING_REP + copy];
      }
      p += n & ~3u;
      for (uint32_t q = 0; q < (n & 3); q++) c = tabs[(((c ^ __ldg(p + q)) & 0xff)) * ING_REP + copy] ^ (c >> 8);
      run.val_crc[i] = c;
    }
  }
}

}  // namespace ybgpu
