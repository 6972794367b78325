// TEST-ONLY: the Snappy kernels of yugabyte-db_b200/csrc/snappy_kernels.cuh — the kernel SOURCE, unchanged — executed
// on the CPU. One warp = 32 cooperative fibers that meet inside every warp-collective intrinsic
// (__ballot_sync, __shfl_sync, __match_any_sync, __syncwarp), so the lanes run in the lock step the kernels are written
// for; __shared__ arrays become statics shared by the 32 threads (one CTA of one warp at a time). The helpers the kernels
// take from engine.cu / encode_kernels.cuh (warp_copy, the strided warp CRC, ldg_u32_unaligned) are plain byte-wise
// stand-ins here: what is under test is the encoder's / decoder's control flow, lane by lane. Not part of the product.
#include <ucontext.h>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <vector>
#include <algorithm>

// One warp = 32 cooperative fibers on one OS thread (ucontext): a lane runs until it reaches a warp-collective intrinsic,
// publishes its operand and yields; when the 32nd lane arrives the collective completes and every lane picks its result
// up. Deterministic, and a "context switch" costs ~100 ns.
namespace emu {
struct Dim { unsigned x = 1, y = 1, z = 1; };
struct Warp {
  ucontext_t main_ctx, lane_ctx[32];
  std::vector<char> stacks[32];
  bool done[32];
  int cur = 0, arrived = 0, alive = 32;
  unsigned generation = 0;
  uint32_t xchg[32];
  std::function<void()> body;
};
inline Warp* g_warp = nullptr;
inline int t_lane = 0;
}  // namespace emu
static emu::Dim threadIdx, blockIdx, blockDim, gridDim;
namespace emu {
inline void yield_lane() { swapcontext(&g_warp->lane_ctx[g_warp->cur], &g_warp->main_ctx); }
// all live lanes meet here
inline void sync() {
  Warp& w = *g_warp;
  const unsigned gen = w.generation;
  if (++w.arrived == w.alive) { w.arrived = 0; w.generation++; return; }
  while (w.generation == gen) yield_lane();
}
inline void lane_entry() {
  Warp& w = *g_warp;
  w.body();
  w.done[w.cur] = true;
  w.alive--;
  // a lane that leaves while others wait at a collective would hang them: the kernels never do that (uniform control flow)
  swapcontext(&w.lane_ctx[w.cur], &w.main_ctx);
}
}  // namespace emu

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __launch_bounds__(...)

static inline uint32_t __ballot_sync(uint32_t, bool pred) {
  emu::g_warp->xchg[emu::t_lane] = pred ? 1u : 0u;
  emu::sync();
  uint32_t m = 0;
  for (int l = 0; l < 32; l++) m |= emu::g_warp->xchg[l] << l;
  emu::sync();
  return m;
}
static inline uint32_t __shfl_sync(uint32_t, uint32_t v, int src) {
  emu::g_warp->xchg[emu::t_lane] = v;
  emu::sync();
  const uint32_t r = emu::g_warp->xchg[src & 31];
  emu::sync();
  return r;
}
static inline uint32_t __match_any_sync(uint32_t, uint32_t v) {
  emu::g_warp->xchg[emu::t_lane] = v;
  emu::sync();
  uint32_t m = 0;
  for (int l = 0; l < 32; l++) m |= (emu::g_warp->xchg[l] == v ? 1u : 0u) << l;
  emu::sync();
  return m;
}
static inline void __syncwarp() { emu::sync(); }
static inline void __syncthreads() { emu::sync(); }                      // one warp per CTA here
static inline int __clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
static inline int __ffs(uint32_t x) { return __builtin_ffs(static_cast<int>(x)); }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) {
  return static_cast<uint32_t>(((static_cast<uint64_t>(hi) << 32) | lo) >> (sh & 31));
}
template <typename T> static inline T __ldg(const T* p) { return *p; }
using std::min;

namespace ybgpu {
// ---- stand-ins for what snappy_kernels.cuh takes from engine.cu / encode_kernels.cuh
struct RunView { const uint8_t* data; const uint64_t* blk_off; const uint32_t* blk_size; };
struct JobDev { int error = 0; uint32_t error_where = 0; };
enum { DEV_ERR_BAD_BLOCK = 3, DEV_ERR_COMPRESSED = 9 };
static inline void dev_fail(JobDev* J, int code, uint32_t where) { if (J->error == 0) { J->error = code; J->error_where = where; } }
static inline uint32_t ldg_u32_unaligned(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline void warp_copy(uint8_t* dst, const uint8_t* src, uint32_t n, int lane) { for (uint32_t i = lane; i < n; i += 32) dst[i] = src[i]; }
static uint32_t g_crc_tab[4][256];
static uint32_t g_crc_s32[4][256];
static uint32_t g_crc_xpow8[256];
static inline uint32_t crc_mask(uint32_t c) { return ((c >> 15) | (c << 17)) + 0xa282ead8u; }
// (the kernel fills its shared byte table with 256 threads per CTA; a CTA is one warp here, so the stand-in reads the
// global table)
static inline uint32_t warp_crc32c_strided(const uint8_t* p, uint64_t len, int, const uint32_t*, const uint32_t (*)[256], uint32_t) {
  uint32_t c = 0xffffffffu;
  for (uint64_t i = 0; i < len; i++) c = g_crc_tab[0][(c ^ p[i]) & 0xff] ^ (c >> 8);
  return ~c;
}
}  // namespace ybgpu

#include "../../yugabyte-db_b200/csrc/snappy_kernels.cuh"

namespace {
template <typename F>
void RunWarp(F&& kernel_body) {           // grid of one CTA of one warp
  emu::Warp* w = new emu::Warp;
  emu::g_warp = w;
  w->body = kernel_body;
  blockIdx = emu::Dim{0, 0, 0}; blockDim = emu::Dim{32, 1, 1}; gridDim = emu::Dim{1, 1, 1};
  for (int l = 0; l < 32; l++) {
    w->done[l] = false;
    w->stacks[l].resize(1 << 20);
    getcontext(&w->lane_ctx[l]);
    w->lane_ctx[l].uc_stack.ss_sp = w->stacks[l].data();
    w->lane_ctx[l].uc_stack.ss_size = w->stacks[l].size();
    w->lane_ctx[l].uc_link = &w->main_ctx;
    makecontext(&w->lane_ctx[l], reinterpret_cast<void (*)()>(emu::lane_entry), 0);
  }
  for (;;) {                               // round robin until every lane has returned
    bool any = false;
    for (int l = 0; l < 32; l++) {
      if (w->done[l]) continue;
      any = true;
      w->cur = l; emu::t_lane = l; threadIdx.x = static_cast<unsigned>(l);
      swapcontext(&w->main_ctx, &w->lane_ctx[l]);
    }
    if (!any) break;
  }
  emu::g_warp = nullptr;
  delete w;
}
void InitCrc() {
  static bool done = false;
  if (done) return;
  done = true;
  for (uint32_t i = 0; i < 256; i++) {
    uint32_t c = i;
    for (int j = 0; j < 8; j++) c = (c >> 1) ^ ((c & 1) ? 0x82F63B78u : 0);
    ybgpu::g_crc_tab[0][i] = c;
  }
}
}  // namespace

extern "C" {

// The engine's output-compression pass over a table of `nblocks` assembled blocks (contents + 5-byte trailer back to back
// at raw_off): k_snappy_compress<variant>, the prefix sum the engine does with its scan kernels, k_snappy_gather.
// out must hold raw_off[nblocks] bytes; final_off gets nblocks + 1 offsets. Returns the final table length.
uint64_t we_compress_table(const uint8_t* raw, const uint64_t* raw_off, uint32_t nblocks, int variant, uint8_t* out, uint64_t* final_off) {
  using namespace ybgpu;
  InitCrc();
  const uint64_t total = raw_off[nblocks];
  std::vector<uint8_t> rawp(total + 128, 0), comp(total + 128, 0);
  memcpy(rawp.data() + 32, raw, total);
  std::vector<unsigned long long> off(raw_off, raw_off + nblocks + 1), fsize(nblocks + 1, 0);
  std::vector<uint32_t> csize(nblocks, 0);
  SnapCompView V{};
  V.raw = rawp.data() + 32; V.raw_off = off.data(); V.comp = comp.data() + 32; V.csize = csize.data(); V.fsize = fsize.data(); V.nblocks = nblocks;
  if (variant == 2) RunWarp([&] { k_snappy_compress<2>(V); });
  else if (variant == 1) RunWarp([&] { k_snappy_compress<1>(V); });
  else RunWarp([&] { k_snappy_compress<0>(V); });
  unsigned long long acc = 0;
  for (uint32_t b = 0; b < nblocks; b++) { const unsigned long long s = fsize[b]; fsize[b] = acc; acc += s; }
  fsize[nblocks] = acc;
  std::vector<uint8_t> outp(acc + 128, 0);
  V.out = outp.data() + 32;
  RunWarp([&] { k_snappy_gather(V); });
  memcpy(out, outp.data() + 32, acc);
  for (uint32_t b = 0; b <= nblocks; b++) final_off[b] = fsize[b];
  return acc;
}

// k_snappy_sizes + k_snappy_decode over the blocks of one file (offsets / sizes exclude the trailer): the uncompressed image
// (contents + 5 zero trailer bytes per block) into out, its block offsets into out_off. Returns the image length, or
// (uint64_t)-1 - error code when the kernels flag a bad block.
uint64_t we_uncompress_table(const uint8_t* data, uint64_t data_len, const uint64_t* blk_off, const uint32_t* blk_size, uint32_t nblocks,
                             uint8_t* out, uint64_t out_cap, uint64_t* out_off) {
  using namespace ybgpu;
  std::vector<uint8_t> in(data_len + 128, 0);
  memcpy(in.data() + 32, data, data_len);
  RunView run{in.data() + 32, blk_off, blk_size};
  uint32_t blk_base[2] = {0, nblocks};
  std::vector<unsigned long long> ooff(nblocks + 1, 0);
  std::vector<uint32_t> usize(nblocks, 0);
  JobDev J;
  SnapView V{};
  V.runs = &run; V.blk_base = blk_base; V.out_off = ooff.data(); V.usize = usize.data(); V.k = 1;
  RunWarp([&] { k_snappy_sizes(V, &J); });
  if (J.error) return ~0ull - static_cast<uint64_t>(J.error);
  unsigned long long acc = 0;
  for (uint32_t b = 0; b < nblocks; b++) { const unsigned long long s = ooff[b]; ooff[b] = acc; acc += s; }
  ooff[nblocks] = acc;
  if (acc > out_cap) return ~0ull;
  std::vector<uint8_t> img(acc + 128, 0xee);
  V.out = img.data() + 32;
  RunWarp([&] { k_snappy_decode(V, &J); });
  if (J.error) return ~0ull - static_cast<uint64_t>(J.error);
  memcpy(out, img.data() + 32, acc);
  for (uint32_t b = 0; b <= nblocks; b++) out_off[b] = ooff[b];
  return acc;
}

}  // extern "C"
