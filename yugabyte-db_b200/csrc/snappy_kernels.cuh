// snappy_kernels.cuh — Snappy-compressed input data blocks.
//
// DocDB's production default is kSnappyCompression (docdb_rocksdb_util.cc:184): a data block is stored compressed
// when that saves at least 12.5 % (block_based_table_builder.cc:109-131), its trailer's type byte says so, and
// ReadBlock uncompresses it after the checksum check (table/format.cc:441-500 UncompressBlockContents). Here the
// compressed blocks of all inputs are uncompressed once, on the GPU, into a second image of the input files (blocks
// that were stored raw are copied); every later kernel runs on that image unchanged.
//
// Format (snappy's format_description.txt; the library itself lives in yugabyte-db-thirdparty and is not vendored):
// varint32 uncompressed length, then elements — tag & 3 == 0: literal (length - 1 in the tag's upper six bits, or in
// the 1..4 following bytes for 60..63), 1: copy with 11-bit offset and length 4..11, 2 / 3: copy with 16- / 32-bit
// offset and length 1..64. A copy may overlap its own output (run-length patterns).
//
// One warp per block. Every lane parses the element stream redundantly (all lanes read the same tag bytes: one
// broadcast load), so the control flow is uniform and nothing is shuffled; the bytes of a literal or a copy are
// spread over the lanes. A copy whose offset is at least 32 proceeds in rounds of 32 bytes (a round only reads what
// earlier rounds or elements wrote), a closer one is a repeating pattern of bytes that were written before the
// element began. __syncwarp() orders the lanes' global stores and loads between rounds.
//
// Included by engine.cu only.
#pragma once

namespace ybgpu {

struct SnapView {
  const RunView* runs;
  const uint32_t* blk_base;          // [k+1]
  unsigned long long* out_off;       // [total blocks + 1] sizes (contents + 5-byte trailer) -> exclusive prefix -> offsets
  uint32_t* usize;                   // [total blocks] uncompressed contents size
  uint8_t* out;                      // the uncompressed image (k_snappy_decode)
  int k;
};

__device__ __forceinline__ int snap_varint32(const uint8_t* p, uint32_t avail, uint32_t* v) {
  uint32_t r = 0;
  for (int i = 0; i < 5 && static_cast<uint32_t>(i) < avail; i++) {
    const uint32_t b = p[i];
    r |= (b & 127) << (7 * i);
    if (!(b & 128)) { *v = r; return i + 1; }
  }
  return 0;
}

// Per block: the size of its uncompressed contents.
__global__ void __launch_bounds__(256) k_snappy_sizes(SnapView V, JobDev* J) {
  const uint32_t total = V.blk_base[V.k];
  for (uint32_t gb = blockIdx.x * blockDim.x + threadIdx.x; gb < total; gb += gridDim.x * blockDim.x) {
    int r = 0;
    while (V.blk_base[r + 1] <= gb) r++;
    const RunView& run = V.runs[r];
    const uint32_t b = gb - V.blk_base[r];
    const uint8_t* blk = run.data + run.blk_off[b];
    const uint32_t size = run.blk_size[b];
    const uint8_t type = blk[size];
    uint32_t u = size;
    if (type == 1) {
      if (!snap_varint32(blk, size, &u) || u >= (1u << 30)) { dev_fail(J, DEV_ERR_BAD_BLOCK, b); u = 0; }
    } else if (type != 0) {
      dev_fail(J, DEV_ERR_COMPRESSED, b);
    }
    V.usize[gb] = u;
    V.out_off[gb] = static_cast<unsigned long long>(u) + 5;
  }
}

__global__ void __launch_bounds__(128) k_snappy_decode(SnapView V, JobDev* J) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t total = V.blk_base[V.k];
  for (uint32_t gb = warp; gb < total; gb += nwarps) {
    int r = 0;
    while (V.blk_base[r + 1] <= gb) r++;
    const RunView& run = V.runs[r];
    const uint32_t b = gb - V.blk_base[r];
    const uint8_t* src = run.data + run.blk_off[b];
    const uint32_t size = run.blk_size[b];
    const uint32_t ulen = V.usize[gb];
    uint8_t* dst = V.out + V.out_off[gb];
    const uint8_t type = src[size];
    if (type == 0) {
      for (uint32_t i = lane; i < size; i += 32) dst[i] = src[i];
    } else if (type == 1) {
      uint32_t u;
      uint32_t ip = static_cast<uint32_t>(snap_varint32(src, size, &u));
      uint32_t op = 0;
      bool bad = ip == 0;
      while (!bad && ip < size) {
        const uint32_t tag = src[ip++];
        const uint32_t kind = tag & 3;
        if (kind == 0) {
          uint32_t len = tag >> 2;
          if (len >= 60) {
            const uint32_t nb = len - 59;
            if (size - ip < nb) { bad = true; break; }
            len = 0;
            for (uint32_t i = 0; i < nb; i++) len |= static_cast<uint32_t>(src[ip + i]) << (8 * i);
            ip += nb;
          }
          len += 1;
          if (size - ip < len || ulen - op < len) { bad = true; break; }
          for (uint32_t i = lane; i < len; i += 32) dst[op + i] = src[ip + i];
          ip += len; op += len;
        } else {
          uint32_t len, off;
          if (kind == 1) { if (ip >= size) { bad = true; break; } len = 4 + ((tag >> 2) & 7); off = ((tag >> 5) << 8) | src[ip]; ip += 1; }
          else if (kind == 2) { if (size - ip < 2) { bad = true; break; } len = 1 + (tag >> 2); off = src[ip] | (static_cast<uint32_t>(src[ip + 1]) << 8); ip += 2; }
          else {
            if (size - ip < 4) { bad = true; break; }
            len = 1 + (tag >> 2);
            off = src[ip] | (static_cast<uint32_t>(src[ip + 1]) << 8) | (static_cast<uint32_t>(src[ip + 2]) << 16) | (static_cast<uint32_t>(src[ip + 3]) << 24);
            ip += 4;
          }
          if (off == 0 || off > op || ulen - op < len) { bad = true; break; }
          __syncwarp();                                        // what earlier elements wrote is visible to every lane
          if (off >= 32) {
            for (uint32_t base = 0; base < len; base += 32) {
              const uint32_t i = base + lane;
              if (i < len) dst[op + i] = dst[op - off + i];
              __syncwarp();
            }
          } else {
            for (uint32_t i = lane; i < len; i += 32) dst[op + i] = dst[op - off + (i % off)];
          }
          op += len;
        }
      }
      __syncwarp();
      if (bad || op != ulen) { if (lane == 0) dev_fail(J, DEV_ERR_BAD_BLOCK, b); }
    }
    // the image's trailer: stored raw; the checksum of the stored (compressed) bytes was verified before
    if (lane < 5) dst[ulen + lane] = 0;
  }
}

}  // namespace ybgpu
