"""Lane-level model of k_snappy_compress (yugabyte-db_b200/csrc/snappy_kernels.cuh), statement for statement: 32
positions are tried at once; a lane's match candidate is the nearest lower lane with the same hash or else the slot's
occupant; the lowest matching lane wins; lanes up to it take their slots. tests/test_oracle_codec.py checks that this
order of events is the scalar encoder's (oracle_sst.cc SnappyCompress), i.e. that the kernel's design is right before
any GPU is involved; tests/test_gpu_parity.py checks the kernel itself. Returns the stream, or None when the encoder
gives up because the block would not shrink by 12.5 % (GoodCompressionRatio)."""
import struct

def clz32(x): return 32 - x.bit_length()
def ffs(x): return (x & -x).bit_length()

def literal_header(L):
    l1 = L - 1
    if l1 < 60: return 1, l1 << 2
    if l1 < 256: return 2, (60 << 2) | (l1 << 8)
    return 3, (61 << 2) | (l1 << 8)

def warp_compress(raw, hash_bits=12, fragment=65536):
    n = len(raw)
    out = bytearray(n + 16)
    limit = n - n // 8
    give_up = n >= 0x7fffffff
    op = 0
    if not give_up:
        v = n
        while v >= 128:
            out[op] = (v | 128) & 0xff; v >>= 7; op += 1
        out[op] = v & 0xff; op += 1
        if op >= limit: give_up = True
    fs = 0
    while fs < n and not give_up:
        f = raw[fs:fs + fragment]; m = len(f)
        T = [0] * (1 << hash_bits)
        G = [0] * (1 << hash_bits)                       # eight more bits of the occupant's hash product
        lit = 0; i = 0
        broke = False
        while i + 4 <= m:
            pos = [i + l for l in range(32)]
            act = [p + 4 <= m for p in pos]
            w = [0] * 32; h = [0x10000 + l for l in range(32)]; tag = [0] * 32
            for l in range(32):
                if act[l]:
                    w[l] = struct.unpack_from("<I", f, pos[l])[0]
                    prod = (w[l] * 0x1e35a7bd) & 0xffffffff
                    h[l] = prod >> (32 - hash_bits); tag[l] = (prod >> (24 - hash_bits)) & 0xff
            grp = [sum(1 << k for k in range(32) if h[k] == h[l]) for l in range(32)]
            cand = [0] * 32; hit = [False] * 32
            for l in range(32):
                lower = grp[l] & ((1 << l) - 1)
                nearest = 31 - clz32(lower) if lower else l
                w_nearest = w[nearest]
                if act[l]:
                    if lower:
                        cand[l] = i + nearest; hit[l] = w_nearest == w[l]
                    else:
                        cand[l] = T[h[l]]
                        hit[l] = cand[l] < pos[l] and G[h[l]] == tag[l] and struct.unpack_from("<I", f, cand[l])[0] == w[l]
            hits = sum(1 << l for l in range(32) if hit[l])
            upto = ffs(hits) - 1 if hits else 31
            writes = {}
            for l in range(32):
                if act[l] and l <= upto:
                    g = grp[l] & (0xffffffff >> (31 - upto))
                    if 31 - clz32(g) == l:
                        assert h[l] not in writes
                        writes[h[l]] = (pos[l] & 0xffff, tag[l])
            for k, v in writes.items(): T[k], G[k] = v
            if not hits:
                i += 32; continue
            mpos = i + upto; c = cand[upto]
            ln = 4
            while True:
                differ = 0
                for l in range(32):
                    q = mpos + ln + l
                    same = q < m and f[c + ln + l] == f[q]
                    if not same: differ |= 1 << l
                if differ:
                    ln += ffs(differ) - 1; break
                ln += 32
            if mpos > lit:
                L = mpos - lit
                hb, hdr = literal_header(L)
                if op + hb + L >= limit: give_up = True; broke = True; break
                for l in range(hb): out[op + l] = (hdr >> (8 * l)) & 0xff
                out[op + hb:op + hb + L] = f[lit:mpos]
                op += hb + L
            off = mpos - c
            nfull = (ln - 68) // 64 + 1 if ln >= 68 else 0
            left = ln - 64 * nfull
            closing = 0; cb = 0
            while left:
                l_ = min(left, 64)
                if left > l_ and left - l_ < 4: l_ = left - 4
                if l_ <= 11 and off < 2048:
                    closing |= (1 | ((l_ - 4) << 2) | ((off >> 8) << 5) | ((off & 0xff) << 8)) << (8 * cb); cb += 2
                else:
                    closing |= (2 | ((l_ - 1) << 2) | (off << 8)) << (8 * cb); cb += 3
                left -= l_
            if op + 3 * nfull + cb >= limit: give_up = True; broke = True; break
            for j in range(nfull):
                out[op + 3 * j] = 2 | (63 << 2); out[op + 3 * j + 1] = off & 0xff; out[op + 3 * j + 2] = off >> 8
            op += 3 * nfull
            for l in range(cb): out[op + l] = (closing >> (8 * l)) & 0xff
            op += cb
            i = mpos + ln; lit = i
        if not give_up and m > lit:
            L = m - lit
            hb, hdr = literal_header(L)
            if op + hb + L >= limit:
                give_up = True; break
            for l in range(hb): out[op + l] = (hdr >> (8 * l)) & 0xff
            out[op + hb:op + hb + L] = f[lit:m]
            op += hb + L
        fs += fragment
    keep = (not give_up) and op < limit
    return bytes(out[:op]) if keep else None

