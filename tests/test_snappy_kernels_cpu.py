"""The Snappy kernels' SOURCE (yugabyte-db_b200/csrc/snappy_kernels.cuh) executed on the CPU: tests/host_harness/warp_emu.cc
runs a warp as 32 lock-stepped fibers with the warp-collective intrinsics implemented on top. The GPU suite checks the
same kernels on the device (tests/test_gpu_parity.py); this keeps them under test in a container without one."""
import random

import pytest

import harness_py as h
import oracle_py as o


def _kvs(seed, n, big=False):
    rng = random.Random(seed)
    words = [bytes(rng.randrange(32, 127) for _ in range(rng.randrange(3, 24))) for _ in range(30)]
    kvs = []
    for i in range(n):
        if (i // 40) % 3 == 2:
            v = bytes(rng.randrange(256) for _ in range(rng.randrange(1, 160)))          # stretches that stay raw
        elif i % 17 == 5:
            v = bytes([rng.randrange(256)]) * rng.randrange(1, 700)                      # runs: chains of 64-byte copies
        else:
            v = b" ".join(rng.choice(words) for _ in range(rng.randrange(0, 16)))
        if big and i == n // 2:
            v = (bytes(range(256)) * 700)[:150007]                                       # a block of three 64 KB fragments
        kvs.append((o.ikey(b"row%06d/c%d" % (i // 2, i % 2), 900 + i), v))
    return kvs


def _blocks(t):
    off, sz = t.block_handles()
    return [int(x) for x in off], [int(x) for x in sz]


@pytest.mark.parametrize("variant,seed,n,bs,big", [(0, 1, 700, 2048, False), (1, 2, 500, 4096, False), (2, 3, 300, 1024, False), (0, 4, 60, 2048, True)])
def test_compress_kernels_write_the_oracles_compressed_table(variant, seed, n, bs, big):
    """k_snappy_compress<variant> + k_snappy_gather over an assembled (uncompressed) data file = the data file the
    oracle's builder writes with kSnappyCompression: same stored form per block (compressed or raw by the 12.5 % rule),
    same offsets, same trailers."""
    kvs = _kvs(seed, n, big)
    plain = o.Sst.build(kvs, o.TableOptions(block_size=bs))
    comp = o.Sst.build(kvs, o.TableOptions(block_size=bs, compression=1))
    off, _ = _blocks(plain)
    data, foff = h.warp_compress_table(bytes(plain.data), off, variant)
    coff, csz = _blocks(comp)
    assert foff[:-1] == coff and foff[-1] == len(comp.data)
    assert data == bytes(comp.data)
    types = {data[a + b] for a, b in zip(coff, csz)}
    assert types == ({0, 1} if not big else types) and 1 in types


def test_decode_kernels_rebuild_the_uncompressed_table():
    """k_snappy_sizes + k_snappy_decode: the image of a compressed table is the uncompressed twin's blocks (zeroed
    trailers); streams written by the real snappy library decode too; a malformed stream is flagged, not decoded."""
    kvs = _kvs(7, 600, big=True)
    plain = o.Sst.build(kvs, o.TableOptions(block_size=4096))
    comp = o.Sst.build(kvs, o.TableOptions(block_size=4096, compression=1))
    poff, psz = _blocks(plain)
    coff, csz = _blocks(comp)
    img, ooff = h.warp_uncompress_table(bytes(comp.data), coff, csz)
    pdata = bytes(plain.data)
    assert img == b"".join(pdata[a:a + b] + bytes(5) for a, b in zip(poff, psz))
    assert ooff == [a - 0 for a in poff] + [len(pdata)]
    try:
        import pyarrow as pa
        lib_ok = pa.Codec.is_available("snappy")
    except ImportError:
        lib_ok = False
    if lib_ok:
        raws = [pdata[a:a + b] for a, b in zip(poff, psz)][:40]
        blob, offs, sizes = b"", [], []
        for r in raws:
            c = pa.compress(r, codec="snappy", asbytes=True)
            offs.append(len(blob)); sizes.append(len(c))
            blob += c + b"\x01" + bytes(4)
        img, _ = h.warp_uncompress_table(blob, offs, sizes)
        assert img == b"".join(r + bytes(5) for r in raws)
    a, b = next((a, b) for a, b in zip(coff, csz) if bytes(comp.data)[a + b] == 1)
    bad = bytearray(comp.data)
    bad[a + b // 2] ^= 0xff
    try:
        got, _ = h.warp_uncompress_table(bytes(bad), coff, csz)
    except RuntimeError as e:
        assert e.args[0] == 3                                   # DEV_ERR_BAD_BLOCK
    else:
        assert got != img                                       # (a flipped literal byte is the checksum's to catch, not the decoder's)


def test_decode_kernel_every_element_kind():
    """Hand-made streams with what neither this repository's encoder nor libsnappy emits but the format allows: literal
    lengths in 1..4 trailing bytes, copies with 4-byte offsets, overlapping copies at every offset around a warp's width
    (1, 2, 3, 5, 31, 32, 33, ...) — k_snappy_decode against the oracle's decoder."""
    import struct
    rng = random.Random(9)

    def varint(n):
        out = b""
        while n >= 128:
            out += bytes([(n & 127) | 128])
            n >>= 7
        return out + bytes([n])

    def lit(data, nb):
        l1 = len(data) - 1
        return bytes([l1 << 2]) + data if nb == 0 else bytes([(59 + nb) << 2]) + l1.to_bytes(nb, "little") + data

    blob, offs, sizes, raws = b"", [], [], []
    for _ in range(300):
        out, stream = bytearray(), b""
        for _ in range(rng.randrange(1, 30)):
            k = rng.randrange(5)
            if k == 0 or not out:
                d = bytes(rng.randrange(256) for _ in range(rng.choice([1, 2, 59, 60, 61, 255, 256, 257, 1000])))
                need = 0 if len(d) - 1 < 60 else (1 if len(d) - 1 < 256 else 2)
                stream += lit(d, rng.choice([need] + [nb for nb in (1, 2, 3, 4) if nb >= max(need, 1)]))
                out += d
            else:
                off = min(rng.choice([1, 2, 3, 5, 31, 32, 33, 64, 100, 2047, 2048, 5000]), len(out))
                ln = rng.randrange(4, 12) if k == 1 else rng.randrange(1, 65)
                if k == 1 and off < 2048:
                    stream += bytes([1 | ((ln - 4) << 2) | ((off >> 8) << 5), off & 0xff])
                elif k == 3:
                    stream += bytes([3 | ((ln - 1) << 2)]) + struct.pack("<I", off)
                else:
                    stream += bytes([2 | ((ln - 1) << 2)]) + struct.pack("<H", off)
                for _ in range(ln):
                    out.append(out[len(out) - off])
        c = varint(len(out)) + stream
        assert o.snappy_uncompress(c) == bytes(out)
        offs.append(len(blob))
        sizes.append(len(c))
        raws.append(bytes(out))
        blob += c + b"\x01" + bytes(4)
    img, _ = h.warp_uncompress_table(blob, offs, sizes)
    assert img == b"".join(r + bytes(5) for r in raws)
