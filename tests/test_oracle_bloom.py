"""Oracle pins for the bloom-filter side of the SST writer: the LevelDB hash,
FixedSizeFilterBitsBuilder, the DocKeyV3Filter key transformer and the filter blocks / filter index /
properties BlockBasedTableBuilder writes into the metadata file. Replays the reference's own tests
(rocksdb/table/fixed_size_filter_block_test.cc:39-125, docdb/docdb_filter_policy-test.cc:42-71)."""
import ctypes as C
import struct

import dockv_util as dk
import oracle_py as o

L = o.lib()
L.orc_bloom_hash.restype = C.c_uint32
L.orc_bloom_hash.argtypes = [C.c_char_p, C.c_uint64]
L.orc_docdb_filter_prefix.restype = C.c_uint64
L.orc_docdb_filter_prefix.argtypes = [C.c_char_p, C.c_uint64]
L.orc_fixed_size_filter.restype = C.c_uint64
L.orc_fixed_size_filter.argtypes = [C.c_uint64, C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64)]
L.orc_filter_may_match.restype = C.c_int
L.orc_filter_may_match.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64]

DEFAULT_BITS = 65536          # FilterPolicy::kDefaultFixedSizeFilterBits (filter_policy.h:179)


def bloom_hash(b):
    return L.orc_bloom_hash(b, len(b))


def build_filter(keys, total_bits=DEFAULT_BITS):
    blob = b"".join(struct.pack("<I", len(k)) + k for k in keys)
    out = C.create_string_buffer(total_bits // 8 + 1024)
    params = (C.c_uint64 * 3)()
    n = L.orc_fixed_size_filter(total_bits, blob, len(keys), out, len(out), params)
    return out.raw[:n], tuple(params)


def may_match(f, k):
    return bool(L.orc_filter_may_match(f, len(f), k, len(k)))


def py_hash(data, seed=0xbc9f1d34):
    """util/hash.cc:32-75 restated independently (signed tail bytes)."""
    m, M = 0xc6a4a793, 0xffffffff
    h = (seed ^ (len(data) * m)) & M
    i = 0
    while i + 4 <= len(data):
        h = (h + struct.unpack_from("<I", data, i)[0]) & M
        h = (h * m) & M
        h ^= h >> 16
        i += 4
    rest = data[i:]
    sc = lambda b: b - 256 if b >= 128 else b
    if len(rest) == 3:
        h = (h + (sc(rest[2]) << 16)) & M
    if len(rest) >= 2:
        h = (h + (sc(rest[1]) << 8)) & M
    if len(rest) >= 1:
        h = (h + sc(rest[0])) & M
        h = (h * m) & M
        h ^= h >> 24
    return h


def test_leveldb_hash_known_answers():
    # LevelDB util/hash_test.cc known answers that do not depend on the signed/unsigned tail quirk
    assert bloom_hash(b"") == 0xbc9f1d34
    assert bloom_hash(bytes([0x62])) == 0xef1345c4
    assert bloom_hash(bytes([0xe1, 0x80, 0xb9, 0x32])) == 0xed21633a
    # tail bytes >= 0x80 are sign-extended in this tree (util/hash.cc:46-66 keeps the on-disk quirk)
    for d in (bytes([0xc3, 0x97]), bytes([0xe2, 0x99, 0xa5]), b"hello world!", bytes(range(200, 255))):
        assert bloom_hash(d) == py_hash(d)
    assert bloom_hash(bytes([0xc3, 0x97])) != 0x5b663814      # what the unsigned variant would give


def test_fixed_size_filter_geometry():
    _, (max_keys, num_lines, num_probes) = build_filter([], 64 * 1024 * 8)
    # bloom.cc:389-415: 1024 lines -> odd (1023); 6 probes at 1 % error; ~54.6k keys per 64 KB block
    assert num_lines == 1023 and num_probes == 6 and max_keys == int(1023 * 512 * 0.4804530139182014 / 4.605170185988091)
    f, (mk, nl, npb) = build_filter([], DEFAULT_BITS)
    assert nl == 127 and len(f) == 127 * 64 + 5 and f[-5] == npb == 6 and struct.unpack("<I", f[-4:])[0] == 127
    _, (_, nl_small, _) = build_filter([], 64 * 8 * 4)       # < 4096 bytes and even -> one line more
    assert nl_small == 5


def test_fixed_size_filter_single_chunk():
    f, _ = build_filter([b"foo", b"bar", b"box", b"hello"])
    for k in (b"foo", b"bar", b"box", b"hello"):
        assert may_match(f, k)
    assert not may_match(f, b"missing") and not may_match(f, b"other")


def test_fixed_size_filter_multiple_chunks():
    f1, _ = build_filter([b"a1", b"b1", b"c1", b"foo", b"bar"])
    f2, _ = build_filter([b"a2", b"b2", b"c2", b"foo", b"bar"])
    f3, _ = build_filter([])
    for k in (b"a1", b"b1", b"c1", b"foo", b"bar"):
        assert may_match(f1, k)
    for k in (b"a2", b"b2", b"c2", b"missing", b"other"):
        assert not may_match(f1, k)
    for k in (b"a2", b"b2", b"c2", b"foo", b"bar"):
        assert may_match(f2, k)
    for k in (b"a1", b"b1", b"c1", b"missing", b"other"):
        assert not may_match(f2, k)
    for k in (b"foo", b"bar", b"a1", b"b1", b"c1", b"a2", b"b2", b"c2", b"missing", b"other"):
        assert not may_match(f3, k)


def filter_key(user_key):
    return user_key[:L.orc_docdb_filter_prefix(user_key, len(user_key))]


def test_docdb_filter_policy_key_matching():
    def enc(hash_key, range_key=b"range_key", sub_key=b"sub_key", micros=12345):
        return dk.sub_doc_key(dk.doc_key([range_key], hash_code=0, hashed=[hash_key]), [dk.kstr(sub_key)], micros=o.YB_EPOCH_US + micros)
    keys = [b"foo", b"bar", b"test"]
    f, _ = build_filter([filter_key(enc(k)) for k in keys])
    for k in keys:
        assert may_match(f, filter_key(enc(k)))
        assert may_match(f, filter_key(enc(k, b"another_range_key", b"another_sub_key", 55555)))
    assert not may_match(f, filter_key(enc(b"fake")))


def test_docdb_filter_prefix_shapes():
    hashed = dk.doc_key(["r1", "r2"], hash_code=0x1234, hashed=["h1", dk.kint64(7)])
    k = dk.sub_doc_key(hashed, [dk.kcol(3)], micros=o.YB_EPOCH_US + 5)
    assert filter_key(k) == b"G\x12\x34" + dk.kstr("h1") + dk.kint64(7) + b"!"          # up to the hashed group end
    ranged = dk.doc_key(["r1", "r2"])
    k = dk.sub_doc_key(ranged, [dk.kcol(3)], micros=o.YB_EPOCH_US + 5)
    assert filter_key(k) == dk.kstr("r1")                                                  # first range component only
    cot = dk.doc_key(["r1"], hash_code=1, hashed=["h"], cotable=bytes(range(16)))
    assert filter_key(dk.sub_doc_key(cot, [], micros=o.YB_EPOCH_US)) == b"y" + bytes(range(16)) + b"G\x00\x01" + dk.kstr("h") + b"!"
    colo = dk.doc_key(["r1"], colocation=77)
    assert filter_key(dk.sub_doc_key(colo, [], micros=o.YB_EPOCH_US)) == b"0" + (77).to_bytes(4, "big") + dk.kstr("r1")
    assert filter_key(dk.table_tombstone_key(colocation=77, micros=o.YB_EPOCH_US)) == b"0" + (77).to_bytes(4, "big") + b"!"
    assert filter_key(b"plainkey") == b""                                                  # not a DocKey: never added
    assert filter_key(b"") == b""


def test_table_builder_writes_filter_blocks_and_index():
    """Filter blocks are cut every max_keys distinct filter keys, interleaved with the index blocks in the
    metadata file; the reader finds them through metaindex -> filter index."""
    cfg = o.GenConfig(seed=9, num_rows=2500, cols=3, versions=2, num_files=1, value_len=30)
    plain = o.Sst.generate(cfg, 0, o.TableOptions(block_size=1024))
    kvs = plain.read_all()
    topt = o.TableOptions(block_size=1024, index_block_size=512, min_keys_per_index_block=4, filter_policy=1, filter_block_size=256)
    sst = o.Sst.build(kvs, topt)
    assert sst.read_all() == kvs and sst.data == plain.data           # data file unaffected
    props = sst.properties()
    assert props["rocksdb.filter.policy"] == b"DocKeyV3Filter"
    _, (max_keys, _, _) = build_filter([], 256 * 8)
    distinct = []
    for k, _ in kvs:
        fk = filter_key(k[:-8])
        if fk and (not distinct or distinct[-1] != fk):
            distinct.append(fk)
    n_blocks = max(1, -(-len(distinct) // max_keys))
    assert o.varint(props["rocksdb.num.filter.blocks"]) == n_blocks and n_blocks > 3
    filters = sst.filter_blocks()
    assert len(filters) == n_blocks
    for i, fk in enumerate(distinct):
        assert may_match(filters[i // max_keys][1], fk)
    assert o.varint(props["rocksdb.filter.size"]) == sum(len(f) + 5 for _, f in filters)
    # index keys: separators between the last filter key of a block and the first of the next
    for b in range(n_blocks - 1):
        sep = filters[b][0]
        assert distinct[(b + 1) * max_keys - 1] <= sep < distinct[(b + 1) * max_keys]
