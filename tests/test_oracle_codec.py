"""Pins the oracle's codecs against the reference's own golden vectors."""
import random

import pytest

import dockv_util as dk
import oracle_py as o

EPOCH = 1500000000 * 1000000
MAXW = 4294967295


def test_doc_hybrid_time_exact_bytes(oracle):
    # common ... server/doc_hybrid_time-test.cc:95-165 TestExactByteRepresentation
    vec = [
        (b"\x80\x07\xc4e5\xff\x80H", EPOCH + 1000000000, 0, 0),
        (b"\x80\x10\xbd\xbf;-\x03\xdf\xff\xff\xff\xec", EPOCH + 1000000, 1234, MAXW),
        (b"\x80\x10\xbd\xbf;-G", EPOCH + 1000000, 1234, 0),
        (b"\x80\x10\xbd\xbf\x80\x03\xdf\xff\xff\xff\xeb", EPOCH + 1000000, 0, MAXW),
        (b"\x80\x10\xbd\xbf\x80F", EPOCH + 1000000, 0, 0),
        (b"\x80<\x17\x80E", EPOCH + 1000, 0, 0),
        (b"\x80?\x0b=\xbfF", EPOCH, 1000000, 0),
        (b"\x80\x80<\x17E", EPOCH, 1000, 0),
        (b"\x80\x80\x80\x0e\x17\xb7\xc7", EPOCH, 0, 1000000),
        (b"\x80\x80\x80\x1f\x82\xc6", EPOCH, 0, 1000),
        (b"\x80\x80\x80D", EPOCH, 0, 0),
        (b"\x80\xc3\xe8\x80E", EPOCH - 1000, 0, 0),
        (b"\x80\xefB@\x80F", EPOCH - 1000000, 0, 0),
        (b"\x80\xf8;\x9a\xca\x00\x80H", EPOCH - 1000000000, 0, 0),
        (b"\x80\xff\x01\xc6\xbfRc@\x00\x80K", 1000000000000000, 0, 0),
        (b"\x80\xff\x05T=\xf7)\xc0\x00\x80K", EPOCH - 1500000000000000, 0, 0),
    ]
    for exp, micros, logical, wid in vec:
        # logical > 4095 does not fit the 12-bit repr; those vectors use the 3-arg constructor
        # (micros, logical, write_id) whose repr is (micros<<12)+logical; emulate the same sum.
        repr_ = (micros << 12) + logical
        got = oracle.encode_doc_ht_repr(repr_, wid)
        assert got == exp, (micros, logical, wid, got, exp)
        assert got[-1] & 0x1f == len(got)
        assert oracle.decode_doc_ht(got) == (repr_ >> 12, repr_ & 0xfff, wid)


def test_doc_hybrid_time_order_is_reversed(oracle):
    # server/doc_hybrid_time-test.cc:38-81: sgn(ts1 <=> ts2) == -sgn(enc1 <=> enc2)
    rng = random.Random(7)
    pts = [(rng.randrange(EPOCH - 10**9, EPOCH + 10**15), rng.randrange(0, 4096), rng.choice([0, 1, 5, 1000, MAXW]))
           for _ in range(400)]
    enc = [oracle.encode_doc_ht(*p) for p in pts]
    for _ in range(2000):
        i, j = rng.randrange(len(pts)), rng.randrange(len(pts))
        a = (pts[i] > pts[j]) - (pts[i] < pts[j])
        b = (enc[i] > enc[j]) - (enc[i] < enc[j])
        assert a == -b


def test_subdockey_exact_bytes(oracle):
    # dockv/doc_key-test.cc:313-332 TestBasicSubDocKeyEncodingDecoding
    key = dk.sub_doc_key(dk.doc_key(["some_doc_key"]), ["sk1", "sk2", dk.kstr_desc(b"sk3\x00")], micros=1000)
    assert key == (b"Ssome_doc_key\x00\x00!Ssk1\x00\x00Ssk2\x00\x00a\x8c\x94\xcc\xff\xfe\xff\xff"
                   b"#\x80\xff\x05T=\xf7)\xbc\x18\x80K")
    ends = oracle.subdockey_ends(key)
    # [id_end, dockey_end, sk1, sk2, sk3]
    assert ends == [0, 16, 22, 28, 36]
    assert key[ends[-1]:ends[-1] + 1] == b"#"


def test_dockey_exact_bytes(oracle):
    # docdb/docdb-test-wrapper.cc:844-880 BasicTest; dockv/doc_key-test.cc:297-311
    assert dk.doc_key(["my_key_where_value_is_a_string"]) == b"Smy_key_where_value_is_a_string\x00\x00!"
    assert dk.doc_key(["mydockey", dk.INT_KEY1]) == b"Smydockey\x00\x00I\x80\x00\x00\x00\x00\x01\xe2@!"
    k = dk.doc_key(["range1", 1000, "range2", 2000], hash_code=0xcafe, hashed=["hashed1", "hashed2"])
    assert k == (b"G\xca\xfeShashed1\x00\x00Shashed2\x00\x00!Srange1\x00\x00I\x80\x00\x00\x00\x00\x00\x03\xe8"
                 b"Srange2\x00\x00I\x80\x00\x00\x00\x00\x00\x07\xd0!")
    sk = dk.sub_doc_key(k, [dk.kcol(3)], micros=EPOCH + 5)
    assert oracle.subdockey_ends(sk) == [0, len(k), len(k) + 2]


def test_table_tombstone_ends(oracle):
    # dockv/doc_key.cc:973-982: id ! # HT => only the id end is pushed
    k = dk.table_tombstone_key(colocation=0x1234, micros=EPOCH + 7)
    assert oracle.subdockey_ends(k) == [5]
    k = dk.table_tombstone_key(cotable=bytes(range(1, 17)), micros=EPOCH + 7)
    assert oracle.subdockey_ends(k) == [17]
    row = dk.sub_doc_key(dk.doc_key(["r"], colocation=0x1234), [dk.kcol(1)], micros=EPOCH)
    assert oracle.subdockey_ends(row) == [5, 5 + 5, 5 + 5 + 2]


def test_crc32c_standard_vectors(oracle):
    # rocksdb/util/crc32c_test.cc:32-70 (rfc3720 B.4)
    assert oracle.crc32c(bytes(32)) == 0x8a9136aa
    assert oracle.crc32c(b"\xff" * 32) == 0x62a8ab43
    assert oracle.crc32c(bytes(range(32))) == 0x46dd794e
    assert oracle.crc32c(bytes(31 - i for i in range(32))) == 0x113fdb5c
    data = bytes([0x01, 0xc0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0x14, 0, 0, 0, 0, 0, 0x04, 0,
                  0, 0, 0, 0x14, 0, 0, 0, 0x18, 0x28, 0, 0, 0, 0, 0, 0, 0, 0x02, 0, 0, 0, 0, 0, 0, 0])
    assert oracle.crc32c(data) == 0xd9963a56
    L = oracle.lib()
    c = oracle.crc32c(b"foo")
    assert L.orc_crc32c_mask(c) != c


def test_fast_varint_known_answers(oracle):
    # util/fast_varint-test.cc:138-143
    assert oracle.signed_varint(0) == b"\x80"
    assert oracle.signed_varint(1) == b"\x81"
    assert oracle.signed_varint(-1) == b"~"
    assert oracle.signed_varint(64) == b"\xc0\x40"
    assert oracle.signed_varint(8191) == b"\xdf\xff"
    import ctypes as C
    L = oracle.lib()
    vals = [0, 1, -1, 63, 64, -64, -65, 8191, 8192, 2**62, -(2**62), 2**63 - 1, -(2**63) + 1, -(2**63)]
    rng = random.Random(3)
    vals += [rng.randrange(-(2**63), 2**63) >> rng.randrange(0, 63) for _ in range(2000)]
    prev = None
    for v in sorted(vals):
        e = oracle.signed_varint(v)
        out = C.c_int64()
        assert L.orc_decode_signed_varint(e, len(e), C.byref(out)) == len(e)
        assert out.value == v
        if prev is not None and prev[0] != v:
            assert prev[1] < e          # order preserving (fast_varint-test.cc lexicographic check)
        prev = (v, e)


VARINT_GOLDEN = [          # util/varint-test.cc:151-190 TestComparableEncoding (exact bit strings)
    (-37618632178637216379216387, "00000000 00000111 11100000 11100001 11110001 11011101 00000001 00110000 11011100 11111011 11101101 01011101 11111100"),
    (-9223372036854775809, "00000000 00111111 01111111 11111111 11111111 11111111 11111111 11111111 11111111 11111110"),
    (-(1 << 63), "00000000 00111111 01111111 11111111 11111111 11111111 11111111 11111111 11111111 11111111"),
    (-(1 << 31) - 1, "00000111 01111111 11111111 11111111 11111110"),
    (-(1 << 31), "00000111 01111111 11111111 11111111 11111111"),
    (-129, "00111111 01111110"), (-128, "00111111 01111111"), (-127, "00111111 10000000"),
    (-23, "01101000"), (0, "10000000"), (38, "10100110"),
    ((1 << 31) - 1, "11111000 01111111 11111111 11111111 11111111"),
    (1 << 31, "11111000 10000000 00000000 00000000 00000000"),
    ((1 << 63) - 1, "11111111 11000000 01111111 11111111 11111111 11111111 11111111 11111111 11111111 11111111"),
    (1 << 63, "11111111 11000000 10000000 00000000 00000000 00000000 00000000 00000000 00000000 00000000"),
    (8429024091289482183283928321, "11111111 11111100 00011011 00111100 01010011 01000111 10010101 11011111 00011011 00001010 01111011 11111101 01101101 00000001"),
]

# util/decimal-test.cc:27-53: test_cases with kComparableEncodingLengths; (digits, exponent, positive)
# in the canonical form 0.d1d2... * 10^exponent the Decimal class keeps.
DECIMAL_GOLDEN = [
    ([9, 8, 4, 7, 2, 3, 6, 7, 7, 6], 2147483658, False, 10), ([9, 8, 4, 7, 2, 3, 6, 7, 8], 2147483657, False, 10),
    ([1, 3, 4], 1, False, 3), ([1, 3, 3, 7], 1, False, 3), ([1, 3, 3, 4], 1, False, 3), ([1, 3, 3], 1, False, 3),
    ([1, 3, 6], -2147483644, False, 7), ([], 0, False, 1), ([5], -1, True, 2), ([1, 1, 5], 1, True, 3),
    ([1, 2], 1, True, 2), ([1, 2], 3, True, 2), ([1, 2], 101, True, 3), ([2, 6, 3, 8, 2], 3628, True, 6),
]


def test_varint_and_decimal_key_components(oracle):
    """kVarInt / kDecimal key entries (primitive_value.cc:1314-1349): the comparable encodings are
    self-delimiting only through their decoders (util/varint.cc:159-205, util/decimal.cc:339-367).
    The test-side encoders are pinned by the reference's exact bit strings / encoded lengths, then the
    oracle's component walk must find every entry's end, for both sort orders, at any position."""
    for v, bits in VARINT_GOLDEN:
        assert dk.varint_comparable(v) == bytes(int(b, 2) for b in bits.split()), v
    encs = [dk.varint_comparable(v) for v, _ in VARINT_GOLDEN]
    assert encs == sorted(encs)                                   # comparable: byte order = numeric order
    for digits, exp, pos, length in DECIMAL_GOLDEN:
        assert len(dk.decimal_comparable(digits, exp, pos)) == length, (digits, exp)
    dencs = [dk.decimal_comparable(d, e, p) for d, e, p, _ in DECIMAL_GOLDEN]
    assert dencs == sorted(dencs)
    comps = [dk.kvarint(v) for v, _ in VARINT_GOLDEN] + [dk.kvarint_desc(v) for v, _ in VARINT_GOLDEN]
    comps += [dk.kdecimal(d, e, p) for d, e, p, _ in DECIMAL_GOLDEN] + [dk.kdecimal_desc(d, e, p) for d, e, p, _ in DECIMAL_GOLDEN]
    rng = random.Random(7)
    for _ in range(300):
        comps.append(dk.kvarint(rng.randrange(-10**rng.randrange(1, 40), 10**rng.randrange(1, 40))))
        digits = [rng.randrange(1, 10)] + [rng.randrange(10) for _ in range(rng.randrange(0, 12))] + [rng.randrange(1, 10)]
        comps.append(dk.kdecimal(digits, rng.randrange(-10**6, 10**6), rng.random() < 0.5))
    for c in comps:
        d = dk.doc_key([c, "tail"], hash_code=3, hashed=[c])
        key = dk.sub_doc_key(d, [dk.kcol(1), c], micros=EPOCH + 5)
        ends = oracle.subdockey_ends(key)
        # [cotable-id end, DocKey end, subkey ends...] (SubDocKey::DecodeDocKeyAndSubKeyEnds, doc_key.cc:963-996)
        assert ends[:4] == [0, len(d), len(d) + 2, len(d) + 2 + len(c)], c


def _snappy_inputs():
    import os as _os
    import random as _random
    rng = _random.Random(5)

    def gen(n, mode):
        if mode == 0:
            return bytes(rng.randrange(256) for _ in range(n))
        if mode == 1:
            return bytes(n)
        if mode == 2:
            words = [bytes(rng.randrange(256) for _ in range(rng.randint(1, 40))) for _ in range(rng.randint(2, 30))]
            b = bytearray()
            while len(b) < n:
                b += rng.choice(words)
            return bytes(b[:n])
        if mode == 3:
            b = bytearray()
            while len(b) < n:
                b += bytes(rng.randrange(256) for _ in range(rng.randint(1, 300))) if rng.random() < 0.5 else bytes([rng.randrange(256)]) * rng.randint(1, 400)
            return bytes(b[:n])
        return bytes(rng.choice(b"ab" if mode == 4 else b"abcdefgh") for _ in range(n))       # many collisions, short matches
    for n in list(range(0, 140, 3)) + [255, 256, 257, 1000, 4096, 5000, 32768, 65535, 65536, 65537, 65540, 70000, 140000]:
        for mode in range(6):
            yield n, mode, gen(n, mode)


def test_snappy_format_round_trip_and_library_cross_check():
    """Snappy is the one third-party codec on the path: the library is not in the tree, so the FORMAT is
    what is pinned — every stream the repository's encoder emits decodes back, the real library (pyarrow bundles
    libsnappy) decodes it to the same bytes, and the oracle's decoder reads the library's own streams."""
    pa = None
    try:
        import pyarrow as pa_mod
        if pa_mod.Codec.is_available("snappy"):
            pa = pa_mod
    except ImportError:
        pass
    for n, mode, raw in _snappy_inputs():
        c = o.snappy_compress(raw)
        assert o.snappy_uncompress(c) == raw, (n, mode)
        if pa is not None and n:
            assert pa.decompress(c, decompressed_size=n, codec="snappy").to_pybytes() == raw, (n, mode)
            assert o.snappy_uncompress(pa.compress(raw, codec="snappy", asbytes=True)) == raw, (n, mode)
    for bad in (b"", b"\x05\x00a", b"\x04\x0cabcd\x01", b"\x08\x0cabcd\x05\x09"):      # empty, short literal, trailing junk, offset beyond the output
        with pytest.raises(ValueError):
            o.snappy_uncompress(bad)


def test_snappy_warp_model_is_the_scalar_encoder():
    """The GPU encoder tries 32 positions per step (snappy_warp_model.py mirrors the kernel statement for statement);
    its element stream must be the scalar encoder's, including the decision to give a block up."""
    from snappy_warp_model import warp_compress
    kept = 0
    for n, mode, raw in _snappy_inputs():
        c = o.snappy_compress(raw)
        want = c if len(c) < n - n // 8 else None
        assert warp_compress(raw) == want, (n, mode)
        kept += want is not None
    assert kept > 100
