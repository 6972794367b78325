"""The per-entry device logic (dev_logic.cuh: key compare, row grouping, CompactionIterator rules,
feed_step) compiled for the CPU by tests/host_harness, against the oracle. Same inputs the GPU
parity tests use, so logic bugs surface without a GPU."""
import pytest

import dockv_util as dk
import harness_py as hh
import oracle_py as o
import workloads as w


def both(runs, **kw):
    p = o.CompactionParams(**kw)
    exp = o.compact_runs(runs, p).kv_list()
    got = hh.compact_runs(runs, p)
    return got, exp


@pytest.mark.parametrize("seed", range(12))
def test_randomized_docdb_runs(seed):
    runs = w.random_docdb_runs(seed, n_runs=1 + seed % 5, n_rows=25 + 5 * seed)
    for kw in w.param_grid():
        got, exp = both(runs, **kw)
        assert got == exp, kw


@pytest.mark.parametrize("seed", range(6))
def test_varint_and_decimal_key_components(seed):
    """kVarInt / kDecimal key entries in DocKeys and subkeys (primitive_value.cc:1314-1349): the device
    walk must delimit them exactly like the reference decoders, or rows and overwrite stacks break."""
    runs = w.random_numeric_key_runs(seed, n_runs=1 + seed % 4, n_rows=30 + 5 * seed)
    for kw in w.param_grid():
        got, exp = both(runs, **kw)
        assert got == exp, kw
    L = hh.lib()
    for k, _ in runs[0][:50]:
        u = k[:-8]
        assert L.hh_group_prefix_len(u, len(u), 1) == o.subdockey_ends(u)[1]


def test_plain_rocksdb_mode():
    seq = 0
    runs = []
    for i in range(2):
        c = []
        for k in range(2000):
            seq += 1
            c.append((o.ikey(str(i * 1000 + k).encode(), seq), str(i * 2000 + k).encode()))
        runs.append(w.sort_run(c))
    got, exp = both(runs, retention=False, bottommost=True, last_sequence=seq + 1)
    assert got == exp and len(got) == 3000
    got, exp = both(runs, retention=False, bottommost=False, last_sequence=seq + 1)
    assert got == exp


def test_key_bounds_and_obsolete_prefix():
    runs = w.random_docdb_runs(99, n_runs=3, n_rows=60)
    keys = sorted(k[:-8] for r in runs for k, _ in r)
    lo, up = keys[len(keys) // 4], keys[3 * len(keys) // 4]
    got, exp = both(runs, cutoff_ht=o.ht_from_micros(w.BASE_US + 50), lower=lo, upper=up)
    assert got == exp and 0 < len(got) < len(keys)


def test_group_prefix_len():
    L = hh.lib()
    d = dk.doc_key(["a", 5], hash_code=7, hashed=["h"])
    k = dk.sub_doc_key(d, [dk.kcol(2), "x"], micros=o.YB_EPOCH_US + 5)
    assert L.hh_group_prefix_len(k, len(k), 1) == len(d)
    assert L.hh_group_prefix_len(k, len(k), 0) == len(k)
    d2 = dk.doc_key(["only-range"])
    k2 = dk.sub_doc_key(d2, [], micros=o.YB_EPOCH_US + 5)
    assert L.hh_group_prefix_len(k2, len(k2), 1) == len(d2)
    cd = dk.doc_key(["r"], colocation=5)
    co = dk.sub_doc_key(cd, [dk.kcol(1)], micros=o.YB_EPOCH_US)
    assert L.hh_group_prefix_len(co, len(co), 1) == len(cd)
    tt = dk.table_tombstone_key(colocation=5, micros=o.YB_EPOCH_US)
    assert L.hh_group_prefix_len(tt, len(tt), 1) == 6         # id + '!'


@pytest.mark.parametrize("seed", range(10))
def test_cotables_and_colocated_tables(seed):
    runs = w.random_cotable_runs(seed, n_runs=1 + seed % 4, colocated=seed % 2 == 0)
    for kw in w.param_grid():
        got, exp = both(runs, **kw)
        assert got == exp, kw
    if seed % 2:        # cotables cutoff (master sys catalog): 'y' keys use their own history cutoff
        kw = dict(bottommost=True, cutoff_ht=o.ht_from_micros(w.BASE_US + 35), cotables_cutoff_ht=o.ht_from_micros(w.BASE_US + 85))
        got, exp = both(runs, **kw)
        assert got == exp


def test_bloom_filter_key_and_hash_match_the_oracle():
    """The device-side DocKeyV3 key transformer (both the stand-alone walk and the by-product of the
    row-group walk the merge kernel stores) and the LevelDB hash against the oracle restatement."""
    import ctypes as C
    import random
    import test_oracle_bloom as ob
    L = hh.lib()
    L.hh_filter_len_pair.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int)]
    L.hh_bloom_hash.argtypes = [C.c_char_p, C.c_uint32]
    L.hh_bloom_hash.restype = C.c_uint32
    keys = []
    for seed in range(6):
        for run in w.random_docdb_runs(seed, n_runs=2, n_rows=80):
            keys += [k[:-8] for k, _ in run]
        for run in w.random_cotable_runs(seed, n_runs=2, n_tables=4, rows_per_table=20, colocated=bool(seed % 2)):
            keys += [k[:-8] for k, _ in run]
    keys += [dk.sub_doc_key(dk.doc_key(["r1", dk.kint64(5)]), [dk.kcol(1)], micros=o.YB_EPOCH_US + 1),
             dk.sub_doc_key(dk.doc_key([], hash_code=7, hashed=["h"]), [], micros=o.YB_EPOCH_US + 1),
             dk.sub_doc_key(dk.doc_key([]), [dk.kcol(2)], micros=o.YB_EPOCH_US + 1)]
    assert len(keys) > 2000
    pair = (C.c_int * 2)()
    rng = random.Random(3)
    for k in keys:
        g = L.hh_filter_len_pair(k, len(k), 1, pair)
        want = len(ob.filter_key(k))
        assert pair[1] == want
        if g >= 0:
            assert pair[0] == want
        L.hh_filter_len_pair(k, len(k), 0, pair)             # plain mode: stand-alone walk
        assert pair[0] == want
        fk = k[:want] + bytes(rng.randrange(256) for _ in range(rng.randrange(4)))
        assert L.hh_bloom_hash(fk, len(fk)) == ob.bloom_hash(fk)


@pytest.mark.parametrize("cut", [1, 2, 3, 7])
def test_state_rebuilt_by_ancestor_replay_inside_groups(cut):
    """Merge tiles may start inside a row group (groups larger than a tile): the tile rebuilds Feed's state by
    replaying the ancestors `P_i # HT` and the earlier versions of its first key (dev_logic.cuh replay_ancestors).
    Here the state is thrown away and rebuilt before every `cut`-th record of every group — cut = 1: before EVERY
    record — and the result must still equal the oracle's single sequential pass."""
    L = hh.lib()
    L.hh_set_cut_every(cut)
    try:
        for seed in range(6):
            runs = w.random_docdb_runs(200 + seed, n_runs=1 + seed % 4, n_rows=20 + 5 * seed)
            for kw in w.param_grid():
                got, exp = both(runs, **kw)
                assert got == exp, (seed, kw)
        for seed in range(4):
            runs = w.random_cotable_runs(300 + seed, n_runs=1 + seed % 3, n_tables=3, rows_per_table=10, colocated=seed % 2 == 0)
            for kw in w.param_grid()[:8]:
                got, exp = both(runs, **kw)
                assert got == exp, (seed, kw)
        for seed in range(3):
            runs = w.random_numeric_key_runs(400 + seed, n_runs=2, n_rows=25)
            for kw in w.param_grid()[:6]:
                got, exp = both(runs, **kw)
                assert got == exp, (seed, kw)
        import test_reference_dumps as t          # TTL / merge-record chains, collections, user timestamps
        for name in sorted(dir(t)):
            if name.startswith("test_") and "parser" not in name:
                getattr(t, name)()
    finally:
        L.hh_set_cut_every(0)


def test_entry_header_parser_all_paths():
    """parse_entry_header has three paths: three one-byte varints, fields of at most two bytes decoded in straight-line
    code (every real key, most values), and the general byte loop. All three against an independent LEB128 encoder, with
    short `avail` and truncated headers."""
    import ctypes as C
    import random
    L = hh.lib()
    L.hh_parse_entry_header.argtypes = [C.c_char_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]

    def leb(v):
        out = bytearray()
        while v >= 128:
            out.append((v & 127) | 128)
            v >>= 7
        out.append(v)
        return bytes(out)
    rng = random.Random(3)
    pools = [[0, 1, 5, 127], [128, 129, 300, 16383], [16384, 70000, (1 << 28) - 1, (1 << 32) - 1]]
    cases = [(a, b, c) for pa in pools for pb in pools for pc in pools for a in pa[:2] for b in pb[-2:] for c in (pc[0], pc[-1])]
    cases += [tuple(rng.choice(rng.choice(pools)) for _ in range(3)) for _ in range(400)]
    oa, ob, oc = C.c_uint32(), C.c_uint32(), C.c_uint32()

    def parse(buf, avail):
        return L.hh_parse_entry_header(buf, avail, C.byref(oa), C.byref(ob), C.byref(oc))
    for a, b, c in cases:
        hdr = leb(a) + leb(b) + leb(c)
        for pad in (0, 1, 9):
            buf = hdr + bytes(rng.randrange(256) for _ in range(pad))
            n = parse(buf + b"\0" * 8, len(buf))
            if len(buf) < 3:
                assert n == 0                      # DecodeEntry needs three bytes (block.cc:70-73)
                continue
            assert n == len(hdr) and (oa.value, ob.value, oc.value) == (a, b, c), (a, b, c, pad)
        if len(hdr) > 3:
            assert parse(hdr + b"\0" * 8, len(hdr) - 1) == 0      # truncated inside a varint
    assert parse(bytes([0x80] * 6 + [1]) + b"\0" * 8, 7) == 0      # a varint32 longer than 5 bytes


def test_hybrid_time_filter_predicate():
    """hidden_by_ht_filters (the device side of HybridTimeFilteringIterator::Satisfied, docdb_rocksdb_util.cc:525-565)
    against the rule written out in Python: global filter, per-database cotable filters (database oid = uuid bytes
    12..15, little endian), keys of other tables, undecodable hybrid times."""
    import ctypes as C
    import struct
    import numpy as np
    L = hh.lib()
    L.hh_hidden_by_ht_filters.argtypes = [C.c_char_p, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32]
    none = (1 << 64) - 2
    oids = np.array([7, 1000, 70000, 0x01020304], np.uint32)
    hts = np.array([o.ht_from_micros(o.YB_EPOCH_US + t) for t in (50, 20, 90, 10)], np.uint64)

    def key(prefix, micros, logical=0):
        return prefix + b"Srow\x00\x00!" + b"K\x81" + b"#" + o.encode_doc_ht(o.YB_EPOCH_US + micros, logical)
    for oid in (7, 8, 1000, 70000, 0x01020304, 0xffffffff):
        uuid = bytes(range(1, 13)) + struct.pack("<I", oid)
        for prefix in (b"y" + uuid, b"0" + struct.pack(">I", oid), b""):
            for micros in (5, 10, 11, 20, 21, 50, 51, 90, 91):
                for glob in (none, o.ht_from_micros(o.YB_EPOCH_US + 30)):
                    k = key(prefix, micros)
                    ht = o.ht_from_micros(o.YB_EPOCH_US + micros)
                    want = glob != none and ht > glob
                    if not want and prefix[:1] == b"y" and oid in oids.tolist():
                        want = ht > int(hts[oids.tolist().index(oid)])
                    got = L.hh_hidden_by_ht_filters(k, len(k), glob, oids.ctypes.data, hts.ctypes.data, len(oids))
                    assert bool(got) == want, (oid, prefix[:1], micros, glob)
                    assert not L.hh_hidden_by_ht_filters(k, len(k), none, None, None, 0)
    bad = b"y" + bytes(16) + b"Sx\x00\x00!" + b"#\x00"          # a DocHybridTime that does not decode: visible
    assert not L.hh_hidden_by_ht_filters(bad, len(bad), 0, oids.ctypes.data, hts.ctypes.data, len(oids))
