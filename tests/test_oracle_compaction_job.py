"""RocksDB-level golden tests (rocksdb/db/compaction_job_test.cc) and SST round trips."""
import random

import oracle_py as o


def plain(runs, **kw):
    return o.compact_runs(runs, o.CompactionParams(retention=False, **kw)).kv_list()


def test_simple_two_files_10k():
    # compaction_job_test.cc:193-231 CreateTwoFiles + :339-347 Simple
    expected = {}
    seq = 0
    runs = []
    for i in range(2):
        contents = {}
        for k in range(10000):
            key = str(i * 5000 + k).encode()
            value = str(i * 10000 + k).encode()
            seq += 1
            contents[o.ikey(key, seq)] = value
            if i == 1 or k < 5000:
                expected[o.ikey(key, seq if key == b"9999" else 0)] = value
        runs.append(sorted(contents.items(), key=lambda kv: (kv[0][:-8], -int.from_bytes(kv[0][-8:], "little"))))
    got = plain(runs, bottommost=True, last_sequence=seq + 1)
    exp = sorted(expected.items(), key=lambda kv: (kv[0][:-8], -int.from_bytes(kv[0][-8:], "little")))
    assert len(got) == 15000
    assert got == exp


def test_simple_overwrite():
    # compaction_job_test.cc:470-490
    f1 = [(o.ikey(b"a", 3), b"val2"), (o.ikey(b"b", 4), b"val3")]
    f2 = [(o.ikey(b"a", 1), b"val"), (o.ikey(b"b", 2), b"val")]
    got = plain([f1, f2], bottommost=True, last_sequence=5)
    assert got == [(o.ikey(b"a", 0), b"val2"), (o.ikey(b"b", 4), b"val3")]


def test_simple_deletion_bottommost():
    # compaction_job_test.cc SimpleDeletion: deletion markers vanish at the bottommost level
    f1 = [(o.ikey(b"c", 4, 0), b""), (o.ikey(b"c", 3), b"val")]
    f2 = [(o.ikey(b"b", 2), b"val"), (o.ikey(b"b", 1), b"val")]
    got = plain([f1, f2], bottommost=True, last_sequence=5)
    assert got == [(o.ikey(b"b", 0), b"val")]


def test_simple_non_last_level_keeps_seqnos():
    f1 = [(o.ikey(b"a", 5), b"val2"), (o.ikey(b"b", 6), b"val3")]
    f2 = [(o.ikey(b"a", 3), b"val"), (o.ikey(b"b", 4), b"val")]
    got = plain([f1, f2], bottommost=False, last_sequence=7)
    assert got == [(o.ikey(b"a", 5), b"val2"), (o.ikey(b"b", 6), b"val3")]


def _rand_kvs(rng, n, klen=(1, 60), vlen=(0, 300)):
    keys = set()
    while len(keys) < n:
        keys.add(bytes(rng.randrange(256) for _ in range(rng.randrange(*klen))))
    kvs = []
    for i, k in enumerate(sorted(keys)):
        kvs.append((o.ikey(k, 1000 + i), bytes(rng.randrange(256) for _ in range(rng.randrange(*vlen)))))
    return kvs


def test_sst_round_trip_both_encodings():
    # table/block_test.cc:536-855 style: encode -> decode must reproduce every KV
    rng = random.Random(11)
    for enc in (1, 2):
        for n, bs in ((1, 4096), (17, 256), (500, 1024), (3000, 4096)):
            kvs = _rand_kvs(rng, n)
            sst = o.Sst.build(kvs, o.TableOptions(block_size=bs, key_encoding=enc))
            assert sst.key_encoding == enc
            assert sst.read_all() == kvs
            assert sst.num_entries == n
            off, sz = sst.block_handles()
            assert len(off) >= 1 and int(off[-1] + sz[-1]) + 5 == len(sst.data)


def test_sst_round_trip_docdb_shaped_keys_three_shared_parts():
    # Keys shaped like DocDB's (same prefix, seq+1 suffixes) exercise the frequent-case encodings
    rng = random.Random(5)
    cfg = o.GenConfig(seed=9, num_rows=300, cols=3, versions=4, num_files=1, value_len=40)
    for enc in (1, 2):
        sst = o.Sst.generate(cfg, 0, o.TableOptions(block_size=2048, key_encoding=enc))
        kvs = sst.read_all()
        assert len(kvs) == 300 * 3 * 4
        keys = [k for k, _ in kvs]
        assert keys == sorted(keys, key=lambda k: (k[:-8], -int.from_bytes(k[-8:], "little")))
        sst2 = o.Sst.build(kvs, o.TableOptions(block_size=2048, key_encoding=enc))
        assert sst2.data == sst.data and sst2.meta == sst.meta


def test_multi_level_index_is_walkable():
    rng = random.Random(2)
    kvs = _rand_kvs(rng, 6000, klen=(8, 24), vlen=(0, 8))
    sst = o.Sst.build(kvs, o.TableOptions(block_size=128, index_block_size=256, min_keys_per_index_block=4))
    assert sst.read_all() == kvs


def test_compact_through_ssts_matches_runs():
    cfg = o.GenConfig(seed=3, num_rows=500, cols=2, versions=5, num_files=4, value_len=64, tombstone_per_1024=100)
    ssts = [o.Sst.generate(cfg, f, o.TableOptions(block_size=1024)) for f in range(4)]
    runs = [s.read_all() for s in ssts]
    cutoff = o.ht_from_micros(cfg.base_micros + 2500)
    p = o.CompactionParams(cutoff_ht=cutoff)
    a = o.compact(ssts, p, o.TableOptions(block_size=1024))
    b = o.compact_runs(runs, p)
    assert a.kv_list() == b.kv_list()
    assert a.stats.num_input_records == 500 * 2 * 5
    assert a.stats.kv_hash == b.stats.kv_hash
    out = a.sst()
    assert out.read_all() == a.kv_list()
