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
