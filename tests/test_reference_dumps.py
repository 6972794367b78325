"""Replays the reference's history-compaction golden tests: the debug dumps that
src/yb/docdb/docdb-test-wrapper.cc and docdb-ttl-test.cc assert before and after
FullyCompactHistoryBefore(cutoff) are parsed (tests/refdump.py) and fed to the oracle and to the
device logic compiled for the CPU (tests/host_harness); both must produce exactly the reference's
expected surviving set, whatever the split of the entries over input files.

FullyCompactHistoryBefore = a major compaction of all files (no other data: other_min = max) with
the given history cutoff (docdb/docdb_test_util.cc)."""
import pytest

import harness_py as hh
import oracle_py as o
import refdump as rd
import workloads as w

SEQ0 = 1 << 50
EXTRA_CHECK = None      # set by the GPU suite: called as EXTRA_CHECK(runs, params_kwargs, want) for every replayed compaction


def us(n):
    return o.ht_from_micros(n)


def split_runs(entries, n_runs, salt):
    runs = [[] for _ in range(n_runs)]
    for i, (k, v) in enumerate(entries):
        r = (i * 7 + salt * 3 + (i * i) % 5) % n_runs
        runs[r].append((o.ikey(k, SEQ0 + i), v))
    return [w.sort_run(r) for r in runs if r]


def fully_compact(state, cutoff_us, expected, **kw):
    """state / expected: dump text or parsed [(user_key, value)]. Returns the new state."""
    entries = rd.parse(state) if isinstance(state, str) else state
    want = rd.parse(expected) if isinstance(expected, str) else expected
    got = None
    for n_runs, salt in ((1, 0), (2, 1), (3, 2), (4, 5)):
        runs = split_runs(entries, n_runs, salt)
        cut = kw.get("cutoff_ht", us(cutoff_us) if cutoff_us is not None else o.HT_MIN)
        p = o.CompactionParams(bottommost=True, other_min_ht=o.HT_MAX, **dict(kw, cutoff_ht=cut))
        got = [(k[:-8], v) for k, v in o.compact_runs(runs, p).kv_list()] if runs else []
        assert got == want, "oracle, %d runs" % n_runs
        dev = [(k[:-8], v) for k, v in hh.compact_runs(runs, p)] if runs else []
        assert dev == want, "device logic, %d runs" % n_runs
        if EXTRA_CHECK is not None and n_runs == 3 and runs:
            EXTRA_CHECK(runs, dict(kw, cutoff_ht=cut), want)
    return got


# docdb/docdb-test.h:111-126 kPredefinedDBStateDebugDumpStr
PREDEFINED = r'''
SubDocKey(DocKey([], ["my_key_where_value_is_a_string"]), [HT{ physical: 1000 }]) -> "value1"
SubDocKey(DocKey([], ["mydockey", 123456]), [HT{ physical: 2000 }]) -> {}
SubDocKey(DocKey([], ["mydockey", 123456]), ["subkey_a"; HT{ physical: 2000 w: 1 }]) -> "value_a"
SubDocKey(DocKey([], ["mydockey", 123456]), ["subkey_b"; HT{ physical: 7000 }]) -> {}
SubDocKey(DocKey([], ["mydockey", 123456]), ["subkey_b"; HT{ physical: 6000 }]) -> DEL
SubDocKey(DocKey([], ["mydockey", 123456]), ["subkey_b"; HT{ physical: 3000 }]) -> {}
SubDocKey(DocKey([], ["mydockey", 123456]), ["subkey_b", "subkey_c"; HT{ physical: 7000 w: 1 }]) \
    -> "value_bc_prime"
SubDocKey(DocKey([], ["mydockey", 123456]), ["subkey_b", "subkey_c"; HT{ physical: 5000 }]) -> DEL
SubDocKey(DocKey([], ["mydockey", 123456]), ["subkey_b", "subkey_c"; HT{ physical: 3000 w: 1 }]) \
    -> "value_bc"
SubDocKey(DocKey([], ["mydockey", 123456]), ["subkey_b", "subkey_d"; HT{ physical: 3500 }]) -> \
    "value_bd"
'''


def test_parser_reproduces_the_reference_key_bytes():
    # docdb-test-wrapper.cc:877-888: PutCF('Smydockey\x00\x00I\x80\x00\x00\x00\x00\x01\xe2@!Ssubkey_a\x00\x00', 'Svalue_a')
    entries = rd.parse(PREDEFINED)
    key, value = entries[2]
    assert key.startswith(b"Smydockey\x00\x00I\x80\x00\x00\x00\x00\x01\xe2@!Ssubkey_a\x00\x00#")
    assert value == b"Svalue_a" and entries[1][1] == b"{" and entries[4][1] == b"X"
    p = o.CompactionParams(bottommost=True, cutoff_ht=o.HT_MIN, other_min_ht=o.HT_MAX)
    kept = [(k[:-8], v) for k, v in o.compact_runs(split_runs(entries, 3, 1), p).kv_list()]
    assert kept == entries                                  # merge order of the oracle = dump order of the reference


def test_basic_test_history_cleanup():
    """docdb-test-wrapper.cc:970-1088 (BasicTest, "Compaction cleanup testing")."""
    after_5000 = r'''
SubDocKey(DocKey([], ["my_key_where_value_is_a_string"]), [HT{ physical: 1000 }]) -> "value1"
SubDocKey(DocKey([], ["mydockey", 123456]), [HT{ physical: 2000 }]) -> {}
SubDocKey(DocKey([], ["mydockey", 123456]), ["subkey_a"; HT{ physical: 2000 w: 1 }]) -> "value_a"
SubDocKey(DocKey([], ["mydockey", 123456]), ["subkey_b"; HT{ physical: 7000 }]) -> {}
SubDocKey(DocKey([], ["mydockey", 123456]), ["subkey_b"; HT{ physical: 6000 }]) -> DEL
SubDocKey(DocKey([], ["mydockey", 123456]), ["subkey_b"; HT{ physical: 3000 }]) -> {}
SubDocKey(DocKey([], ["mydockey", 123456]), ["subkey_b", "subkey_c"; HT{ physical: 7000 w: 1 }]) \
    -> "value_bc_prime"
SubDocKey(DocKey([], ["mydockey", 123456]), ["subkey_b", "subkey_d"; HT{ physical: 3500 }]) -> \
    "value_bd"
'''
    after_6000 = r'''
SubDocKey(DocKey([], ["my_key_where_value_is_a_string"]), [HT{ physical: 1000 }]) -> "value1"
SubDocKey(DocKey([], ["mydockey", 123456]), [HT{ physical: 2000 }]) -> {}
SubDocKey(DocKey([], ["mydockey", 123456]), ["subkey_a"; HT{ physical: 2000 w: 1 }]) -> "value_a"
SubDocKey(DocKey([], ["mydockey", 123456]), ["subkey_b"; HT{ physical: 7000 }]) -> {}
SubDocKey(DocKey([], ["mydockey", 123456]), ["subkey_b", "subkey_c"; HT{ physical: 7000 w: 1 }]) \
    -> "value_bc_prime"
'''
    s0 = rd.parse(PREDEFINED)
    s1 = fully_compact(s0, 5000, after_5000)
    # "starting both from the initial state as well as from the state with the first history compaction"
    for snap in (s0, s1):
        s2 = fully_compact(snap, 6000, after_6000)
    # the document is overwritten with an empty object at 8000 on top of every snapshot (:1026-1046)
    overwrite = rd.parse('SubDocKey(DocKey([], ["mydockey", 123456]), [HT{ physical: 8000 }]) -> {}')

    def plus_overwrite(state):
        return sorted(state + overwrite, key=lambda kv: next(i for i, k in enumerate(order) if k == kv[0]))
    with_8000 = r'''
SubDocKey(DocKey([], ["my_key_where_value_is_a_string"]), [HT{ physical: 1000 }]) -> "value1"
SubDocKey(DocKey([], ["mydockey", 123456]), [HT{ physical: 8000 }]) -> {}
SubDocKey(DocKey([], ["mydockey", 123456]), [HT{ physical: 2000 }]) -> {}
SubDocKey(DocKey([], ["mydockey", 123456]), ["subkey_a"; HT{ physical: 2000 w: 1 }]) -> "value_a"
SubDocKey(DocKey([], ["mydockey", 123456]), ["subkey_b"; HT{ physical: 7000 }]) -> {}
SubDocKey(DocKey([], ["mydockey", 123456]), ["subkey_b"; HT{ physical: 6000 }]) -> DEL
SubDocKey(DocKey([], ["mydockey", 123456]), ["subkey_b"; HT{ physical: 3000 }]) -> {}
SubDocKey(DocKey([], ["mydockey", 123456]), ["subkey_b", "subkey_c"; HT{ physical: 7000 w: 1 }]) \
    -> "value_bc_prime"
SubDocKey(DocKey([], ["mydockey", 123456]), ["subkey_b", "subkey_c"; HT{ physical: 5000 }]) -> DEL
SubDocKey(DocKey([], ["mydockey", 123456]), ["subkey_b", "subkey_c"; HT{ physical: 3000 w: 1 }]) \
    -> "value_bc"
SubDocKey(DocKey([], ["mydockey", 123456]), ["subkey_b", "subkey_d"; HT{ physical: 3500 }]) -> \
    "value_bd"
'''
    order = [k for k, _ in rd.parse(with_8000)]
    after_7999 = r'''
SubDocKey(DocKey([], ["my_key_where_value_is_a_string"]), [HT{ physical: 1000 }]) -> "value1"
SubDocKey(DocKey([], ["mydockey", 123456]), [HT{ physical: 8000 }]) -> {}
SubDocKey(DocKey([], ["mydockey", 123456]), [HT{ physical: 2000 }]) -> {}
SubDocKey(DocKey([], ["mydockey", 123456]), ["subkey_a"; HT{ physical: 2000 w: 1 }]) -> "value_a"
SubDocKey(DocKey([], ["mydockey", 123456]), ["subkey_b"; HT{ physical: 7000 }]) -> {}
SubDocKey(DocKey([], ["mydockey", 123456]), ["subkey_b", "subkey_c"; HT{ physical: 7000 w: 1 }]) \
    -> "value_bc_prime"
'''
    after_8000 = r'''
SubDocKey(DocKey([], ["my_key_where_value_is_a_string"]), [HT{ physical: 1000 }]) -> "value1"
SubDocKey(DocKey([], ["mydockey", 123456]), [HT{ physical: 8000 }]) -> {}
'''
    full = plus_overwrite(s0)
    assert full == rd.parse(with_8000)
    s3 = fully_compact(full, 7999, after_7999)
    # "Starting with each snapshot, perform the final history compaction and verify we always get the same result."
    for snap in (full, s3, plus_overwrite(s1), plus_overwrite(s2)):
        fully_compact(snap, 8000, after_8000)


def test_static_column_compaction():
    """docdb-test-wrapper.cc:1404-1515 (StaticColumnCompaction): TTL expiry, overwritten versions and a
    tombstone at the cutoff, static columns (DocKey with hashed components only)."""
    before = r'''
SubDocKey(DocKey(0x0000, ["h1"], []), ["s1"; HT{ physical: 1000 }]) -> "v1"; ttl: 0.001s
SubDocKey(DocKey(0x0000, ["h1"], []), ["s2"; HT{ physical: 1000 }]) -> "v2"; ttl: 0.002s
SubDocKey(DocKey(0x0000, ["h1"], []), ["s3"; HT{ physical: 3000 }]) -> "v3new"
SubDocKey(DocKey(0x0000, ["h1"], []), ["s3"; HT{ physical: 1000 }]) -> "v3old"
SubDocKey(DocKey(0x0000, ["h1"], []), ["s4"; HT{ physical: 3000 }]) -> DEL
SubDocKey(DocKey(0x0000, ["h1"], []), ["s4"; HT{ physical: 1000 }]) -> "v4"
SubDocKey(DocKey(0x0000, ["h1"], ["r1"]), ["c5"; HT{ physical: 1000 }]) -> "v51"; ttl: 0.001s
SubDocKey(DocKey(0x0000, ["h1"], ["r1"]), ["c6"; HT{ physical: 1000 }]) -> "v61"; ttl: 0.002s
SubDocKey(DocKey(0x0000, ["h1"], ["r1"]), ["c7"; HT{ physical: 3000 }]) -> "v71new"
SubDocKey(DocKey(0x0000, ["h1"], ["r1"]), ["c7"; HT{ physical: 1000 }]) -> "v71old"
SubDocKey(DocKey(0x0000, ["h1"], ["r1"]), ["c8"; HT{ physical: 1000 }]) -> "v81"
SubDocKey(DocKey(0x0000, ["h1"], ["r2"]), ["c5"; HT{ physical: 1000 }]) -> "v52"; ttl: 0.001s
SubDocKey(DocKey(0x0000, ["h1"], ["r2"]), ["c6"; HT{ physical: 1000 }]) -> "v62"; ttl: 0.002s
SubDocKey(DocKey(0x0000, ["h1"], ["r2"]), ["c7"; HT{ physical: 1000 }]) -> "v72"
SubDocKey(DocKey(0x0000, ["h1"], ["r2"]), ["c8"; HT{ physical: 5000 }]) -> DEL
SubDocKey(DocKey(0x0000, ["h1"], ["r2"]), ["c8"; HT{ physical: 1000 }]) -> "v82"
'''
    after = r'''
SubDocKey(DocKey(0x0000, ["h1"], []), ["s2"; HT{ physical: 1000 }]) -> "v2"; ttl: 0.002s
SubDocKey(DocKey(0x0000, ["h1"], []), ["s3"; HT{ physical: 3000 }]) -> "v3new"
SubDocKey(DocKey(0x0000, ["h1"], ["r1"]), ["c6"; HT{ physical: 1000 }]) -> "v61"; ttl: 0.002s
SubDocKey(DocKey(0x0000, ["h1"], ["r1"]), ["c7"; HT{ physical: 3000 }]) -> "v71new"
SubDocKey(DocKey(0x0000, ["h1"], ["r1"]), ["c8"; HT{ physical: 1000 }]) -> "v81"
SubDocKey(DocKey(0x0000, ["h1"], ["r2"]), ["c6"; HT{ physical: 1000 }]) -> "v62"; ttl: 0.002s
SubDocKey(DocKey(0x0000, ["h1"], ["r2"]), ["c7"; HT{ physical: 1000 }]) -> "v72"
SubDocKey(DocKey(0x0000, ["h1"], ["r2"]), ["c8"; HT{ physical: 5000 }]) -> DEL
SubDocKey(DocKey(0x0000, ["h1"], ["r2"]), ["c8"; HT{ physical: 1000 }]) -> "v82"
'''
    fully_compact(before, 3000, after)


def test_compaction_with_user_timestamp():
    """docdb-test-wrapper.cc:1580-1655 (TestCompactionWithUserTimestamp)."""
    fully_compact(r'''
      SubDocKey(DocKey([], ["k1"]), ["s1"; HT{ physical: 5000 }]) -> DEL
      SubDocKey(DocKey([], ["k1"]), ["s1"; HT{ physical: 3000 }]) -> "v11"
      ''', 5000, "")                                        # "Compaction takes away everything."
    fully_compact(r'''
      SubDocKey(DocKey([], ["k1"]), ["s1"; HT{ physical: 3000 }]) -> "v13"; timestamp: 4000
      SubDocKey(DocKey([], ["k1"]), ["s2"; HT{ physical: 3000 }]) -> "v11"; ttl: 0.001s
      ''', 5000, r'''
      SubDocKey(DocKey([], ["k1"]), ["s1"; HT{ physical: 3000 }]) -> "v13"; timestamp: 4000
      ''')


def test_compaction_with_transactions_regular_db():
    """docdb-test-wrapper.cc:1657-1810 (CompactionWithTransactions): the regular-DB records; the intents
    the test also dumps live in the intents DB, which a regular-DB compaction does not touch."""
    fully_compact(r'''
SubDocKey(DocKey([], ["mydockey", 123456]), [HT{ physical: 4000 }]) -> {}
SubDocKey(DocKey([], ["mydockey", 123456]), [HT{ physical: 1000 }]) -> {}
SubDocKey(DocKey([], ["mydockey", 123456]), ["subkey1"; HT{ physical: 3000 }]) -> "value3"
SubDocKey(DocKey([], ["mydockey", 123456]), ["subkey1"; HT{ physical: 2000 }]) -> "value2"
SubDocKey(DocKey([], ["mydockey", 123456]), ["subkey1"; HT{ physical: 1000 }]) -> "value1"
''', 3500, r'''
SubDocKey(DocKey([], ["mydockey", 123456]), [HT{ physical: 4000 }]) -> {}
SubDocKey(DocKey([], ["mydockey", 123456]), [HT{ physical: 1000 }]) -> {}
SubDocKey(DocKey([], ["mydockey", 123456]), ["subkey1"; HT{ physical: 3000 }]) -> "value3"
''')


def test_expired_value_compaction():
    """docdb-ttl-test.cc:25-70 (ExpiredValueCompactionTest)."""
    fully_compact(r'''
      SubDocKey(DocKey([], ["k1"]), ["s1"; HT{ physical: 5000 }]) -> "v14"
      SubDocKey(DocKey([], ["k1"]), ["s1"; HT{ physical: 1000 }]) -> "v11"; ttl: 0.001s
      SubDocKey(DocKey([], ["k1"]), ["s2"; HT{ physical: 5000 }]) -> "v24"
      SubDocKey(DocKey([], ["k1"]), ["s2"; HT{ physical: 1000 }]) -> "v21"; ttl: 0.003s
      ''', 3000, r'''
SubDocKey(DocKey([], ["k1"]), ["s1"; HT{ physical: 5000 }]) -> "v14"
SubDocKey(DocKey([], ["k1"]), ["s2"; HT{ physical: 5000 }]) -> "v24"
SubDocKey(DocKey([], ["k1"]), ["s2"; HT{ physical: 1000 }]) -> "v21"; ttl: 0.003s
''')


def test_ttl_compaction_sequence():
    """docdb-ttl-test.cc:883-991 (TTLCompactionTest): liveness columns, column TTLs expiring one by one,
    then tombstones at and above the cutoff."""
    s = rd.parse(r'''
SubDocKey(DocKey([], ["k1"]), [SystemColumnId(0); HT{ physical: 1000 }]) -> null; ttl: 0.001s
SubDocKey(DocKey([], ["k1"]), [ColumnId(0); HT{ physical: 1000 }]) -> "v1"; ttl: 0.002s
SubDocKey(DocKey([], ["k1"]), [ColumnId(1); HT{ physical: 1000 }]) -> "v2"; ttl: 0.003s
SubDocKey(DocKey([], ["k1"]), [ColumnId(2); HT{ physical: 1000 }]) -> "v3"
SubDocKey(DocKey([], ["k1"]), [ColumnId(3); HT{ physical: 1000 }]) -> "v4"
SubDocKey(DocKey([], ["k2"]), [SystemColumnId(0); HT{ physical: 1000 }]) -> null; ttl: 0.003s
SubDocKey(DocKey([], ["k2"]), [ColumnId(0); HT{ physical: 1000 }]) -> "v1"; ttl: 0.002s
SubDocKey(DocKey([], ["k2"]), [ColumnId(1); HT{ physical: 1000 }]) -> "v2"; ttl: 0.001s
''')
    s = fully_compact(s, 3000, r'''
SubDocKey(DocKey([], ["k1"]), [ColumnId(0); HT{ physical: 1000 }]) -> "v1"; ttl: 0.002s
SubDocKey(DocKey([], ["k1"]), [ColumnId(1); HT{ physical: 1000 }]) -> "v2"; ttl: 0.003s
SubDocKey(DocKey([], ["k1"]), [ColumnId(2); HT{ physical: 1000 }]) -> "v3"
SubDocKey(DocKey([], ["k1"]), [ColumnId(3); HT{ physical: 1000 }]) -> "v4"
SubDocKey(DocKey([], ["k2"]), [SystemColumnId(0); HT{ physical: 1000 }]) -> null; ttl: 0.003s
SubDocKey(DocKey([], ["k2"]), [ColumnId(0); HT{ physical: 1000 }]) -> "v1"; ttl: 0.002s
''')
    s = fully_compact(s, 4000, r'''
SubDocKey(DocKey([], ["k1"]), [ColumnId(1); HT{ physical: 1000 }]) -> "v2"; ttl: 0.003s
SubDocKey(DocKey([], ["k1"]), [ColumnId(2); HT{ physical: 1000 }]) -> "v3"
SubDocKey(DocKey([], ["k1"]), [ColumnId(3); HT{ physical: 1000 }]) -> "v4"
SubDocKey(DocKey([], ["k2"]), [SystemColumnId(0); HT{ physical: 1000 }]) -> null; ttl: 0.003s
''')
    s = fully_compact(s, 5000, r'''
SubDocKey(DocKey([], ["k1"]), [ColumnId(2); HT{ physical: 1000 }]) -> "v3"
SubDocKey(DocKey([], ["k1"]), [ColumnId(3); HT{ physical: 1000 }]) -> "v4"
''')
    tombstoned = r'''
SubDocKey(DocKey([], ["k1"]), [ColumnId(2); HT{ physical: 2000 }]) -> DEL
SubDocKey(DocKey([], ["k1"]), [ColumnId(2); HT{ physical: 1000 }]) -> "v3"
SubDocKey(DocKey([], ["k1"]), [ColumnId(3); HT{ physical: 2000 }]) -> DEL
SubDocKey(DocKey([], ["k1"]), [ColumnId(3); HT{ physical: 1000 }]) -> "v4"
'''
    s = fully_compact(tombstoned, 1000, tombstoned)          # "Nothing is removed."
    fully_compact(s, 2000, "")                              # "Next compactions removes everything."


def test_table_ttl_compaction_sequence():
    """docdb-ttl-test.cc:993-1055 (TableTTLCompactionTest): table-level TTL of 2 ms next to column TTLs;
    a TTL of 0 means "never expires"."""
    kw = dict(table_ttl_ns=2 * 10**6)
    s = fully_compact(r'''
      SubDocKey(DocKey([], ["k1"]), ["s1"; HT{ physical: 1000 }]) -> "v1"; ttl: 0.001s
      SubDocKey(DocKey([], ["k1"]), ["s2"; HT{ physical: 1000 }]) -> "v2"
      SubDocKey(DocKey([], ["k1"]), ["s3"; HT{ physical: 2000 }]) -> "v3"; ttl: 0.000s
      SubDocKey(DocKey([], ["k1"]), ["s4"; HT{ physical: 1000 }]) -> "v4"; ttl: 0.003s
      ''', 3000, r'''
SubDocKey(DocKey([], ["k1"]), ["s2"; HT{ physical: 1000 }]) -> "v2"
SubDocKey(DocKey([], ["k1"]), ["s3"; HT{ physical: 2000 }]) -> "v3"; ttl: 0.000s
SubDocKey(DocKey([], ["k1"]), ["s4"; HT{ physical: 1000 }]) -> "v4"; ttl: 0.003s
''', **kw)
    s = fully_compact(s, 4000, r'''
SubDocKey(DocKey([], ["k1"]), ["s3"; HT{ physical: 2000 }]) -> "v3"; ttl: 0.000s
SubDocKey(DocKey([], ["k1"]), ["s4"; HT{ physical: 1000 }]) -> "v4"; ttl: 0.003s
''', **kw)
    fully_compact(s, 5000, r'''
SubDocKey(DocKey([], ["k1"]), ["s3"; HT{ physical: 2000 }]) -> "v3"; ttl: 0.000s
''', **kw)


COTABLE1 = "0000400200003000800000000000400a"
COTABLE2 = "0000400500003000800000000000400a"


def _versions(dockey, upto=4):
    return "\n".join('SubDocKey(DocKey(%s), ["subkey1"; HT{ physical: %d }]) -> "value%d"' % (dockey, 1000 * i, i)
                     for i in range(1, upto + 1))


def _sorted(entries):
    return [kv for run in [w.sort_run([(o.ikey(k, SEQ0), v) for k, v in entries])] for kv in [(k[:-8], v) for k, v in run]]


def test_history_retention_with_cotables():
    """docdb-test-wrapper.cc:2125-2242 (HistoryRetentionWithCotables): HistoryCutoff{cotables 3000,
    primary 2000} — cotable rows keep 3000 and 4000, the id-less table keeps 2000, 3000 and 4000."""
    c1 = 'CoTableId=%s, 0x0001, ["cotablekey", 10000], []' % COTABLE1
    c2 = 'CoTableId=%s, 0x0003, ["cotablekey2", 10000], []' % COTABLE2
    nc = '0x0002, ["noncotablekey", 10000], []'
    state = _sorted(rd.parse(_versions(c1) + "\n" + _versions(c2, 1) + "\n" + _versions(nc)))
    keep = lambda dkey, first: [kv for kv in rd.parse(_versions(dkey)) if int(kv[1][6:]) >= first]   # noqa: E731
    want = _sorted(keep(c1, 3) + rd.parse(_versions(c2, 1)) + keep(nc, 2))
    fully_compact(state, None, want, cutoff_ht=us(2000), cotables_cutoff_ht=us(3000))


def test_history_retention_with_colocated_tables():
    """docdb-test-wrapper.cc:2244-2361 (HistoryRetentionWithColocatedTables): HistoryCutoff{invalid, 3000}:
    colocated tables follow the primary cutoff."""
    tabs = ['ColocationId=16384, 0x0001, ["colocationkey", 10000], []', 'ColocationId=16385, 0x0002, ["colocationkey2", 10000], []']
    t3 = 'ColocationId=16386, 0x0003, ["colocationkey3", 10000], []'
    state = _sorted(rd.parse("\n".join(_versions(t) for t in tabs) + "\n" + _versions(t3, 1)))
    want = _sorted([kv for t in tabs for kv in rd.parse(_versions(t)) if int(kv[1][6:]) >= 3] + rd.parse(_versions(t3, 1)))
    fully_compact(state, 3000, want)


def test_history_retention_with_non_colocated_tables():
    """docdb-test-wrapper.cc:2363-2432 (HistoryRetentionWithNonColocatedTables): cutoff 2000."""
    nc = '0x0002, ["noncotablekey", 10000], []'
    extra = 'SubDocKey(DocKey(%s), ["subkey2"; HT{ physical: 1000 }]) -> "value1"' % nc
    state = rd.parse(_versions(nc)[::1]) + rd.parse(extra)
    state = _sorted(state)
    want = _sorted([kv for kv in rd.parse(_versions(nc)) if int(kv[1][6:]) >= 2] + rd.parse(extra))
    fully_compact(state, 2000, want)


@pytest.mark.parametrize("bad", ['SubDocKey(DocKey([], [1.5]), [HT{ physical: 1 }]) -> "x"', "garbage"])
def test_parser_rejects_what_it_does_not_know(bad):
    with pytest.raises(ValueError):
        rd.parse(bad)


def test_redis_ttl_compaction():
    """docdb-ttl-test.cc:674-881 (RedisTTLCompactionTest): value TTLs on whole documents, then TTL merge
    records (merge flags: 1) whose TTL the compaction folds into the value below them (rewritten
    values: docdb_compaction_context.cc:1252-1293). t[i] = 1000 + 1000 i microseconds."""
    s = fully_compact(r'''
SubDocKey(DocKey([], ["k0"]), [HT{ physical: 3000 }]) -> "v0"; ttl: 0.004s
SubDocKey(DocKey([], ["k0"]), [HT{ physical: 1000 }]) -> "v1"; ttl: 0.003s
SubDocKey(DocKey([], ["k1"]), [HT{ physical: 6000 }]) -> "v3"; ttl: 0.001s
SubDocKey(DocKey([], ["k1"]), [HT{ physical: 4000 }]) -> "v2"; ttl: 0.008s
SubDocKey(DocKey([], ["k2"]), [HT{ physical: 12000 }]) -> "v6"
SubDocKey(DocKey([], ["k2"]), [HT{ physical: 8000 }]) -> "v5"; ttl: 0.005s
SubDocKey(DocKey([], ["k2"]), [HT{ physical: 6000 }]) -> "v4"; ttl: 0.003s
SubDocKey(DocKey([], ["k3"]), [HT{ physical: 14000 }]) -> "v9"; ttl: 0.001s
SubDocKey(DocKey([], ["k3"]), [HT{ physical: 5000 }]) -> "v8"
SubDocKey(DocKey([], ["k3"]), [HT{ physical: 2000 }]) -> "v7"; ttl: 0.004s
SubDocKey(DocKey([], ["k4"]), [HT{ physical: 13000 }]) -> DEL
SubDocKey(DocKey([], ["k5"]), [HT{ physical: 10000 }]) -> DEL
SubDocKey(DocKey([], ["k5"]), [HT{ physical: 9000 }]) -> "v:"; ttl: 0.009s
SubDocKey(DocKey([], ["k6"]), [HT{ physical: 9000 }]) -> "v;"; ttl: 0.009s
SubDocKey(DocKey([], ["k6"]), [HT{ physical: 7000 }]) -> DEL
''', 11000, r'''
SubDocKey(DocKey([], ["k2"]), [HT{ physical: 12000 }]) -> "v6"
SubDocKey(DocKey([], ["k2"]), [HT{ physical: 8000 }]) -> "v5"; ttl: 0.005s
SubDocKey(DocKey([], ["k3"]), [HT{ physical: 14000 }]) -> "v9"; ttl: 0.001s
SubDocKey(DocKey([], ["k3"]), [HT{ physical: 5000 }]) -> "v8"
SubDocKey(DocKey([], ["k4"]), [HT{ physical: 13000 }]) -> DEL
SubDocKey(DocKey([], ["k6"]), [HT{ physical: 9000 }]) -> "v;"; ttl: 0.009s
''')
    s = fully_compact(s, 15000, r'''
SubDocKey(DocKey([], ["k2"]), [HT{ physical: 12000 }]) -> "v6"
SubDocKey(DocKey([], ["k3"]), [HT{ physical: 14000 }]) -> "v9"; ttl: 0.001s
SubDocKey(DocKey([], ["k6"]), [HT{ physical: 9000 }]) -> "v;"; ttl: 0.009s
''')
    fully_compact(s, 20000, r'''
SubDocKey(DocKey([], ["k2"]), [HT{ physical: 12000 }]) -> "v6"
''')
    # "Checking TTL rows now"
    fully_compact(r'''
SubDocKey(DocKey([], ["k0"]), [HT{ physical: 6000 }]) -> ""; merge flags: 1; ttl: 0.006s
SubDocKey(DocKey([], ["k0"]), [HT{ physical: 3000 }]) -> ""; merge flags: 1; ttl: 0.004s
SubDocKey(DocKey([], ["k0"]), [HT{ physical: 1000 }]) -> "v0"; ttl: 0.003s
SubDocKey(DocKey([], ["k1"]), [HT{ physical: 6000 }]) -> ""; merge flags: 1; ttl: 0.003s
SubDocKey(DocKey([], ["k1"]), [HT{ physical: 4000 }]) -> "v1"; ttl: 0.008s
SubDocKey(DocKey([], ["k2"]), [HT{ physical: 13000 }]) -> ""; merge flags: 1
SubDocKey(DocKey([], ["k2"]), [HT{ physical: 12000 }]) -> "v6"
SubDocKey(DocKey([], ["k2"]), [HT{ physical: 8000 }]) -> ""; merge flags: 1; ttl: 0.005s
SubDocKey(DocKey([], ["k2"]), [HT{ physical: 6000 }]) -> "v2"; ttl: 0.003s
SubDocKey(DocKey([], ["k3"]), [HT{ physical: 14000 }]) -> "v4"; ttl: 0.001s
SubDocKey(DocKey([], ["k3"]), [HT{ physical: 5000 }]) -> ""; merge flags: 1
SubDocKey(DocKey([], ["k3"]), [HT{ physical: 2000 }]) -> "v3"; ttl: 0.004s
SubDocKey(DocKey([], ["k4"]), [HT{ physical: 13000 }]) -> DEL
SubDocKey(DocKey([], ["k5"]), [HT{ physical: 10000 }]) -> DEL
SubDocKey(DocKey([], ["k5"]), [HT{ physical: 9000 }]) -> "v5"; ttl: 0.009s
SubDocKey(DocKey([], ["k6"]), [HT{ physical: 11000 }]) -> ""; merge flags: 1; ttl: 0.004s
SubDocKey(DocKey([], ["k6"]), [HT{ physical: 9000 }]) -> "v6"; ttl: 0.009s
SubDocKey(DocKey([], ["k6"]), [HT{ physical: 7000 }]) -> DEL
''', 10000, r'''
SubDocKey(DocKey([], ["k0"]), [HT{ physical: 1000 }]) -> "v0"; ttl: 0.011s
SubDocKey(DocKey([], ["k2"]), [HT{ physical: 13000 }]) -> ""; merge flags: 1
SubDocKey(DocKey([], ["k2"]), [HT{ physical: 12000 }]) -> "v6"
SubDocKey(DocKey([], ["k2"]), [HT{ physical: 6000 }]) -> "v2"; ttl: 0.007s
SubDocKey(DocKey([], ["k3"]), [HT{ physical: 14000 }]) -> "v4"; ttl: 0.001s
SubDocKey(DocKey([], ["k3"]), [HT{ physical: 2000 }]) -> "v3"
SubDocKey(DocKey([], ["k4"]), [HT{ physical: 13000 }]) -> DEL
SubDocKey(DocKey([], ["k6"]), [HT{ physical: 11000 }]) -> ""; merge flags: 1; ttl: 0.004s
SubDocKey(DocKey([], ["k6"]), [HT{ physical: 9000 }]) -> "v6"; ttl: 0.009s
''')


def test_redis_collection_ttl_compaction_chain():
    """docdb-ttl-test.cc:72-672 (RedisCollectionTTLCompactionTest): 13 successive history compactions of
    collections with init markers, collection-level TTL merge records, tombstoned and re-created
    collections. Dumps extracted verbatim by tests/golden/extract_reference_dumps.py."""
    import json
    import os
    fx = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "redis_collection_ttl_compaction.json")))
    assert len(fx["steps"]) == 13
    state = rd.parse(fx["initial"])
    for step in fx["steps"]:
        state = fully_compact(state, step["cutoff_us"], step["expected"])


COLLECTION_WITH_TTL = r'''
            SubDocKey(DocKey([], ["c"]), [HT{ physical: 1000 }]) -> {}; ttl: 10.000s               // file 1
            SubDocKey(DocKey([], ["c"]), ["k0"; HT{ physical: 1100 }]) -> "vv0"; ttl: 20.000s      // file 2
            SubDocKey(DocKey([], ["c"]), ["k0"; HT{ physical: 1000 w: 1 }]) -> "v0"; ttl: 10.000s  // file 1
            SubDocKey(DocKey([], ["c"]), ["k1"; HT{ physical: 1100 }]) -> "vv1"; ttl: 21.000s      // file 3
            SubDocKey(DocKey([], ["c"]), ["k1"; HT{ physical: 1000 w: 2 }]) -> "v1"; ttl: 10.000s  // file 1
            SubDocKey(DocKey([], ["c"]), ["k2"; HT{ physical: 1100 }]) -> "vv2"; ttl: 22.000s      // file 4
            SubDocKey(DocKey([], ["c"]), ["k2"; HT{ physical: 1000 w: 3 }]) -> "v2"; ttl: 10.000s  // file 1
            SubDocKey(DocKey([], ["c"]), ["k3"; HT{ physical: 1100 }]) -> "vv3"; ttl: 23.000s      // file 5
            SubDocKey(DocKey([], ["c"]), ["k4"; HT{ physical: 1100 }]) -> "vv4"; ttl: 24.000s      // file 6
            SubDocKey(DocKey([], ["c"]), ["k5"; HT{ physical: 1100 }]) -> "vv5"; ttl: 25.000s      // file 7
'''


def test_compaction_for_collections_with_ttl():
    """docdb-ttl-test.cc:1090-1098 + docdb-test.h:394-425 (TestCompactionForCollectionsWithTTL): after the
    collection's init marker and first values expire, a major compaction leaves no tombstone for them."""
    fully_compact(COLLECTION_WITH_TTL, 1050 + 10 * 1000000, r'''
            SubDocKey(DocKey([], ["c"]), ["k0"; HT{ physical: 1100 }]) -> "vv0"; ttl: 20.000s
            SubDocKey(DocKey([], ["c"]), ["k1"; HT{ physical: 1100 }]) -> "vv1"; ttl: 21.000s
            SubDocKey(DocKey([], ["c"]), ["k2"; HT{ physical: 1100 }]) -> "vv2"; ttl: 22.000s
            SubDocKey(DocKey([], ["c"]), ["k3"; HT{ physical: 1100 }]) -> "vv3"; ttl: 23.000s
            SubDocKey(DocKey([], ["c"]), ["k4"; HT{ physical: 1100 }]) -> "vv4"; ttl: 24.000s
            SubDocKey(DocKey([], ["c"]), ["k5"; HT{ physical: 1100 }]) -> "vv5"; ttl: 25.000s
''')


def test_minor_compactions_for_collections_with_ttl():
    """docdb-ttl-test.cc:1117-1161 (MinorCompactionsForCollectionsWithTTL): compactions of some of the
    files ("// file N" tags of the reference dumps); expired values become delete markers because older
    data may exist in the files left out (docdb_compaction_context.cc:1268-1277)."""
    def minor(state, files, cutoff_us, expected_dump, new_file):
        runs = [w.sort_run([(o.ikey(k, SEQ0 + f * 100 + i), v) for i, (k, v, ff) in enumerate(state) if ff == f]) for f in files]
        p = o.CompactionParams(bottommost=False, cutoff_ht=us(cutoff_us), other_min_ht=o.HT_MIN)
        want_all = rd.parse(expected_dump, with_files=True)
        want = [(k, v) for k, v, f in want_all if f == new_file]
        assert [(k[:-8], v) for k, v in o.compact_runs(runs, p).kv_list()] == want
        assert [(k[:-8], v) for k, v in hh.compact_runs(runs, p)] == want
        if EXTRA_CHECK is not None:
            EXTRA_CHECK(runs, dict(bottommost=False, cutoff_ht=us(cutoff_us), other_min_ht=o.HT_MIN), want)
        untouched = [e for e in state if e[2] not in files]
        assert sorted((k, v) for k, v, f in want_all if f != new_file) == sorted((k, v) for k, v, _ in untouched)
        return want_all

    s0 = rd.parse(COLLECTION_WITH_TTL, with_files=True)
    # MinorCompaction(cutoff, num_files_to_compact = 2, start_index = 1): files 2 and 3 -> file 8
    s1 = minor(s0, [2, 3], 1100 + 20 * 1000000 + 1, r'''
SubDocKey(DocKey([], ["c"]), [HT{ physical: 1000 }]) -> {}; ttl: 10.000s               // file 1
SubDocKey(DocKey([], ["c"]), ["k0"; HT{ physical: 1100 }]) -> DEL                      // file 8
SubDocKey(DocKey([], ["c"]), ["k0"; HT{ physical: 1000 w: 1 }]) -> "v0"; ttl: 10.000s  // file 1
SubDocKey(DocKey([], ["c"]), ["k1"; HT{ physical: 1100 }]) -> "vv1"; ttl: 21.000s      // file 8
SubDocKey(DocKey([], ["c"]), ["k1"; HT{ physical: 1000 w: 2 }]) -> "v1"; ttl: 10.000s  // file 1
SubDocKey(DocKey([], ["c"]), ["k2"; HT{ physical: 1100 }]) -> "vv2"; ttl: 22.000s      // file 4
SubDocKey(DocKey([], ["c"]), ["k2"; HT{ physical: 1000 w: 3 }]) -> "v2"; ttl: 10.000s  // file 1
SubDocKey(DocKey([], ["c"]), ["k3"; HT{ physical: 1100 }]) -> "vv3"; ttl: 23.000s      // file 5
SubDocKey(DocKey([], ["c"]), ["k4"; HT{ physical: 1100 }]) -> "vv4"; ttl: 24.000s      // file 6
SubDocKey(DocKey([], ["c"]), ["k5"; HT{ physical: 1100 }]) -> "vv5"; ttl: 25.000s      // file 7
''', 8)
    # "Compact files 4, 5, 6, 7, 8" -> file 9
    minor(s1, [4, 5, 6, 7, 8], 1100 + 24 * 1000000 + 1, r'''
SubDocKey(DocKey([], ["c"]), [HT{ physical: 1000 }]) -> {}; ttl: 10.000s               // file 1
SubDocKey(DocKey([], ["c"]), ["k0"; HT{ physical: 1100 }]) -> DEL                      // file 9
SubDocKey(DocKey([], ["c"]), ["k0"; HT{ physical: 1000 w: 1 }]) -> "v0"; ttl: 10.000s  // file 1
SubDocKey(DocKey([], ["c"]), ["k1"; HT{ physical: 1100 }]) -> DEL                      // file 9
SubDocKey(DocKey([], ["c"]), ["k1"; HT{ physical: 1000 w: 2 }]) -> "v1"; ttl: 10.000s  // file 1
SubDocKey(DocKey([], ["c"]), ["k2"; HT{ physical: 1100 }]) -> DEL                      // file 9
SubDocKey(DocKey([], ["c"]), ["k2"; HT{ physical: 1000 w: 3 }]) -> "v2"; ttl: 10.000s  // file 1
SubDocKey(DocKey([], ["c"]), ["k3"; HT{ physical: 1100 }]) -> DEL                      // file 9
SubDocKey(DocKey([], ["c"]), ["k4"; HT{ physical: 1100 }]) -> DEL                      // file 9
SubDocKey(DocKey([], ["c"]), ["k5"; HT{ physical: 1100 }]) -> "vv5"; ttl: 25.000s      // file 9
''', 9)
