"""Pins the oracle's DocDBCompactionFeed restatement against the reference's golden tests
(textual dumps of every surviving KV)."""
import dockv_util as dk
import oracle_py as o

SEQ0 = 1 << 50


def us(n):
    return o.ht_from_micros(n)


def run_of(entries, seq_start):
    """entries: list of (user_key, value) -> sorted run of (internal key, value)."""
    kvs = [(o.ikey(k, seq_start + i), v) for i, (k, v) in enumerate(entries)]
    kvs.sort(key=lambda kv: (kv[0][:-8], -int.from_bytes(kv[0][-8:], "little")))
    return kvs


def user_kvs(result):
    return [(k[:-8], v) for k, v in result.kv_list()]


def compact(runs, cutoff_us, major, bottommost=None, **kw):
    p = o.CompactionParams(bottommost=major if bottommost is None else bottommost,
                           cutoff_ht=us(cutoff_us),
                           other_min_ht=o.HT_MAX if major else o.HT_MIN, **kw)
    return o.compact_runs(runs, p)


def test_history_compaction_first_row_handling_regression():
    # docdb/docdb-test-wrapper.cc:99-133
    d = dk.doc_key(["mydockey", dk.INT_KEY1])
    run = run_of([
        (dk.sub_doc_key(d, [], micros=4000), dk.OBJECT),
        (dk.sub_doc_key(d, [], micros=1000), dk.OBJECT),
        (dk.sub_doc_key(d, ["subkey1"], micros=3000), dk.vstr("value3")),
        (dk.sub_doc_key(d, ["subkey1"], micros=2000), dk.vstr("value2")),
        (dk.sub_doc_key(d, ["subkey1"], micros=1000), dk.vstr("value1")),
    ], SEQ0)
    got = user_kvs(compact([run], 3500, major=True))
    assert got == [
        (dk.sub_doc_key(d, [], micros=4000), dk.OBJECT),
        (dk.sub_doc_key(d, [], micros=1000), dk.OBJECT),
        (dk.sub_doc_key(d, ["subkey1"], micros=3000), dk.vstr("value3")),
    ]


def _minor_sequence(values):
    """Six single-entry files f1..f6 (f_i at i*1000us); compact the two newest repeatedly with
    cutoff 5000 (docdb-test-wrapper.cc:695-842). Returns the dump after each step."""
    d = dk.doc_key(["k"])
    files = [run_of([(dk.sub_doc_key(d, [], micros=i * 1000), values[i])], SEQ0 + i) for i in range(1, 7)]
    dumps = []
    while len(files) > 1:
        newest2 = files[-2:]
        major = len(files) == 2           # "file_names.size() == compaction_input_file_names.size()"
        res = compact(newest2, 5000, major=major)
        merged = res.kv_list()
        files = files[:-2] + [merged]
        allkv = sorted((kv for f in files for kv in f),
                       key=lambda kv: (kv[0][:-8], -int.from_bytes(kv[0][-8:], "little")))
        dumps.append([(int.from_bytes(b"", "big") or k[:-8], v) for k, v in allkv])
    return d, dumps


def test_minor_compaction_no_deletions():
    vals = {i: dk.vstr("v%d" % i) for i in range(1, 7)}
    d, dumps = _minor_sequence(vals)

    def expect(times):
        return [(dk.sub_doc_key(d, [], micros=t * 1000), vals[t]) for t in times]
    assert dumps[0] == expect([6, 5, 4, 3, 2, 1])
    assert dumps[1] == expect([6, 5, 3, 2, 1])
    assert dumps[2] == expect([6, 5, 2, 1])
    assert dumps[3] == expect([6, 5, 1])
    assert dumps[4] == expect([6, 5])


def test_minor_compaction_with_deletions():
    vals = {i: dk.vstr("v%d" % i) for i in range(1, 7)}
    vals[5] = dk.TOMBSTONE
    d, dumps = _minor_sequence(vals)

    def expect(times):
        return [(dk.sub_doc_key(d, [], micros=t * 1000), vals[t]) for t in times]
    assert dumps[0] == expect([6, 5, 4, 3, 2, 1])
    assert dumps[1] == expect([6, 5, 3, 2, 1])
    assert dumps[2] == expect([6, 5, 2, 1])
    assert dumps[3] == expect([6, 5, 1])
    # last step is a major compaction: the tombstone is gone too
    assert dumps[4] == expect([6])


def test_overwrite_stack_worked_example():
    # docdb/docdb_compaction_context.cc:842-868 (history_cutoff = 25)
    d = dk.doc_key(["doc_key1"])
    entries = [
        (dk.sub_doc_key(d, [], micros=30), dk.OBJECT),            # keep (above cutoff)
        (dk.sub_doc_key(d, [], micros=20), dk.TOMBSTONE),         # keep in minor (20 >= MinHT)
        (dk.sub_doc_key(d, [], micros=10), dk.OBJECT),            # 10 < 20 -> deleted
        (dk.sub_doc_key(d, ["subkey1"], micros=35), dk.vstr("value4")),   # keep
        (dk.sub_doc_key(d, ["subkey1"], micros=23), dk.vstr("value3")),   # keep: 23 >= 20
        (dk.sub_doc_key(d, ["subkey1"], micros=21), dk.vstr("value2")),   # 21 < 23 -> deleted
        (dk.sub_doc_key(d, ["subkey1"], micros=15), dk.vstr("value1")),   # deleted
    ]
    got = user_kvs(compact([run_of(entries, SEQ0)], 25, major=False))
    assert got == [entries[0], entries[1], entries[3], entries[4]]
    # In a major compaction the tombstone itself is GC'ed but still shadows.
    got = user_kvs(compact([run_of(entries, SEQ0)], 25, major=True))
    assert got == [entries[0], entries[3], entries[4]]


def test_second_worked_example_cutoff_12():
    # docdb/docdb_compaction_context.cc:1031-1046
    d = dk.doc_key(["k1"])
    e = [
        (dk.sub_doc_key(d, [], micros=10), dk.OBJECT),
        (dk.sub_doc_key(d, [], micros=5), dk.OBJECT),                 # 5 < 10 removed
        (dk.sub_doc_key(d, ["col1"], micros=11), dk.vstr("a")),
        (dk.sub_doc_key(d, ["col1"], micros=7), dk.vstr("b")),        # 7 < 11 removed
        (dk.sub_doc_key(d, ["col2"], micros=9), dk.vstr("c")),        # 9 < 10 removed
    ]
    got = user_kvs(compact([run_of(e, SEQ0)], 12, major=True))
    assert got == [e[0], e[2]]


def test_obsolete_intent_prefix_dropped_and_bounds():
    d = dk.doc_key(["a"])
    e = [(b"\x0a" + b"junk" + b"#" + o.encode_doc_ht(5), b"Sx"),
         (dk.sub_doc_key(d, [dk.kcol(1)], micros=100), dk.vstr("v"))]
    got = user_kvs(compact([run_of(e, SEQ0)], 50, major=True))
    assert got == [e[1]]


def test_ttl_expiry_minor_vs_major():
    # docdb/docdb_compaction_context.cc:1257-1277; dockv/doc_ttl_util.cc
    d = dk.doc_key(["r"])
    base = o.YB_EPOCH_US + 10_000_000
    e = [(dk.sub_doc_key(d, [dk.kcol(1)], micros=base), dk.with_ttl(dk.vstr("v"), 1000))]   # 1 s TTL
    cutoff = base + 5_000_000
    # major: expired => dropped
    assert user_kvs(compact([run_of(e, SEQ0)], cutoff, major=True)) == []
    # minor: expired => rewritten as tombstone
    assert user_kvs(compact([run_of(e, SEQ0)], cutoff, major=False)) == [(e[0][0], dk.TOMBSTONE)]
    # not yet expired at cutoff
    assert user_kvs(compact([run_of(e, SEQ0)], base + 500_000, major=True)) == e
    # table-level TTL applies when the value has none
    e2 = [(dk.sub_doc_key(d, [dk.kcol(1)], micros=base), dk.vstr("v"))]
    assert user_kvs(compact([run_of(e2, SEQ0)], cutoff, major=True, table_ttl_ns=10**9)) == []
    assert user_kvs(compact([run_of(e2, SEQ0)], cutoff, major=True)) == e2


def test_intent_doc_ht_is_stripped():
    # docdb/docdb_compaction_context.cc:1298-1307
    d = dk.doc_key(["r"])
    v = dk.with_intent_ht(dk.vstr("payload"), o.YB_EPOCH_US + 5, 0, 3)
    e = [(dk.sub_doc_key(d, [dk.kcol(1)], micros=o.YB_EPOCH_US + 10), v)]
    got = user_kvs(compact([run_of(e, SEQ0)], o.YB_EPOCH_US + 100, major=True))
    assert got == [(e[0][0], dk.vstr("payload"))]
    # above cutoff: untouched
    got = user_kvs(compact([run_of(e, SEQ0)], o.YB_EPOCH_US + 1, major=True))
    assert got == e
