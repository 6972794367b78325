"""Shared randomized DocDB-shaped workloads for the parity tests (CPU device-logic harness and GPU)."""
import random

import dockv_util as dk
import oracle_py as o

BASE_US = o.YB_EPOCH_US + 100_000_000


def sort_run(kvs):
    kvs.sort(key=lambda kv: (kv[0][:-8], -int.from_bytes(kv[0][-8:], "little")))
    return kvs


def random_docdb_runs(seed, n_runs=4, n_rows=40, deep=True, ttl=True, dup_rate=0.1, n_ht=12):
    """Rows with nested subkeys, tombstones, TTLs, merge (TTL) rows, intent doc HTs, and exact
    duplicate user keys across runs (rule A). Returns list of sorted runs of (ikey, value)."""
    rng = random.Random(seed)
    runs = [[] for _ in range(n_runs)]
    seq = [(1 << 50) + (r << 30) for r in range(n_runs)]
    used = set()

    def put(user_key, value):
        r = rng.randrange(n_runs)
        seq[r] += 1
        runs[r].append((o.ikey(user_key, seq[r]), value))
        if rng.random() < dup_rate:
            r2 = rng.randrange(n_runs)
            if r2 != r:
                seq[r2] += 1
                runs[r2].append((o.ikey(user_key, seq[r2]), dk.vstr("dup%d" % rng.randrange(100))))

    def rand_value():
        x = rng.random()
        if x < 0.15:
            return dk.TOMBSTONE
        v = dk.vstr("v" + "x" * rng.randrange(0, 40) + str(rng.randrange(1000)))
        if x < 0.2:
            v = dk.OBJECT
        if ttl and rng.random() < 0.15:
            v = dk.with_ttl(v, rng.choice([0, 1, 50, 5000, 10**7]))
        if rng.random() < 0.08:
            v = dk.with_intent_ht(v, BASE_US + rng.randrange(n_ht * 10), 0, rng.randrange(4))
        if ttl and rng.random() < 0.05:
            v = dk.ttl_merge_row(rng.choice([1, 100, 10**6]))
        return v

    for row in range(n_rows):
        style = rng.random()
        if style < 0.5:
            d = dk.doc_key([rng.choice(["r", "row", "k"]) + str(row), rng.randrange(5)],
                           hash_code=rng.randrange(65536), hashed=["h%d" % row])
        else:
            d = dk.doc_key(["plain%04d" % row] + ([rng.randrange(-5, 5)] if rng.random() < 0.5 else []))
        paths = [[]] if rng.random() < 0.5 else []
        for c in range(rng.randrange(1, 5)):
            p = [dk.kcol(c + 1) if rng.random() < 0.7 else "sub%d" % c]
            paths.append(p)
            if deep and rng.random() < 0.4:
                for j in range(rng.randrange(1, 3)):
                    paths.append(p + ["leaf%d" % j])
                    if rng.random() < 0.3:
                        paths.append(p + ["leaf%d" % j, dk.kint64(j)])
        for p in paths:
            for _ in range(rng.randrange(1, 5)):
                ht = (BASE_US + rng.randrange(n_ht) * 10, rng.randrange(2), rng.randrange(3))
                uk = dk.sub_doc_key(d, p, ht=ht)
                if uk in used:
                    continue
                used.add(uk)
                put(uk, rand_value())
    return [sort_run(r) for r in runs]


def param_grid():
    cut = [o.HT_MIN, o.ht_from_micros(BASE_US + 35), o.ht_from_micros(BASE_US + 75, 1), o.ht_from_micros(BASE_US + 10**6)]
    out = []
    for c in cut:
        for major in (True, False):
            out.append(dict(bottommost=major, cutoff_ht=c, other_min_ht=o.HT_MAX if major else o.HT_MIN))
    out.append(dict(bottommost=True, cutoff_ht=cut[2], other_min_ht=o.HT_MAX, retain_delete_markers=True))
    out.append(dict(bottommost=True, cutoff_ht=cut[3], other_min_ht=o.ht_from_micros(BASE_US + 50), table_ttl_ns=20 * 10**3))
    out.append(dict(bottommost=False, cutoff_ht=cut[1], other_min_ht=o.HT_MIN, table_ttl_ns=10**9))
    return out


def random_cotable_runs(seed, n_runs=3, n_tables=4, rows_per_table=12, colocated=True, n_ht=12):
    """Colocated ('0' + u32 id) or cotable ('y' + 16-byte uuid) tables with optional table tombstones
    (id ! # HT) at several hybrid times, mixed with plain rows of an id-less table."""
    rng = random.Random(seed)
    runs = [[] for _ in range(n_runs)]
    seq = [(1 << 50) + (r << 30) for r in range(n_runs)]

    def put(user_key, value):
        r = rng.randrange(n_runs)
        seq[r] += 1
        runs[r].append((o.ikey(user_key, seq[r]), value))
    used = set()
    for t in range(n_tables):
        kw = dict(colocation=1000 + t) if colocated else dict(cotable=bytes([(t * 37 + j) % 251 + 1 for j in range(16)]))
        n_tomb = rng.choice([0, 1, 2, 3])
        for _ in range(n_tomb):
            k = dk.table_tombstone_key(micros=BASE_US + rng.randrange(n_ht) * 10, **kw)
            if k not in used:
                used.add(k)
                put(k, dk.TOMBSTONE)
        for row in range(rows_per_table):
            d = dk.doc_key(["r%03d" % row], **kw) if rng.random() < 0.5 else dk.doc_key([row], hash_code=rng.randrange(65536), hashed=["h%d" % row], **kw)
            for c in range(rng.randrange(1, 4)):
                for _ in range(rng.randrange(1, 4)):
                    uk = dk.sub_doc_key(d, [dk.kcol(c + 1)], ht=(BASE_US + rng.randrange(n_ht) * 10, rng.randrange(2), 0))
                    if uk in used:
                        continue
                    used.add(uk)
                    put(uk, dk.TOMBSTONE if rng.random() < 0.15 else dk.vstr("v%d" % rng.randrange(1000)))
    for row in range(10):
        uk = dk.sub_doc_key(dk.doc_key(["plain%d" % row]), [dk.kcol(1)], ht=(BASE_US + rng.randrange(n_ht) * 10, 0, 0))
        if uk not in used:
            used.add(uk)
            put(uk, dk.vstr("p"))
    return [sort_run(r) for r in runs]


def random_numeric_key_runs(seed, n_runs=3, n_rows=40, n_ht=12):
    """Rows whose DocKeys / subkeys carry kVarInt / kDecimal / kBson components (YSQL numeric, YCQL varint /
    decimal primary keys, bson keys) in both sort orders; several versions per column, tombstones."""
    rng = random.Random(seed)
    runs = [[] for _ in range(n_runs)]
    seq = [(1 << 50) + (r << 30) for r in range(n_runs)]
    used = set()

    def numeric():
        x = rng.randrange(5)
        if x == 4:                                           # kBson / kBsonDescending: (complement) zero-encoded bytes
            raw = bytes(rng.randrange(256) for _ in range(rng.randrange(0, 12)))
            return dk.kbson(raw) if rng.random() < 0.5 else dk.kbson_desc(raw)
        if x == 0:
            return dk.kvarint(rng.randrange(-10**rng.randrange(1, 30), 10**rng.randrange(1, 30)))
        if x == 1:
            return dk.kvarint_desc(rng.randrange(-10**6, 10**6))
        digits = [rng.randrange(1, 10)] + [rng.randrange(10) for _ in range(rng.randrange(0, 9))] + [rng.randrange(1, 10)]
        if rng.random() < 0.1:
            digits = []
        f = dk.kdecimal if x == 2 else dk.kdecimal_desc
        return f(digits, rng.randrange(-400, 400), rng.random() < 0.6)

    for row in range(n_rows):
        if rng.random() < 0.5:
            d = dk.doc_key([numeric()] + (["s%d" % row] if rng.random() < 0.5 else []), hash_code=rng.randrange(65536), hashed=[numeric(), "h%d" % row])
        else:
            d = dk.doc_key([numeric(), numeric(), row])
        paths = [[dk.kcol(c + 1)] for c in range(rng.randrange(1, 4))]
        if rng.random() < 0.4:
            paths.append([dk.kcol(7), numeric()])
        for p in paths:
            for _ in range(rng.randrange(1, 5)):
                uk = dk.sub_doc_key(d, p, ht=(BASE_US + rng.randrange(n_ht) * 10, rng.randrange(2), rng.randrange(3)))
                if uk in used:
                    continue
                used.add(uk)
                r = rng.randrange(n_runs)
                seq[r] += 1
                runs[r].append((o.ikey(uk, seq[r]), dk.TOMBSTONE if rng.random() < 0.15 else dk.vstr("v%d" % rng.randrange(1000))))
    return [sort_run(r) for r in runs]


def giant_row_runs(seed, n_runs=3, cols=250, versions=20, collection=2000, small_rows=30, colocated=False, n_ht=40):
    """Rows far larger than a merge tile: one row of cols x versions column entries with row-level markers and
    tombstones (`DocKey # HT`), one row holding a collection of `collection` elements x 3 versions under a column
    that itself has several versions / tombstones (`DocKey col # HT`), surrounded by small rows; optionally inside a
    colocated table with table tombstones. The ancestors' overwrite times shadow entries thousands of records later."""
    rng = random.Random(seed)
    runs = [[] for _ in range(n_runs)]
    seq = [(1 << 50) + (r << 30) for r in range(n_runs)]
    used = set()
    kw = dict(colocation=77) if colocated else {}

    def put(uk, value):
        if uk in used:
            return
        used.add(uk)
        r = rng.randrange(n_runs)
        seq[r] += 1
        runs[r].append((o.ikey(uk, seq[r]), value))

    def ht():
        return (BASE_US + rng.randrange(n_ht) * 10, rng.randrange(2), rng.randrange(3))

    def val():
        x = rng.random()
        return dk.TOMBSTONE if x < 0.12 else dk.vstr("v%d" % rng.randrange(10**6) + "y" * rng.randrange(0, 30))

    if colocated:
        for _ in range(2):
            put(dk.table_tombstone_key(micros=BASE_US + rng.randrange(n_ht) * 10, **kw), dk.TOMBSTONE)
    for i in range(small_rows // 2):
        put(dk.sub_doc_key(dk.doc_key(["a%03d" % i], **kw), [dk.kcol(1)], ht=ht()), val())
    wide = dk.doc_key(["m-wide"], hash_code=4242, hashed=["w"], **kw)
    for _ in range(3):
        put(dk.sub_doc_key(wide, [], ht=ht()), rng.choice([dk.OBJECT, dk.TOMBSTONE, dk.OBJECT]))
    for c in range(cols):
        for _ in range(versions):
            put(dk.sub_doc_key(wide, [dk.kcol(c + 1)], ht=ht()), val())
    coll = dk.doc_key(["n-coll"], **kw)
    put(dk.sub_doc_key(coll, [], ht=ht()), dk.OBJECT)
    for _ in range(4):
        put(dk.sub_doc_key(coll, [dk.kcol(7)], ht=ht()), rng.choice([dk.OBJECT, dk.TOMBSTONE]))
    for e in range(collection):
        for _ in range(3):
            put(dk.sub_doc_key(coll, [dk.kcol(7), "elem%06d" % e], ht=ht()), val())
        if e % 97 == 0:
            for _ in range(2):
                put(dk.sub_doc_key(coll, [dk.kcol(7), "elem%06d" % e, dk.kint64(e)], ht=ht()), val())
    for c in range(3):
        put(dk.sub_doc_key(coll, [dk.kcol(8 + c)], ht=ht()), val())
    for i in range(small_rows // 2):
        put(dk.sub_doc_key(dk.doc_key(["z%03d" % i], **kw), [dk.kcol(1)], ht=ht()), val())
    return [sort_run(r) for r in runs]
