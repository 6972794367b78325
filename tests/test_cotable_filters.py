"""Per-database cotable HybridTime filters of an input file (the tail of FdWithBoundaries::user_filter_data:
docdb/docdb_rocksdb_util.cc:503-509; HybridTimeFilteringIterator::Satisfied, :525-565 — set on the master's sys catalog
by a restore). No reference test holds vectors for it, so the oracle is checked against the rule applied by hand: a
compaction WITH the filters equals a compaction of inputs from which the hidden entries were removed beforehand."""
import importlib
import os
import struct
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import oracle_py as o          # noqa: E402
import workloads as w          # noqa: E402


def _uuid(t):
    return bytes([(t * 37 + j) % 251 + 1 for j in range(16)])


def _db_oid(t):
    return struct.unpack("<I", _uuid(t)[12:16])[0]


def _ht_of(user_key):
    n = user_key[-1] & 0x1f                       # DocHybridTime::EncodedFromEnd: the length sits in the last byte
    micros, logical, _ = o.decode_doc_ht(user_key[-n:])
    return (micros << 12) | logical


def _hidden(user_key, global_filter, cot):
    """The reference rule, written from docdb_rocksdb_util.cc:525-565."""
    ht = _ht_of(user_key)
    if global_filter != o.HT_INVALID and ht > global_filter:
        return True
    if not cot or user_key[:1] != b"y":
        return False
    oid = struct.unpack("<I", user_key[13:17])[0]
    for d, f in zip(*cot):
        if d == oid:
            return ht > f
    return False


def _case(seed):
    runs = w.random_cotable_runs(seed, n_runs=3, n_tables=5, rows_per_table=10, colocated=False)
    tabs = sorted(range(5), key=_db_oid)
    f_lo, f_mid = o.ht_from_micros(w.BASE_US + 35), o.ht_from_micros(w.BASE_US + 75, 1)
    # file 0: two databases filtered; file 1: one database + a global filter; file 2: a database no key belongs to
    cot = [([_db_oid(tabs[0]), _db_oid(tabs[2])], [f_lo, f_mid]), ([_db_oid(tabs[3])], [f_lo]), ([7], [f_lo])]
    for oids, _ in cot:
        assert oids == sorted(oids)
    glob = [o.HT_INVALID, o.ht_from_micros(w.BASE_US + 95), o.HT_INVALID]
    return runs, glob, cot


@pytest.mark.parametrize("seed", range(4))
def test_oracle_applies_the_rule(seed):
    runs, glob, cot = _case(seed)
    ssts = [o.Sst.build(r, o.TableOptions(block_size=1024)) for r in runs]
    kept_runs = [[(k, v) for k, v in r if not _hidden(k[:-8], glob[i], cot[i])] for i, r in enumerate(runs)]
    assert sum(len(r) for r in kept_runs) < sum(len(r) for r in runs)        # the filters do hide something
    plain = [o.Sst.build(r, o.TableOptions(block_size=1024)) for r in kept_runs]
    # Compaction::GetLargestUserKey comes from the files' metadata, i.e. from the UNFILTERED files (db/compaction.cc:318)
    largest = max(k[:-8] for r in runs for k, _ in r)
    for kw in w.param_grid()[:5]:
        got = o.compact(ssts, o.CompactionParams(**kw), o.TableOptions(block_size=1024), ht_filters=glob, cotable_filters=cot)
        want = o.compact(plain, o.CompactionParams(largest_user_key=largest, **kw), o.TableOptions(block_size=1024))
        assert got.kv_list() == want.kv_list()
        assert got.stats.num_input_records == want.stats.num_input_records   # hidden entries are not even counted


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(4))
def test_gpu_applies_cotable_filters(seed):
    pkg = importlib.import_module("yugabyte-db_b200")
    runs, glob, cot = _case(seed)
    for enc in (1, 2):                                # fused ingest pass / general decode kernels
        ssts = [o.Sst.build(r, o.TableOptions(block_size=1024, key_encoding=enc)) for r in runs]
        for kw in w.param_grid()[:5]:
            exp = o.compact(ssts, o.CompactionParams(**kw), o.TableOptions(block_size=2048), ht_filters=glob, cotable_filters=cot)
            job = pkg.GpuCompactionJob(block_size=2048, **kw)
            for i, s in enumerate(ssts):
                job.add_input_sst(s.meta_view(), s.data_view(), ht_filter=glob[i])
                job.set_cotable_filters(*cot[i])
            job.run()
            assert job.kv_list() == exp.kv_list()
            st = job.stats()
            assert st.num_input_records == exp.stats.num_input_records and st.num_output_records == exp.stats.num_output_records
            data, meta = job.fetch_output()
            ref = exp.sst()
            assert (data.tobytes(), meta.tobytes()) == ((ref.data, ref.meta) if ref is not None else (b"", b""))
    # the pipelined path hands the filters to every range's job
    ssts = [o.Sst.build(r, o.TableOptions(block_size=1024)) for r in runs]
    kw = w.param_grid()[2]
    exp = o.compact(ssts, o.CompactionParams(**kw), o.TableOptions(block_size=2048), ht_filters=glob, cotable_filters=cot)
    res = pkg.compact_files([(s.meta_view(), s.data_view()) for s in ssts], max_subcompactions=3, max_in_flight=2, ht_filters=glob,
                            cotable_filters=cot, block_size=2048, **kw)
    got = []
    for data, meta in res.files():
        got += o.Sst.from_bytes(meta.tobytes(), data.tobytes()).read_all()
    assert got == exp.kv_list()
