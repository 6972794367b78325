"""GPU parity: the CUDA path, called through the C ABI, against the oracle on the same inputs.
Bit-exact KV stream (integer/byte work), stats, and output SST bytes."""
import importlib

import numpy as np
import pytest

import dockv_util as dk
import oracle_py as o
import workloads as w

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    m = importlib.import_module("yugabyte-db_b200")
    assert m.device_count() >= 1, "GPU tests need a CUDA device"
    return m


def gpu_compact(pkg, ssts, ht_filters=None, **kw):
    job = pkg.GpuCompactionJob(**kw)
    for i, s in enumerate(ssts):
        job.add_input_sst(s.meta_view(), s.data_view(), ht_filter=(ht_filters[i] if ht_filters else pkg.HT_INVALID))
    job.run()
    return job


def okw(kw):
    """binding kwargs -> oracle CompactionParams kwargs"""
    m = dict(kw)
    m.pop("block_size", None)
    m.pop("output_key_encoding", None)
    m.pop("filter_policy", None)
    m.pop("filter_block_size", None)
    m.pop("output_compression", None)
    return m


def check(pkg, ssts, block_size=4096, ht_filters=None, **kw):
    topt = o.TableOptions(block_size=block_size, key_encoding=kw.get("output_key_encoding", 1),
                          filter_policy=kw.get("filter_policy", 0), filter_block_size=kw.get("filter_block_size", 65536),
                          compression=kw.get("output_compression", 0))
    exp = o.compact(ssts, o.CompactionParams(**okw(kw)), topt, ht_filters=ht_filters)
    job = gpu_compact(pkg, ssts, ht_filters=ht_filters, block_size=block_size, **kw)
    st = job.stats()
    # FileMetaData::smallest / largest come from the block encoder's boundary records (no KV stream)
    ekv = exp.kv_list()
    assert job.boundaries() == ((ekv[0][0], ekv[-1][0]) if ekv else (b"", b""))
    assert job.kv_list() == ekv
    es = exp.stats
    assert st.num_input_records == es.num_input_records
    assert st.num_output_records == es.num_output_records
    assert st.num_record_drop_hidden == es.num_dropped_hidden
    assert st.num_record_drop_obsolete == es.num_dropped_obsolete
    assert st.num_record_drop_feed == es.num_dropped_feed
    assert st.total_input_raw_key_bytes == es.in_key_bytes and st.total_input_raw_value_bytes == es.in_val_bytes
    assert st.total_output_raw_key_bytes == es.out_key_bytes and st.total_output_raw_value_bytes == es.out_val_bytes
    assert job.digest() == es.kv_hash
    data, meta = job.fetch_output()
    ref = exp.sst()
    if ref is None:
        assert data.size == 0 and meta.size == 0
    else:
        assert data.tobytes() == ref.data
        assert meta.tobytes() == ref.meta
    assert st.gpu_kernel_launches > 0
    assert st.path_flags & (pkg.PATH_FUSED_INGEST | pkg.PATH_GENERAL_DECODE) and (st.path_flags & pkg.PATH_ENCODER_V4 or not ekv)
    return job, exp


def runs_to_ssts(runs, block_size=1024):
    return [o.Sst.build(r, o.TableOptions(block_size=block_size)) for r in runs if r]


def test_config1_two_sst_minor_10k(pkg):
    """BASELINE config 1: 2 SSTs x 10k entries, 5k overlapping user keys (mirrors
    rocksdb/db/compaction_job_test.cc:193-231), minor compaction, cutoff = min."""
    base = o.YB_EPOCH_US + 10**9
    runs = []
    seq = 1 << 50
    for f in range(2):
        kvs = []
        for k in range(10000):
            row = f * 5000 + k
            d = dk.doc_key([("%024d" % row)], hash_code=(row * 65536) // 15000, hashed=[])
            uk = dk.sub_doc_key(d, [dk.kcol(1)], micros=base + 1000)
            seq += 1
            kvs.append((o.ikey(uk, seq), b"S" + bytes([(row + j) % 251 + 1 for j in range(255)])))
        runs.append(w.sort_run(kvs))
    ssts = [o.Sst.build(r, o.TableOptions(block_size=32768)) for r in runs]
    job, exp = check(pkg, ssts, block_size=32768, bottommost=False, cutoff_ht=o.HT_MIN, other_min_ht=o.HT_MIN)
    assert job.stats().num_output_records == 15000
    assert job.stats().num_record_drop_hidden == 5000


def test_plain_rocksdb_compaction_job_test_simple(pkg):
    # compaction_job_test.cc:339-347 through real SSTs, no DocDB context
    seq = 0
    runs = []
    for i in range(2):
        c = []
        for k in range(10000):
            seq += 1
            c.append((o.ikey(str(i * 5000 + k).encode(), seq), str(i * 10000 + k).encode()))
        runs.append(w.sort_run(c))
    ssts = runs_to_ssts(runs, 4096)
    job, exp = check(pkg, ssts, retention=False, bottommost=True, last_sequence=seq + 1)
    kv = dict(job.kv_list())
    assert kv[o.ikey(b"9999", 15000)] == b"14999"          # largest user key keeps its seqno
    assert o.ikey(b"0", 0) in kv


@pytest.mark.parametrize("seed", range(8))
def test_randomized_docdb(pkg, seed):
    runs = w.random_docdb_runs(seed, n_runs=1 + seed % 6, n_rows=150 + 40 * seed)
    ssts = runs_to_ssts(runs, 512 if seed % 2 else 2048)
    for kw in w.param_grid():
        check(pkg, ssts, block_size=1024, **kw)


@pytest.mark.parametrize("seed", range(3))
def test_varint_and_decimal_key_components(pkg, seed):
    """DocKeys / subkeys with kVarInt / kDecimal entries (YSQL numeric, YCQL varint / decimal keys;
    primitive_value.cc:1314-1349), with the DocKeyV3 bloom filter keyed by them."""
    runs = w.random_numeric_key_runs(10 + seed, n_runs=2 + seed, n_rows=300)
    ssts = runs_to_ssts(runs, 1024)
    for kw in w.param_grid()[1::3]:
        check(pkg, ssts, block_size=1024, filter_policy=1, filter_block_size=1024, **kw)


def test_golden_first_row_regression(pkg):
    # docdb/docdb-test-wrapper.cc:99-133
    d = dk.doc_key(["mydockey", dk.INT_KEY1])
    e = [(dk.sub_doc_key(d, [], micros=4000), dk.OBJECT), (dk.sub_doc_key(d, [], micros=1000), dk.OBJECT),
         (dk.sub_doc_key(d, ["subkey1"], micros=3000), dk.vstr("value3")),
         (dk.sub_doc_key(d, ["subkey1"], micros=2000), dk.vstr("value2")),
         (dk.sub_doc_key(d, ["subkey1"], micros=1000), dk.vstr("value1"))]
    run = w.sort_run([(o.ikey(k, (1 << 50) + i), v) for i, (k, v) in enumerate(e)])
    job, _ = check(pkg, runs_to_ssts([run]), cutoff_ht=o.ht_from_micros(3500))
    assert [k[:-8] for k, _ in job.kv_list()] == [e[0][0], e[1][0], e[2][0]]


def test_generated_shapes_and_mvcc_heavy(pkg):
    # config-4 shape scaled: 20 versions per key, cutoff above 19 of them
    cfg = o.GenConfig(seed=2, num_rows=5000, cols=1, versions=20, num_files=8, value_len=64)
    ssts = o.Sst.generate_all(cfg, o.TableOptions(block_size=8192))
    cutoff = o.ht_from_micros(cfg.base_micros + 18 * 1000 + 500)
    job, exp = check(pkg, ssts, block_size=8192, cutoff_ht=cutoff)
    assert job.stats().num_output_records == 5000 * 2


def test_config2_shape_scaled_8way(pkg):
    cfg = o.GenConfig(seed=7, num_rows=200000, cols=1, versions=1, num_files=8, value_len=256)
    ssts = o.Sst.generate_all(cfg, o.TableOptions(block_size=32768))
    job, exp = check(pkg, ssts, block_size=32768)
    assert job.stats().num_output_records == 200000
    # the bench shape takes the fused pass (TMA-staged verify + value CRCs + decode) and the CRC-linearity block assembler
    assert job.stats().path_flags & pkg.PATH_FUSED_INGEST and not job.stats().path_flags & pkg.PATH_GENERAL_DECODE


def test_hybrid_time_filter(pkg):
    cfg = o.GenConfig(seed=5, num_rows=800, cols=2, versions=6, num_files=3, value_len=32)
    ssts = o.Sst.generate_all(cfg, o.TableOptions(block_size=2048))
    filt = [o.ht_from_micros(cfg.base_micros + 3500), o.HT_INVALID, o.ht_from_micros(cfg.base_micros + 1500)]
    check(pkg, ssts, block_size=2048, ht_filters=filt, cutoff_ht=o.ht_from_micros(cfg.base_micros + 2500))


def test_edge_cases(pkg):
    # single entry, single file
    d = dk.doc_key(["only"])
    one = [(o.ikey(dk.sub_doc_key(d, [dk.kcol(1)], micros=o.YB_EPOCH_US + 1), 1 << 50), dk.vstr("x"))]
    check(pkg, runs_to_ssts([one]))
    # everything dropped -> empty output, no files
    tomb = [(o.ikey(dk.sub_doc_key(d, [dk.kcol(1)], micros=o.YB_EPOCH_US + 1), 1 << 50), dk.TOMBSTONE)]
    job, _ = check(pkg, runs_to_ssts([tomb]), cutoff_ht=o.ht_from_micros(o.YB_EPOCH_US + 100))
    assert job.stats().num_output_records == 0
    # empty values, long keys (> 128 bytes), ragged sizes
    long_rows = []
    for i in range(300):
        dd = dk.doc_key(["L" * (5 + (i * 7) % 200) + "%05d" % i])
        long_rows.append((o.ikey(dk.sub_doc_key(dd, [dk.kcol(1)], micros=o.YB_EPOCH_US + i), (1 << 50) + i),
                          b"" if i % 3 == 0 else dk.vstr("v" * (i % 50))))
    check(pkg, runs_to_ssts([w.sort_run(long_rows)], 256))


def test_errors_are_loud(pkg):
    # packed rows need the tablet's SchemaPackingProvider: NotSupported, never silently mishandled
    d = dk.doc_key(["r"])
    run = [(o.ikey(dk.sub_doc_key(d, [], micros=o.YB_EPOCH_US + 1), 1 << 50), b"z\x01packed")]
    job = pkg.GpuCompactionJob()
    s = runs_to_ssts([run])[0]
    job.add_input_sst(s.meta_view(), s.data_view())
    with pytest.raises(pkg.YbGpuError) as e:
        job.run()
    assert e.value.status_name == "NotSupported"
    # corrupted block checksum
    s2 = runs_to_ssts([[(o.ikey(b"Sabc\x00\x00!#" + o.encode_doc_ht(o.YB_EPOCH_US), 5), b"Sv")]])[0]
    data = s2.data_view().copy()
    data[-1] ^= 0xff                    # damage the stored CRC, contents stay decodable
    job = pkg.GpuCompactionJob()
    job.add_input_sst(s2.meta_view(), data)
    with pytest.raises(pkg.YbGpuError) as e:
        job.run()                       # input CRC32C is verified on the GPU
    assert e.value.status_name == "Corruption"
    job = pkg.GpuCompactionJob(verify_checksums=False)
    job.add_input_sst(s2.meta_view(), data)
    job.run()                           # like ReadOptions::verify_checksums = false


def test_gpu_block_encoder_cut_rules(pkg):
    """Block cuts (flush_block_policy.cc:45-76) under many block sizes / entry-size mixes, including
    blocks longer than a chain segment and deviation = 0."""
    import random
    rng = random.Random(17)
    rows = []
    for i in range(30000):
        d = dk.doc_key(["k%07d" % i])
        vlen = rng.choice([0, 1, 3, 20, 100, 300, 2000]) if i % 7 else rng.randrange(0, 5000)
        rows.append((o.ikey(dk.sub_doc_key(d, [dk.kcol(1 + i % 3)], micros=o.YB_EPOCH_US + i), (1 << 50) + i),
                     b"S" + bytes(rng.randrange(256) for _ in range(vlen))))
    run = w.sort_run(rows)
    sst = o.Sst.build(run, o.TableOptions(block_size=4096))
    for bs, dev, ri in ((256, 10, 16), (4096, 10, 16), (32768, 10, 16), (1 << 20, 10, 16), (4096, 0, 16), (4096, 10, 4), (2048, 50, 1)):
        exp = o.compact([sst], o.CompactionParams(), o.TableOptions(block_size=bs, deviation=dev, restart=ri))
        job = pkg.GpuCompactionJob(block_size=bs, deviation=dev, restart_interval=ri)
        job.add_input_sst(sst.meta_view(), sst.data_view())
        job.run()
        data, meta = job.fetch_output()
        assert data.tobytes() == exp.sst().data, (bs, dev, ri)
        assert meta.tobytes() == exp.sst().meta, (bs, dev, ri)
    # blocks far longer than a chain segment (4096 entries): tiny entries, 1 MB blocks
    tiny = w.sort_run([(o.ikey(dk.sub_doc_key(dk.doc_key(["t%06d" % i]), [], micros=o.YB_EPOCH_US + 5), (1 << 50) + i), b"")
                       for i in range(40000)])
    sst = o.Sst.build(tiny, o.TableOptions(block_size=4096))
    for bs in (1 << 20, 1 << 17):
        exp = o.compact([sst], o.CompactionParams(), o.TableOptions(block_size=bs))
        job = pkg.GpuCompactionJob(block_size=bs)
        job.add_input_sst(sst.meta_view(), sst.data_view())
        job.run()
        data, meta = job.fetch_output()
        assert data.tobytes() == exp.sst().data and meta.tobytes() == exp.sst().meta


def test_three_shared_parts_inputs(pkg):
    """kKeyDeltaEncodingThreeSharedParts (YCQL's data-block encoding, block_builder.cc:248-333) inputs,
    mixed with shared-prefix inputs in one job."""
    cfg = o.GenConfig(seed=19, num_rows=6000, cols=3, versions=4, num_files=4, value_len=48, tombstone_per_1024=50)
    ssts = [o.Sst.generate(cfg, f, o.TableOptions(block_size=2048, key_encoding=2 if f != 2 else 1)) for f in range(4)]
    assert [s.key_encoding for s in ssts] == [2, 2, 1, 2]
    job, _ = check(pkg, ssts, block_size=4096, cutoff_ht=o.ht_from_micros(cfg.base_micros + 2500))
    assert job.stats().path_flags & pkg.PATH_GENERAL_DECODE      # three_shared_parts inputs: the general kernels
    runs = w.random_docdb_runs(4, n_runs=3, n_rows=200)
    ssts = [o.Sst.build(r, o.TableOptions(block_size=512, key_encoding=2)) for r in runs]
    for kw in w.param_grid()[:4]:
        check(pkg, ssts, block_size=1024, **kw)


@pytest.mark.parametrize("seed", range(4))
def test_three_shared_parts_output(pkg, seed):
    """Output encoded as kKeyDeltaEncodingThreeSharedParts on the GPU (ThreeSharedPartsEncoder,
    block_builder.cc:265-333: shared prefix + shared middle + reused / incremented last component),
    byte-identical files; inputs in either encoding."""
    if seed == 0:
        # config-2 shaped rows: consecutive columns of a row differ in one byte + the sequence number
        cfg = o.GenConfig(seed=23, num_rows=3000, cols=4, versions=3, num_files=3, value_len=40, tombstone_per_1024=30)
        ssts = [o.Sst.generate(cfg, f, o.TableOptions(block_size=2048, key_encoding=1 + f % 2)) for f in range(3)]
        check(pkg, ssts, block_size=4096, output_key_encoding=2)
        check(pkg, ssts, block_size=32768, output_key_encoding=2, cutoff_ht=o.ht_from_micros(cfg.base_micros + 1500))
    else:
        runs = w.random_docdb_runs(40 + seed, n_runs=1 + seed, n_rows=120 + 50 * seed)
        ssts = [o.Sst.build(r, o.TableOptions(block_size=1024, key_encoding=1 + (i + seed) % 2)) for i, r in enumerate(runs) if r]
        for kw in w.param_grid()[:5]:
            check(pkg, ssts, block_size=1024 if seed % 2 else 4096, output_key_encoding=2, **kw)
    # plain RocksDB keys (no DocDB structure): sequence numbers zeroed at the bottommost level make
    # the last component reusable, increasing ones exercise the "+0x100" path
    seq = 0
    runs = []
    for i in range(2):
        c = []
        for k in range(3000):
            seq += 1
            c.append((o.ikey(b"user%06d" % (i * 1500 + k), seq), b"v" * (k % 7)))
        runs.append(w.sort_run(c))
    ssts = runs_to_ssts(runs, 4096)
    check(pkg, ssts, block_size=2048, retention=False, bottommost=bool(seed % 2), last_sequence=seq + 1, output_key_encoding=2)


@pytest.mark.parametrize("seed", range(4))
def test_bloom_filter_blocks(pkg, seed):
    """DocKeyV3Filter fixed-size bloom filter blocks built on the GPU (hash, bit setting, block cuts every
    max_keys distinct filter keys) and placed by the host among the index blocks: metadata file
    byte-identical to the oracle's BlockBasedTableBuilder restatement."""
    if seed == 0:
        cfg = o.GenConfig(seed=31, num_rows=6000, cols=2, versions=3, num_files=4, value_len=40, tombstone_per_1024=30)
        ssts = o.Sst.generate_all(cfg, o.TableOptions(block_size=4096))
        for fbs in (256, 2048, 65536):
            check(pkg, ssts, block_size=2048, filter_policy=1, filter_block_size=fbs)
        check(pkg, ssts, block_size=32768, filter_policy=1, filter_block_size=512, cutoff_ht=o.ht_from_micros(cfg.base_micros + 1500))
    elif seed == 1:
        # cotables / colocated tables: table tombstones and rows share the id prefix; range-only keys
        runs = w.random_cotable_runs(5, n_runs=3, n_tables=5, rows_per_table=80, colocated=True)
        ssts = runs_to_ssts(runs, 512)
        for kw in w.param_grid()[:3]:
            check(pkg, ssts, block_size=1024, filter_policy=1, filter_block_size=128, **kw)
    else:
        runs = w.random_docdb_runs(60 + seed, n_runs=2 + seed, n_rows=200 + 100 * seed)
        ssts = runs_to_ssts(runs, 1024)
        for kw in w.param_grid()[:4]:
            check(pkg, ssts, block_size=1024, filter_policy=1, filter_block_size=128 * seed, output_key_encoding=seed - 1, **kw)
    # keys that are not DocKeys never enter the filter; an empty filter block is still written
    seq = 0
    c = []
    for k in range(500):
        seq += 1
        c.append((o.ikey(b"\x7fplain%05d" % k, seq), b"v%d" % k))
    check(pkg, runs_to_ssts([w.sort_run(c)], 1024), block_size=1024, retention=False, filter_policy=1, filter_block_size=256)


@pytest.mark.parametrize("seed", range(6))
def test_cotables_and_colocated_tables(pkg, seed):
    """Table tombstones (id ! # HT) shadow every older row of their table across row groups and
    tiles (slot 0 of the overwrite stack, docdb_compaction_context.cc:999-1024)."""
    runs = w.random_cotable_runs(seed, n_runs=1 + seed % 4, n_tables=6, rows_per_table=60, colocated=seed % 2 == 0)
    ssts = runs_to_ssts(runs, 512)
    for kw in w.param_grid()[:6]:
        check(pkg, ssts, block_size=1024, **kw)
    if seed % 2:
        check(pkg, ssts, block_size=1024, bottommost=True, cutoff_ht=o.ht_from_micros(w.BASE_US + 35),
              cotables_cutoff_ht=o.ht_from_micros(w.BASE_US + 85))


@pytest.mark.parametrize("enc,in_flight", [(1, 3), (2, 1)])
def test_subcompactions_pipelined(pkg, enc, in_flight):
    """ybgpu_compact_files: one compaction cut into key ranges on row boundaries, ranges pipelined on
    private streams (compaction_job.cc:409-519,532-552). Every range's output SST is byte-identical to
    the oracle's compaction of the same key range, the outputs concatenate to the single-job KV stream,
    the seqno-zeroing exception key is the compaction's (not the range's) largest user key, and
    FileMetaData boundaries come back per output."""
    cfg = o.GenConfig(seed=23, num_rows=9000, cols=2, versions=4, num_files=6, value_len=90, tombstone_per_1024=50)
    ssts = o.Sst.generate_all(cfg, o.TableOptions(block_size=4096))
    cutoff = o.ht_from_micros(cfg.base_micros + 2500)
    topt = o.TableOptions(block_size=4096, key_encoding=enc, filter_policy=1, filter_block_size=4096)
    exp = o.compact(ssts, o.CompactionParams(cutoff_ht=cutoff), topt)
    files = [(s.meta_view(), s.data_view()) for s in ssts]
    res = pkg.compact_files(files, max_subcompactions=5, max_in_flight=in_flight, cutoff_ht=cutoff, block_size=4096,
                            output_key_encoding=enc, filter_policy=1, filter_block_size=4096)
    assert 3 <= len(res.outputs) <= 5
    all_kvs = [s.read_all() for s in ssts]
    largest = max(kvs[-1][0][:-8] for kvs in all_kvs)
    got = []
    for out in res.outputs:
        lo, hi = out.lower, out.upper
        part = [[kv for kv in kvs if (not lo or kv[0][:-8] >= lo) and (not hi or kv[0][:-8] < hi)] for kvs in all_kvs]
        part_ssts = [o.Sst.build(p, o.TableOptions(block_size=4096)) for p in part if p]
        ref = o.compact(part_ssts, o.CompactionParams(cutoff_ht=cutoff, largest_user_key=largest), topt)
        assert out.stats.num_input_records == ref.stats.num_input_records
        assert out.stats.num_output_records == ref.stats.num_output_records
        rs = ref.sst()
        if rs is None:
            assert out.data_len == 0
            continue
        data = res.data_arena[out.data_offset:out.data_offset + out.data_len].tobytes()
        meta = res.meta_arena[out.meta_offset:out.meta_offset + out.meta_len].tobytes()
        assert data == rs.data and meta == rs.meta
        kvs = ref.kv_list()
        assert (out.smallest, out.largest) == (kvs[0][0], kvs[-1][0])
        got += kvs
    assert got == exp.kv_list()
    assert res.total.num_input_records == exp.stats.num_input_records
    assert res.total.num_output_records == exp.stats.num_output_records
    assert [o_.lower for o_ in res.outputs[1:]] == [o_.upper for o_ in res.outputs[:-1]]


@pytest.mark.parametrize("seed", range(6))
def test_subcompactions_inside_cotables(pkg, seed):
    """Ranges that start inside a cotable / colocated table still see the table's tombstones (`id ! # HT`), which
    sort before the range: their blocks are loaded out of range and seed slot 0 of the overwrite stack
    (docdb_compaction_context.cc:999-1024). The range outputs must concatenate to the single-pass oracle output."""
    runs = w.random_cotable_runs(100 + seed, n_runs=2 + seed % 3, n_tables=3, rows_per_table=150, colocated=seed % 2 == 0)
    ssts = runs_to_ssts(runs, 512)
    files = [(s.meta_view(), s.data_view()) for s in ssts]
    inside = 0
    for kw in [w.param_grid()[i] for i in (0, 2, 4, 6)]:
        exp = o.compact(ssts, o.CompactionParams(**okw(kw)), o.TableOptions(block_size=1024))
        res = pkg.compact_files(files, max_subcompactions=7, max_in_flight=2, block_size=1024, **kw)
        assert len(res.outputs) >= 4
        got = []
        for out in res.outputs:
            if out.lower and out.lower[:1] in (b"y", b"0") and len(out.lower) > (17 if out.lower[:1] == b"y" else 5) + 1:
                inside += 1
            if out.data_len:
                got += o.Sst.from_bytes(res.meta_arena[out.meta_offset:out.meta_offset + out.meta_len].tobytes(),
                                        res.data_arena[out.data_offset:out.data_offset + out.data_len].tobytes()).read_all()
        assert got == exp.kv_list()
        assert res.total.num_input_records == exp.stats.num_input_records
        assert res.total.num_output_records == exp.stats.num_output_records
        assert res.total.num_record_drop_feed == exp.stats.num_dropped_feed
    assert inside >= 4, "the test must cut inside tables"


@pytest.mark.parametrize("seed,cols,versions,collection,colocated", [(0, 250, 20, 2000, False), (1, 120, 45, 1500, True), (2, 500, 100, 4000, False)])
def test_rows_larger_than_a_merge_tile(pkg, seed, cols, versions, collection, colocated):
    """a9 without a row-size limit: DocDBCompactionFeed carries its overwrite stack over any number of entries of one
    DocKey (docdb_compaction_context.cc:999-1024). Rows of ~5 000 and ~50 000 entries (many columns x versions, a
    collection under a column with its own tombstones) are cut across merge tiles; tiles that start inside the row
    rebuild the state by replaying the ancestors of their first key. Full-file parity with the oracle."""
    runs = w.giant_row_runs(seed, n_runs=2 + seed, cols=cols, versions=versions, collection=collection, colocated=colocated)
    assert sum(len(r) for r in runs) > (40000 if seed == 2 else 5000)
    ssts = runs_to_ssts(runs, 4096)
    grid = w.param_grid()
    for kw in ([grid[i] for i in (0, 2, 3, 4, 6, 8, 9)] if seed < 2 else [grid[2], grid[4]]):
        job, _ = check(pkg, ssts, block_size=4096, **kw)
        assert job.stats().tiles_inside_rows > 0            # the rows really were cut across tiles


def test_plain_mode_user_key_with_thousands_of_versions(pkg):
    """retention off (plain RocksDB): a row group is one user key; 6 000 versions of one key span many tiles and
    rule A (compaction_iterator.cc:388-400) must still keep exactly the newest one."""
    seq = 1 << 40
    runs = [[], [], []]
    for i in range(6000):
        seq += 1
        runs[i % 3].append((o.ikey(b"hot-key", seq), b"v%d" % i))
    for j in range(300):
        seq += 1
        runs[j % 3].append((o.ikey(b"key%04d" % j, seq), b"x" * (j % 50)))
    ssts = runs_to_ssts([w.sort_run(r) for r in runs], 2048)
    job, exp = check(pkg, ssts, block_size=2048, retention=False, bottommost=True)
    assert job.stats().num_output_records == 301
    check(pkg, ssts, block_size=2048, retention=False, bottommost=False)


def test_emit_kv_stream_is_the_kv_list(pkg):
    """a8: ybgpu_job_emit_kv_stream hands the surviving stream to a CompactionFeed-shaped callback in output order;
    a non-zero return aborts with that status (compaction_job.cc:797-800)."""
    cfg = o.GenConfig(seed=5, num_rows=2000, cols=2, versions=3, num_files=3, value_len=60, tombstone_per_1024=60)
    ssts = o.Sst.generate_all(cfg, o.TableOptions(block_size=4096))
    cutoff = o.ht_from_micros(cfg.base_micros + 1500)
    exp = o.compact(ssts, o.CompactionParams(cutoff_ht=cutoff), o.TableOptions(block_size=4096))
    job = gpu_compact(pkg, ssts, cutoff_ht=cutoff, block_size=4096)
    got = []
    job.emit_kv_stream(lambda k, v: got.append((k, v)) or 0)
    assert got == exp.kv_list() == job.kv_list()
    seen = []

    def failing(k, v):
        if len(seen) == 100:
            return 5            # IOError
        seen.append(k)
        return 0
    with pytest.raises(pkg.YbGpuError) as e:
        job.emit_kv_stream(failing)
    assert e.value.status == 5 and len(seen) == 100


def test_concurrent_jobs_on_private_streams(pkg):
    """Jobs of different tablets run concurrently on one device (one PriorityThreadPool worker per
    CompactionJob, db_impl.cc:397-403): each on its own stream, results unchanged."""
    import threading
    cfgs = [o.GenConfig(seed=40 + i, num_rows=3000 + 500 * i, cols=2, versions=3, num_files=3, value_len=70, tombstone_per_1024=30) for i in range(4)]
    tablets = [o.Sst.generate_all(c, o.TableOptions(block_size=4096)) for c in cfgs]
    cutoffs = [o.ht_from_micros(c.base_micros + 1500) for c in cfgs]
    exps = [o.compact(t, o.CompactionParams(cutoff_ht=c), o.TableOptions(block_size=4096)) for t, c in zip(tablets, cutoffs)]
    results = [None] * len(tablets)

    def work(i):
        for _ in range(3):
            job = gpu_compact(pkg, tablets[i], cutoff_ht=cutoffs[i], block_size=4096, cuda_stream=pkg.STREAM_PRIVATE)
            data, meta = job.fetch_output()
            results[i] = (data.tobytes(), meta.tobytes(), job.stats().num_input_records)
            job.close()

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(tablets))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for r, e in zip(results, exps):
        assert r is not None and r[0] == e.sst().data and r[1] == e.sst().meta and r[2] == e.stats.num_input_records


def test_reference_golden_dumps_on_gpu(pkg):
    """Every history compaction of tests/test_reference_dumps.py (the reference's own golden dumps:
    docdb-test-wrapper.cc BasicTest / StaticColumnCompaction / user timestamps / cotables, docdb-ttl-test.cc
    TTL sequences) replayed through the CUDA path: surviving set = the reference's expected dump, and
    files / stats = the oracle's."""
    import test_reference_dumps as t
    replayed = []

    def on_gpu(runs, kw, want):
        job, _ = check(pkg, runs_to_ssts(runs), block_size=512, **kw)
        assert [(k[:-8], v) for k, v in job.kv_list()] == want
        replayed.append(len(want))

    t.EXTRA_CHECK = on_gpu
    try:
        for name in sorted(dir(t)):
            if name.startswith("test_") and "parser" not in name:
                getattr(t, name)()
    finally:
        t.EXTRA_CHECK = None
    assert len(replayed) >= 20


def test_subcompaction_outputs_concatenate_into_one_table(pkg):
    """ybgpu_compact_files + ybgpu_sst_concat_meta: the GPU's range outputs, appended in range order under one
    rebuilt index / filter index, are a single table holding exactly the single-job KV stream."""
    cfg = o.GenConfig(seed=37, num_rows=8000, cols=2, versions=3, num_files=5, value_len=80, tombstone_per_1024=40)
    ssts = o.Sst.generate_all(cfg, o.TableOptions(block_size=4096))
    cutoff = o.ht_from_micros(cfg.base_micros + 1500)
    topt = o.TableOptions(block_size=4096, filter_policy=1, filter_block_size=4096)
    exp = o.compact(ssts, o.CompactionParams(cutoff_ht=cutoff), topt)
    res = pkg.compact_files([(s.meta_view(), s.data_view()) for s in ssts], max_subcompactions=6, max_in_flight=3,
                            cutoff_ht=cutoff, block_size=4096, filter_policy=1, filter_block_size=4096)
    outs = [x for x in res.outputs if x.data_len]
    assert len(outs) >= 3
    pieces = [(res.meta_arena[x.meta_offset:x.meta_offset + x.meta_len], x.data_len, x.smallest, x.largest) for x in outs]
    meta = pkg.sst_concat_meta(pieces, block_size=4096, filter_policy=1, filter_block_size=4096)
    data = b"".join(res.data_arena[x.data_offset:x.data_offset + x.data_len].tobytes() for x in outs)
    whole = o.Sst.from_bytes(meta, data)
    assert whole.read_all() == exp.kv_list()
    assert o.varint(whole.properties()["rocksdb.num.entries"]) == exp.stats.num_output_records
    # and the table is a valid INPUT of the next compaction
    job = gpu_compact(pkg, [whole], cutoff_ht=cutoff, block_size=4096)
    assert job.kv_list() == exp.kv_list()


@pytest.mark.parametrize("enc,in_flight,nsub", [(1, 3, 6), (2, 2, 9), (1, 1, 3)])
def test_compact_files_one_table(pkg, enc, in_flight, nsub):
    """ybgpu_compact_files_one_table: the pipelined key ranges land back to back in ONE data file and one metadata
    file is assembled while later ranges run. The table must (i) hold exactly the single-job KV stream, (ii) equal,
    byte for byte, what ybgpu_sst_concat_meta builds from the per-range outputs of ybgpu_compact_files, (iii) carry
    the compaction's FileMetaData boundaries, (iv) be a valid input of the next compaction."""
    cfg = o.GenConfig(seed=51, num_rows=9000, cols=2, versions=3, num_files=5, value_len=80, tombstone_per_1024=40)
    ssts = o.Sst.generate_all(cfg, o.TableOptions(block_size=4096))
    cutoff = o.ht_from_micros(cfg.base_micros + 1500)
    kw = dict(cutoff_ht=cutoff, block_size=4096, output_key_encoding=enc, filter_policy=1, filter_block_size=4096)
    exp = o.compact(ssts, o.CompactionParams(cutoff_ht=cutoff), o.TableOptions(block_size=4096, key_encoding=enc, filter_policy=1, filter_block_size=4096))
    files = [(s.meta_view(), s.data_view()) for s in ssts]
    data, meta, res, total = pkg.compact_files_one_table(files, max_subcompactions=nsub, max_in_flight=in_flight, **kw)
    assert res.num_pieces >= 2 and res.num_ranges >= res.num_pieces
    whole = o.Sst.from_bytes(meta.tobytes(), data.tobytes())
    ekv = exp.kv_list()
    assert whole.read_all() == ekv
    assert (res.smallest, res.largest) == (ekv[0][0], ekv[-1][0])
    assert total.num_input_records == exp.stats.num_input_records and total.num_output_records == exp.stats.num_output_records
    assert total.output_data_file_size == data.size and total.output_meta_file_size == meta.size
    ranges = pkg.compact_files(files, max_subcompactions=nsub, max_in_flight=in_flight, **kw)
    outs = [x for x in ranges.outputs if x.data_len]
    pieces = [(ranges.meta_arena[x.meta_offset:x.meta_offset + x.meta_len], x.data_len, x.smallest, x.largest) for x in outs]
    assert meta.tobytes() == bytes(pkg.sst_concat_meta(pieces, block_size=4096, output_key_encoding=enc, filter_policy=1, filter_block_size=4096))
    assert data.tobytes() == b"".join(ranges.data_arena[x.data_offset:x.data_offset + x.data_len].tobytes() for x in outs)
    job = gpu_compact(pkg, [whole], cutoff_ht=cutoff, block_size=4096)
    assert job.kv_list() == ekv


@pytest.mark.parametrize("seed", range(5))
def test_user_boundary_values(pkg, seed):
    """a19: FileMetaData user boundary values. DocDBCompactionFeed::UpdateBoundaryValues keeps, per range component of
    the DocKeys it passes on, the bytewise smallest / largest encoded value (docdb_compaction_context.cc:754-773,
    doc_boundary_values_extractor.cc:40-64); the GPU reduces them over the first surviving entry of every row."""
    if seed < 3:
        runs = w.random_docdb_runs(500 + seed, n_runs=2 + seed, n_rows=120)
    elif seed == 3:
        runs = w.random_cotable_runs(510, n_runs=3, n_tables=4, rows_per_table=40, colocated=True)
    else:
        runs = w.random_numeric_key_runs(520, n_runs=3, n_rows=80)
    ssts = runs_to_ssts(runs, 1024)
    for kw in [w.param_grid()[i] for i in (0, 2, 4, 6)]:
        exp = o.compact(ssts, o.CompactionParams(**okw(kw)), o.TableOptions(block_size=1024))
        job = gpu_compact(pkg, ssts, block_size=1024, user_boundary_values=True, **kw)
        assert job.kv_list() == exp.kv_list()
        assert job.user_values() == exp.user_values(), kw
    # rows larger than a merge tile: the first surviving entry of a row may be met in several tiles — same extrema
    big = w.giant_row_runs(7, n_runs=3, cols=150, versions=12, collection=800)
    ssts = runs_to_ssts(big, 4096)
    if seed == 0:
        kw = w.param_grid()[6]
        exp = o.compact(ssts, o.CompactionParams(**okw(kw)), o.TableOptions(block_size=4096))
        job = gpu_compact(pkg, ssts, block_size=4096, user_boundary_values=True, **kw)
        assert job.user_values() == exp.user_values()


@pytest.mark.parametrize("seed", range(4))
def test_snappy_compressed_inputs(pkg, seed):
    """f2: Snappy-compressed input data blocks (DocDB's production default, docdb_rocksdb_util.cc:184). The stored bytes'
    checksum is verified, the blocks are uncompressed on the GPU (table/format.cc:441-500) and the compaction proceeds as
    for raw inputs: output files identical to the oracle's, which reads the same compressed inputs. Files mix compressed
    and raw blocks (a block is stored compressed only if that saves 12.5 %, block_based_table_builder.cc:109-131)."""
    if seed < 2:
        runs = w.random_docdb_runs(600 + seed, n_runs=3, n_rows=150 + 100 * seed)
        kws = [w.param_grid()[i] for i in (0, 2, 6)]
    else:
        cfg = o.GenConfig(seed=60 + seed, num_rows=4000, cols=2, versions=3, num_files=3, value_len=24 if seed == 2 else 200, tombstone_per_1024=50)
        runs = [s.read_all() for s in o.Sst.generate_all(cfg, o.TableOptions(block_size=4096))]
        if seed == 3:
            # random 200-byte strings do not compress; every other stretch of 150 entries gets a repetitive string, so
            # the files hold compressed and raw blocks side by side
            runs = [[(k, v[:1] + (k[-12:-8] * 50)[:len(v) - 1]) if v[:1] == b"S" and (i // 150) % 2 else (k, v) for i, (k, v) in enumerate(r)] for r in runs]
        kws = [dict(cutoff_ht=o.ht_from_micros(cfg.base_micros + 1500))]
    ssts = [o.Sst.build(r, o.TableOptions(block_size=2048 if seed < 2 else 8192, compression=1)) for r in runs if r]
    raw = sum(sum(len(k) + len(v) for k, v in r) for r in runs)
    assert sum(len(s.data) for s in ssts) < raw                      # something was compressed
    def stored_compressed(s):                                       # any block whose trailer says kSnappyCompression
        offs, sizes = s.block_handles()
        d = bytes(s.data)
        return any(d[int(off) + int(size)] == 1 for off, size in zip(offs, sizes))
    def stored_raw(s):
        offs, sizes = s.block_handles()
        d = bytes(s.data)
        return any(d[int(off) + int(size)] == 0 for off, size in zip(offs, sizes))
    any_compressed = any(stored_compressed(s) for s in ssts)
    assert any_compressed and (seed != 3 or all(stored_raw(s) for s in ssts))
    for kw in kws:
        job, _ = check(pkg, ssts, block_size=4096, filter_policy=1, filter_block_size=4096, **kw)
        assert bool(job.stats().path_flags & pkg.PATH_SNAPPY) == any_compressed
    # pipelined key ranges over the same production-shaped inputs (compressed index blocks are read by the host planner, the
    # last data block by the last-key helper): the ranges' tables hold the single job's KV stream
    exp = o.compact(ssts, o.CompactionParams(**kws[0]), o.TableOptions(block_size=4096))
    res = pkg.compact_files([(s.meta_view(), s.data_view()) for s in ssts], max_subcompactions=3, max_in_flight=2, block_size=4096, **kws[0])
    got = []
    for data, meta in res.files():
        got += o.Sst.from_bytes(meta.tobytes(), data.tobytes()).read_all()
    assert got == exp.kv_list()
    # a flipped bit in a compressed block is a checksum error, not garbage
    bad = bytearray(ssts[0].data)
    bad[len(bad) // 2] ^= 0x10
    job = pkg.GpuCompactionJob(block_size=4096)
    job.add_input_sst(ssts[0].meta_view(), np.frombuffer(bytes(bad), np.uint8))
    with pytest.raises(pkg.YbGpuError) as e:
        job.run()
    assert e.value.status_name == "Corruption"


def _phrase_runs(seed, n_runs, n_rows, vmax=160, random_every=None):
    """DocDB-shaped runs whose values are made of recurring phrases (compressible); with random_every, every other
    stretch of that many rows carries random bytes instead, so that tables hold blocks worth compressing next to
    blocks that are not."""
    import random as _random
    rng = _random.Random(seed)
    runs = w.random_docdb_runs(seed, n_runs=n_runs, n_rows=n_rows)
    words = [bytes(rng.randrange(32, 127) for _ in range(rng.randrange(2, 20))) for _ in range(30)]
    out = []
    for r in runs:
        rr = []
        for i, (k, v) in enumerate(r):
            if v[:1] == b"S":                                      # string values only: control fields / tombstones keep their bytes
                if random_every and (i // random_every) % 2:
                    v = b"S" + bytes(rng.randrange(256) for _ in range(rng.randrange(vmax)))
                else:
                    v = b"S" + b" ".join(rng.choice(words) for _ in range(rng.randrange(vmax // 8)))
            rr.append((k, v))
        out.append(rr)
    return out


def _stored_types(data, meta, pkg):
    off, sz, _ = pkg.sst_block_handles(np.frombuffer(meta, np.uint8))
    return [data[int(a) + int(b)] for a, b in zip(off, sz)], off, sz


@pytest.mark.parametrize("seed", range(4))
def test_snappy_compressed_output(pkg, seed):
    """a17: BlockBasedTableBuilder::WriteBlock -> CompressBlock with kSnappyCompression, DocDB's production setting
    (block_based_table_builder.cc:109-131,630-655; docdb_rocksdb_util.cc:184). Every assembled data block goes through
    the GPU encoder (k_snappy_compress) and is stored compressed when that saves 12.5 %, its checksum then covers the
    compressed bytes; index blocks and the filter index likewise (host). The oracle's builder uses the same encoder, so
    whole files are compared byte for byte; the block CONTENTS are pinned independently: each stored block, decoded by
    the real snappy library (pyarrow) when it is there and by the oracle's decoder otherwise, is the block the
    uncompressed twin of the table holds at that index."""
    if seed == 0:
        runs, bs, enc, kws = _phrase_runs(900, 3, 300), 1024, 1, [w.param_grid()[i] for i in (0, 2, 6)]
    elif seed == 1:
        runs, bs, enc, kws = _phrase_runs(901, 4, 1500, random_every=120), 4096, 2, [w.param_grid()[0]]
    elif seed == 2:
        runs, bs, enc, kws = _phrase_runs(902, 3, 4000, vmax=400, random_every=700), 32768, 1, [w.param_grid()[2]]
    else:
        # values beyond a fragment (64 KB): blocks of several fragments, long runs (copy elements of 64 bytes), tiny blocks
        runs = _phrase_runs(903, 2, 60)
        big = [(k, b"S" + (bytes(range(256)) * 700)[:150000 + 7 * i] if i % 9 == 4 and v[:1] == b"S" else
                (b"S" + b"\0" * (3000 + i) if i % 9 == 7 and v[:1] == b"S" else v)) for i, (k, v) in enumerate(runs[0])]
        runs, bs, enc, kws = [big, runs[1]], 2048, 1, [w.param_grid()[0]]
    # production shape on both sides: the inputs are Snappy tables too
    ssts = [o.Sst.build(r, o.TableOptions(block_size=bs, compression=1)) for r in runs if r]
    try:
        import pyarrow as pa
        lib_ok = pa.Codec.is_available("snappy")
    except ImportError:
        lib_ok = False
    for kw in kws:
        job, exp = check(pkg, ssts, block_size=bs, output_key_encoding=enc, filter_policy=1, filter_block_size=4096, output_compression=1, **kw)
        assert job.stats().path_flags & pkg.PATH_SNAPPY_OUTPUT
        data, meta = job.fetch_output()
        data, meta = data.tobytes(), meta.tobytes()
        types, off, sz = _stored_types(data, meta, pkg)
        plain = gpu_compact(pkg, ssts, block_size=bs, output_key_encoding=enc, filter_policy=1, filter_block_size=4096, **kw)
        assert not plain.stats().path_flags & pkg.PATH_SNAPPY_OUTPUT
        pdata, pmeta = plain.fetch_output()
        pdata, pmeta = pdata.tobytes(), pmeta.tobytes()
        ptypes, poff, psz = _stored_types(pdata, pmeta, pkg)
        assert len(types) == len(ptypes) and set(ptypes) == {0} and 1 in types
        assert len(data) < len(pdata)
        if seed in (1, 2):
            assert 0 in types                                      # the random stretches were not worth it
        for t, a, b, pa_, pb in zip(types, off, sz, poff, psz):
            stored = data[int(a):int(a) + int(b)]
            want = pdata[int(pa_):int(pa_) + int(pb)]
            if t == 0:
                assert stored == want
            else:
                assert len(stored) < len(want) - len(want) // 8
                assert o.snappy_uncompress(stored) == want
                if lib_ok:
                    assert pa.decompress(stored, decompressed_size=len(want), codec="snappy").to_pybytes() == want
        assert pkg.sst_verify_blocks(np.frombuffer(meta, np.uint8), np.frombuffer(data, np.uint8)) == (len(types), 0)
    # the compressed table is a valid input of the next compaction, and pipelined key ranges write compressed pieces that
    # assemble into one table with the single job's entries
    kw = kws[0]
    exp = o.compact(ssts, o.CompactionParams(**okw(kw)), o.TableOptions(block_size=bs, key_encoding=enc))
    again = gpu_compact(pkg, [o.Sst.from_bytes(meta, data)], block_size=bs, **kw)
    assert again.kv_list() == o.compact([o.Sst.from_bytes(meta, data)], o.CompactionParams(**okw(kw)), o.TableOptions(block_size=bs)).kv_list()
    files = [(s.meta_view(), s.data_view()) for s in ssts]
    d1, m1, res, total = pkg.compact_files_one_table(files, max_subcompactions=4, max_in_flight=2, block_size=bs, output_key_encoding=enc,
                                                     filter_policy=1, filter_block_size=4096, output_compression=1, **kw)
    whole = o.Sst.from_bytes(m1.tobytes(), d1.tobytes())
    assert whole.read_all() == exp.kv_list()
    assert total.path_flags & pkg.PATH_SNAPPY_OUTPUT and 1 in _stored_types(d1.tobytes(), m1.tobytes(), pkg)[0]


def test_yield_points(pkg):
    """The scheduler's pause hook (PriorityThreadPoolSuspender::PauseIfNecessary in the reference) is honoured
    between kernel phases of a job and between the ranges of a compaction run as subcompactions."""
    cfg = o.GenConfig(seed=3, num_rows=3000, cols=2, versions=2, num_files=3, value_len=40)
    ssts = o.Sst.generate_all(cfg, o.TableOptions(block_size=4096))
    calls = []
    job = gpu_compact(pkg, ssts, block_size=4096, yield_fn=lambda: calls.append(1))
    assert len(calls) >= 3 and job.stats().num_input_records == 12000
    calls.clear()
    res = pkg.compact_files([(s.meta_view(), s.data_view()) for s in ssts], max_subcompactions=4, max_in_flight=2,
                            block_size=4096, yield_fn=lambda: calls.append(1))
    assert len(calls) >= len(res.outputs) and res.total.num_input_records == 12000


@pytest.mark.parametrize("assembler", ["v5", "v4"])
def test_block_assemblers(pkg, assembler, monkeypatch):
    """Both block assemblers produce the oracle's bytes: k_encode_v5 (warp per block, the default) and k_encode_v4
    (CTA per block with a shared-memory image; what runs when a scratch row per lane does not fit, and the A/B switch
    YBGPU_ENC_V4) — TTL rewrites and tombstoned values (the entries whose bytes are not copied from the inputs), both
    key encodings, tiny and large blocks, empty values."""
    if assembler == "v4":
        monkeypatch.setenv("YBGPU_ENC_V4", "1")
    want = pkg.PATH_ENCODER_V5 if assembler == "v5" else 0
    runs = w.random_docdb_runs(77, n_runs=4, n_rows=400)
    ssts = runs_to_ssts(runs, 1024)
    for kw in w.param_grid():
        for enc in (1, 2):
            job, _ = check(pkg, ssts, block_size=512 if enc == 1 else 8192, output_key_encoding=enc, **kw)
            assert (job.stats().path_flags & pkg.PATH_ENCODER_V5) == want
    cfg = o.GenConfig(seed=5, num_rows=4000, cols=2, versions=3, num_files=3, value_len=700, tombstone_per_1024=100)
    ssts = o.Sst.generate_all(cfg, o.TableOptions(block_size=8192))
    for bs in (300, 4096, 65536, 1 << 20):
        check(pkg, ssts, block_size=bs, cutoff_ht=o.ht_from_micros(cfg.base_micros + 1500), filter_policy=1)
    cfg = o.GenConfig(seed=6, num_rows=3000, cols=1, versions=2, num_files=2, value_len=0)
    check(pkg, o.Sst.generate_all(cfg, o.TableOptions(block_size=2048)), block_size=2048)


def test_small_upload_ring_wraps(pkg, monkeypatch):
    """Parameter uploads go through a ring inside the job's host-mapped page (no DMA on the critical path); with a
    tiny ring every few uploads wrap (stream synchronise, start over) — many input files make the run table the
    largest upload."""
    monkeypatch.setenv("YBGPU_UPLOAD_RING_BYTES", "2048")
    runs = w.random_docdb_runs(5, n_runs=6, n_rows=300)
    ssts = runs_to_ssts(runs, 1024)
    for kw in w.param_grid()[:3]:
        check(pkg, ssts, block_size=1024, filter_policy=1, filter_block_size=2048, **kw)
    cfg = o.GenConfig(seed=9, num_rows=5000, cols=2, versions=2, num_files=24, value_len=60)
    check(pkg, o.Sst.generate_all(cfg, o.TableOptions(block_size=2048)), block_size=4096)


def test_kv_stream_inputs_flush_path(pkg):
    """ybgpu_job_add_input_kv: sorted runs held in memory instead of table files — the flush path's shape (BuildTable,
    rocksdb/db/builder.cc:119-318: memtable iterator -> CompactionIterator -> TableBuilder). The job must write exactly
    what a compaction of table files with the same entries writes: KV stream, counters, both output files."""
    topt = dict(block_size=2048, filter_policy=1, filter_block_size=2048)
    runs = [r for r in w.random_docdb_runs(12, n_runs=3, n_rows=300) if r]
    ssts = runs_to_ssts(runs, 1024)
    for kw in (dict(retention=False, bottommost=False), dict(retention=False, bottommost=True, last_sequence=o.MAX_SEQ), w.param_grid()[1], w.param_grid()[4]):
        exp = o.compact(ssts, o.CompactionParams(**kw), o.TableOptions(**topt))
        job = pkg.GpuCompactionJob(**topt, **kw)
        for r in runs:
            job.add_input_kv(r)
        job.run()
        st = job.stats()
        assert st.path_flags & pkg.PATH_KV_INPUT and not st.path_flags & (pkg.PATH_FUSED_INGEST | pkg.PATH_GENERAL_DECODE)
        assert job.kv_list() == exp.kv_list()
        assert (st.num_input_records, st.num_output_records) == (exp.stats.num_input_records, exp.stats.num_output_records)
        assert job.digest() == exp.stats.kv_hash
        data, meta = job.fetch_output()
        ref = exp.sst()
        assert (data.tobytes(), meta.tobytes()) == ((ref.data, ref.meta) if ref is not None else (b"", b""))
    # a flush proper: ONE memtable -> one L0 table, plain RocksDB rules (duplicates of a user key: the newest wins)
    mem = w.sort_run([(o.ikey(b"k%05d" % (i // 3), (1 << 50) + i), b"v%d" % i) for i in range(6000)])
    exp = o.compact([o.Sst.build(mem, o.TableOptions(block_size=4096))], o.CompactionParams(retention=False, bottommost=False), o.TableOptions(block_size=4096))
    job = pkg.GpuCompactionJob(block_size=4096, retention=False, bottommost=False)
    job.add_input_kv(mem)
    job.run()
    assert job.kv_list() == exp.kv_list() and len(job.kv_list()) == 2000
    data, meta = job.fetch_output()
    assert data.tobytes() == exp.sst().data and meta.tobytes() == exp.sst().meta
    # empty and single-entry inputs; mixing with table files is refused
    job = pkg.GpuCompactionJob(block_size=4096, retention=False)
    job.add_input_kv([])
    job.add_input_kv(mem[:1])
    job.run()
    assert [v for _, v in job.kv_list()] == [mem[0][1]]
    job = pkg.GpuCompactionJob(block_size=4096)
    job.add_input_kv(mem[:10])
    with pytest.raises(pkg.YbGpuError):
        job.add_input_sst(ssts[0].meta_view(), ssts[0].data_view())
