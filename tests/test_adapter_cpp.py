"""The C++ host adapter (rocksdb::CompactionJob / TableBuilder shaped classes over the C ABI)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "adapter_cpp_test")


def build_bin():
    import __graft_entry__ as g
    g.build()
    lib_dir = os.path.join(ROOT, "yugabyte-db_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", BIN, os.path.join(ROOT, "tests", "adapter_cpp_test.cc"),
                           "-L" + lib_dir, "-lybgpu", "-Wl,-rpath," + lib_dir, "-L/usr/local/cuda/lib64",
                           "-Wl,-rpath,/usr/local/cuda/lib64"])


def test_adapter_compiles_and_fails_loudly_without_gpu():
    build_bin()
    out = subprocess.check_output([BIN, "cpu"], text=True)
    assert out.strip().endswith("OK")


def _adapter_file_filter(frontiers, table_ttl_ns, now, primary=2**64 - 1, cotables=2**64 - 2, mode=0):
    args = [BIN, "filefilter", str(table_ttl_ns), str(primary), str(cotables), str(now), str(mode)]
    for f in frontiers:
        args += ["0", "0", str(2**64 - 2)] if f is None else ["1", str(f[0]), str(f[1])]
    marks, expired = subprocess.check_output(args, text=True).rstrip("\n").split(" ")
    return [c == "D" for c in marks], [c == "E" for c in expired]


def test_adapter_file_filter_reference_known_answers():
    """f4: whole-file TTL expiration as the C++ adapter restates it (DocDBCompactionFileFilter + factory, MarkExpiredFiles)
    against the reference's own known answers (docdb/compaction_file_filter-test.cc:321-484, the table in
    tests/test_oracle_file_filter.py) and, on random frontiers / TTLs / cutoffs / modes, against the oracle."""
    import random
    import oracle_py as o
    import test_oracle_file_filter as K
    build_bin()
    for name, ttl, mode, frontiers, want in K.FILTER_CASES:
        got, _ = _adapter_file_filter(frontiers, ttl, K.NOW, mode=mode)
        assert got == want, name
    rng = random.Random(11)
    specials = [o.NO_EXPIRATION, o.USE_DEFAULT_TTL, o.HT_INVALID]
    for _ in range(150):
        now = K.NOW + rng.randrange(-10**6, 10**6)
        n = rng.randrange(0, 7)
        fr = []
        for _ in range(n):
            if rng.random() < 0.1:
                fr.append(None)
                continue
            created = K.ht_add_seconds(now, rng.choice([-10000, -100, -50, -2, -1, 0, 1, 30, 10000])) + rng.randrange(3)
            if rng.random() < 0.05:
                created = rng.choice([o.HT_MIN, o.HT_MAX, o.HT_INVALID])
            vttl = rng.choice(specials) if rng.random() < 0.5 else K.ht_add_seconds(now, rng.choice([-10000, -100, -10, -1, 0, 1, 10, 10000]))
            fr.append((created, vttl))
        ttl = rng.choice([o.MAX_TTL_NS, 0, 1, 10**9, 20 * 10**9, 1000 * 10**9, 2**62])
        primary = rng.choice([o.HT_MAX, o.HT_INVALID, o.HT_MIN, K.ht_add_seconds(now, rng.choice([-200, -60, -1, 5]))])
        cot = rng.choice([o.HT_INVALID, o.HT_MAX, K.ht_add_seconds(now, rng.choice([-200, -60, 5]))])
        mode = rng.randrange(3)
        got, expired = _adapter_file_filter(fr, ttl, now, primary, cot, mode)
        assert got == o.file_filter(fr, ttl, now, primary, cot, mode), (fr, ttl, now, primary, cot, mode)
        for f, e in zip(fr, expired):
            vt, cr = (o.NO_EXPIRATION, o.HT_MAX) if f is None else (f[1] if f[1] != o.HT_INVALID else o.NO_EXPIRATION, f[0])
            assert e == o.ttl_is_expired(vt, cr, ttl, now, mode)


def test_adapter_check_output_file(tmp_path):
    """CompactionJob::CheckOutputFile (compaction_job.cc:932-971) as the adapter offers it: the table must open; with
    paranoid_file_checks every stored block must match its trailer; the tail-of-zeros check of the data file."""
    import oracle_py as o
    build_bin()
    cfg = o.GenConfig(seed=19, num_rows=2000, cols=2, versions=2, num_files=1, value_len=60)
    kvs = o.Sst.generate(cfg, 0, o.TableOptions(block_size=2048)).read_all()

    def check(meta, data, n=len(kvs), paranoid=1, tail=0):
        b, d = tmp_path / "t.sst", tmp_path / "t.sst.sblock.0"
        b.write_bytes(bytes(meta))
        d.write_bytes(bytes(data))
        return int(subprocess.check_output([BIN, "checkfile", str(b), str(d), str(n), str(paranoid), str(tail)], text=True))

    for comp in (0, 1):
        t = o.Sst.build(kvs, o.TableOptions(block_size=2048, index_block_size=512, min_keys_per_index_block=4, filter_policy=1,
                                            filter_block_size=1024, compression=comp))
        meta, data = bytes(t.meta), bytes(t.data)
        assert check(meta, data) == 0 and check(meta, data, paranoid=0) == 0 and check(meta, data, tail=64) == 0
        bad = bytearray(data)
        bad[len(bad) // 2] ^= 1                                    # a flipped bit in the middle of the data file
        assert check(meta, bad) == 2                               # Corruption
        assert check(meta, bad, paranoid=0) == 0                   # the quick check only opens the table (first block)
        bad = bytearray(data)
        bad[3] ^= 1
        assert check(meta, bad, paranoid=0) == 2
        assert check(meta[:-1], data) == 2 and check(meta[:40], data) == 2          # truncated metadata file
        assert check(meta, data[:len(data) // 2]) == 2                               # truncated data file
        assert check(meta, data + bytes(100), tail=64) == 2 and check(meta, data + bytes(100), tail=0) == 0
        assert check(b"", b"", n=0) == 0                            # nothing survived: no file, nothing to check (:950-952)


@pytest.mark.gpu
def test_adapter_runs_compaction(tmp_path):
    import oracle_py as o
    build_bin()
    cfg = o.GenConfig(seed=8, num_rows=3000, cols=2, versions=3, num_files=3, value_len=80)
    ssts = o.Sst.generate_all(cfg)
    args = [BIN, "gpu"]
    for i, s in enumerate(ssts):
        b, d = tmp_path / ("%d.sst" % i), tmp_path / ("%d.sst.sblock.0" % i)
        b.write_bytes(s.meta)
        d.write_bytes(s.data)
        args += [str(b), str(d)]
    out = subprocess.check_output(args, text=True)
    assert "OK in=18000" in out
    exp = o.compact(ssts, o.CompactionParams())
    assert (tmp_path / "0.sst.out.data").read_bytes() == exp.sst().data
    assert (tmp_path / "0.sst.out.base").read_bytes() == exp.sst().meta


@pytest.mark.gpu
def test_adapter_snappy_in_snappy_out(tmp_path):
    """The production configuration end to end through the C++ adapter: Snappy-compressed input tables, Snappy-compressed
    output (Options::compression = kSnappyCompression, docdb_rocksdb_util.cc:184); files equal the oracle's."""
    import oracle_py as o
    import test_gpu_parity as T
    build_bin()
    runs = T._phrase_runs(77, 3, 800, random_every=150)
    ssts = [o.Sst.build(r, o.TableOptions(block_size=4096, compression=1)) for r in runs]
    args = [BIN, "gpusnappy"]
    for i, s in enumerate(ssts):
        b, d = tmp_path / ("%d.sst" % i), tmp_path / ("%d.sst.sblock.0" % i)
        b.write_bytes(bytes(s.meta))
        d.write_bytes(bytes(s.data))
        args += [str(b), str(d)]
    out = subprocess.check_output(args, text=True)
    assert "OK in=%d" % sum(len(r) for r in runs) in out
    exp = o.compact(ssts, o.CompactionParams(), o.TableOptions(compression=1))
    plain = o.compact(ssts, o.CompactionParams(), o.TableOptions())
    assert (tmp_path / "0.sst.out.data").read_bytes() == bytes(exp.sst().data)
    assert (tmp_path / "0.sst.out.base").read_bytes() == bytes(exp.sst().meta)
    assert len(exp.sst().data) < 0.9 * len(plain.sst().data)


@pytest.mark.gpu
def test_adapter_runs_subcompactions(tmp_path):
    """GpuCompactionJob with max_subcompactions = 4: one output file per key range, in range order;
    together they hold exactly the single-job KV stream."""
    import oracle_py as o
    build_bin()
    cfg = o.GenConfig(seed=9, num_rows=6000, cols=2, versions=3, num_files=4, value_len=80, tombstone_per_1024=40)
    ssts = o.Sst.generate_all(cfg, o.TableOptions(block_size=4096))
    args = [BIN, "gpusub"]
    for i, s in enumerate(ssts):
        b, d = tmp_path / ("%d.sst" % i), tmp_path / ("%d.sst.sblock.0" % i)
        b.write_bytes(s.meta)
        d.write_bytes(s.data)
        args += [str(b), str(d)]
    out = subprocess.check_output(args, text=True)
    assert "OK in=36000" in out
    n_files = int(out.strip().split("files=")[1])
    assert 2 <= n_files <= 4
    exp = o.compact(ssts, o.CompactionParams())
    got = []
    for i in range(n_files):
        part = o.Sst.from_bytes((tmp_path / ("0.sst.sub%d.base" % i)).read_bytes(), (tmp_path / ("0.sst.sub%d.data" % i)).read_bytes())
        got += part.read_all()
    assert got == exp.kv_list()
    # ... and as ONE table (ConcatenatedOutput): the same key/value stream behind one index
    one = o.Sst.from_bytes((tmp_path / "0.sst.one.base").read_bytes(), (tmp_path / "0.sst.one.data").read_bytes())
    assert one.read_all() == exp.kv_list()


@pytest.mark.gpu
def test_adapter_run_into_feed(tmp_path):
    """a8: CompactionFeed interface. RunIntoFeed pushes the surviving stream through a host feed into the host
    TableBuilder: the files equal the oracle's (= the GPU-built ones); a failing Feed aborts with its own status."""
    import oracle_py as o
    build_bin()
    cfg = o.GenConfig(seed=12, num_rows=2500, cols=2, versions=3, num_files=3, value_len=70, tombstone_per_1024=50)
    ssts = o.Sst.generate_all(cfg)
    args = [BIN, "gpufeed"]
    for i, s in enumerate(ssts):
        b, d = tmp_path / ("%d.sst" % i), tmp_path / ("%d.sst.sblock.0" % i)
        b.write_bytes(s.meta)
        d.write_bytes(s.data)
        args += [str(b), str(d)]
    out = subprocess.check_output(args, text=True)
    assert "OK in=15000" in out, out
    exp = o.compact(ssts, o.CompactionParams())
    assert (tmp_path / "0.sst.feed.data").read_bytes() == exp.sst().data
    assert (tmp_path / "0.sst.feed.base").read_bytes() == exp.sst().meta
