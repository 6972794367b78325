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
