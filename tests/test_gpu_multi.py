"""Multi-GPU paths on real GPUs (skipped on a 1-GPU box): key-range sharded compaction of one tablet
with one NCCL all_to_all (BASELINE config 5, scaled), and the key-range filter on a single GPU."""
import importlib
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _ranges_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import numpy as np
    import torch
    import torch.distributed as dist
    import oracle_py as o
    pkg = importlib.import_module("yugabyte-db_b200")
    rs = importlib.import_module("yugabyte-db_b200.range_sharded")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    cfg = o.GenConfig(seed=31, num_rows=60000, cols=2, versions=3, num_files=8, value_len=120, tombstone_per_1024=40)
    ssts = o.Sst.generate_all(cfg, o.TableOptions(block_size=8192))
    cutoff = o.ht_from_micros(cfg.base_micros + 1500)
    all_last = max(s.read_all()[-1][0][:-8] for s in ssts)
    mine = [(s.meta_view().copy(), s.data_view().copy()) for f, s in enumerate(ssts) if f % world == rank]
    job, my_range, info = rs.compact(mine, rank, world, rank, all_last, cutoff_ht=cutoff, block_size=8192)
    piece = job.kv_list() if job is not None else []
    pieces = [None] * world
    dist.all_gather_object(pieces, piece)
    stats = [None] * world
    dist.all_gather_object(stats, (job.stats().num_input_records if job else 0, info["sent_bytes"], info["recv_bytes"]))
    if rank == 0:
        exp = o.compact(ssts, o.CompactionParams(cutoff_ht=cutoff), o.TableOptions(block_size=8192))
        ok = [kv for p in pieces for kv in p] == exp.kv_list()
        ok = ok and sum(s[0] for s in stats) == exp.stats.num_input_records
        ok = ok and all(len(p) > 0 for p in pieces)
        q.put((ok, [len(p) for p in pieces], stats))
    dist.barrier()
    dist.destroy_process_group()


def test_key_range_sharded_compaction_two_gpus():
    pkg = importlib.import_module("yugabyte-db_b200")
    if pkg.device_count() < 2:                     # checked without importing torch (cold import is slow)
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    mp.spawn(_ranges_worker, args=(2, port, q), nprocs=2, join=True)
    ok, lens, stats = q.get(timeout=10)
    assert ok, (lens, stats)


def test_key_range_filter_single_gpu():
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import oracle_py as o
    pkg = importlib.import_module("yugabyte-db_b200")
    sh = importlib.import_module("yugabyte-db_b200.sharding")
    cfg = o.GenConfig(seed=13, num_rows=20000, cols=2, versions=3, num_files=4, value_len=60, tombstone_per_1024=30)
    ssts = o.Sst.generate_all(cfg, o.TableOptions(block_size=4096))
    cutoff = o.ht_from_micros(cfg.base_micros + 1500)
    exp = o.compact(ssts, o.CompactionParams(cutoff_ht=cutoff), o.TableOptions(block_size=4096))
    keys = sorted(k[:-8] for s in ssts for k, _ in s.read_all())
    splitters = [keys[len(keys) // 3][:32], keys[2 * len(keys) // 3][:32]]       # DocKey (32 B) aligned
    out, n_in = [], 0
    for r in range(3):
        lo, hi = sh.range_of_rank(splitters, r)
        job = pkg.GpuCompactionJob(cutoff_ht=cutoff, block_size=4096, largest_user_key=keys[-1], range_lower=lo, range_upper=hi)
        for s in ssts:
            job.add_input_sst(s.meta_view(), s.data_view())
        st = job.run()
        n_in += st.num_input_records
        out += job.kv_list()
    assert out == exp.kv_list()
    assert n_in == exp.stats.num_input_records
