"""Multi-process (gloo, world_size 2, CPU) tests of the host-side multi-GPU logic: tablet placement,
result aggregation, and the key-range plan + all_to_all exchange used for one oversized tablet.
The per-rank compaction itself is played by the oracle here (no GPU in this container)."""
import importlib
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def sharding():
    spec = importlib.util.spec_from_file_location("ybgpu_sharding", os.path.join(ROOT, "yugabyte-db_b200", "sharding.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_assign_tablets_balanced():
    sh = sharding()
    sizes = [31, 7, 7, 7, 20, 1, 1, 12, 5, 9]
    for n in (1, 2, 3, 8):
        parts = sh.assign_tablets(sizes, n)
        assert sorted(i for p in parts for i in p) == list(range(len(sizes)))
        loads = [sum(sizes[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(sizes)
    assert sh.assign_tablets([5] * 64, 8) == [list(range(g, 64, 8)) for g in range(8)]


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_py as o
    sh = sharding()
    # ---- tablet sharding: 6 tablets x 3-way compaction, placed on 2 ranks
    n_tablets = 6
    cfgs = [o.GenConfig(seed=40 + t, num_rows=300 + 100 * t, cols=2, versions=3, num_files=3, value_len=40,
                        row_offset=t * 10000, hash_rows_total=10**6) for t in range(n_tablets)]
    sizes = [c.num_rows for c in cfgs]
    cutoff = o.ht_from_micros(cfgs[0].base_micros + 1500)

    def run(t):
        ssts = o.Sst.generate_all(cfgs[t], o.TableOptions(block_size=2048), max_threads=1)
        r = o.compact(ssts, o.CompactionParams(cutoff_ht=cutoff), o.TableOptions(block_size=2048))
        return (r.stats.num_input_records, r.stats.num_output_records, r.stats.kv_hash)
    mine = sh.run_sharded(cfgs, rank, world, run, sizes)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    merged = {}
    for g in gathered:
        assert not (set(g) & set(merged))
        merged.update(g)
    assert sorted(merged) == list(range(n_tablets))
    tot = torch.tensor([sum(v[0] for v in mine.values()), sum(v[1] for v in mine.values())], dtype=torch.int64)
    dist.all_reduce(tot)
    if rank == 0:
        single = {t: run(t) for t in range(n_tablets)}
        assert single == merged
        assert tot.tolist() == [sum(v[0] for v in single.values()), sum(v[1] for v in single.values())]

    # ---- one oversized tablet: key-range plan + all_to_all of input slices
    big = o.GenConfig(seed=77, num_rows=3000, cols=2, versions=4, num_files=4, value_len=60)
    ssts = o.Sst.generate_all(big, o.TableOptions(block_size=1024), max_threads=1)
    my_files = [f for f in range(big.num_files) if f % world == rank]     # files staged round-robin
    kvs = {f: ssts[f].read_all() for f in my_files}
    # samples: every 16th user key of my files, weight = bytes since the previous sample
    samples = []
    for f in my_files:
        acc = 0
        for i, (k, v) in enumerate(kvs[f]):
            acc += len(k) + len(v)
            if i % 16 == 15:
                samples.append((k[:-8], acc))
                acc = 0
    all_samples = [None] * world
    dist.all_gather_object(all_samples, samples)
    flat = sorted(s for part in all_samples for s in part)
    splitters = sh.plan_key_ranges([k for k, _ in flat], [w for _, w in flat], world)
    assert len(splitters) == world - 1
    # slice my files by owner range and exchange (variable sizes -> all_to_all of byte tensors)
    send = []
    for dst in range(world):
        lo, hi = sh.range_of_rank(splitters, dst)
        part = [(k, v) for f in my_files for (k, v) in kvs[f] if (not lo or k[:-8] >= lo) and (not hi or k[:-8] < hi)]
        send.append(part)
    recv = [None] * world
    dist.all_to_all_object = None
    out_lists = [None] * world
    for dst in range(world):                       # gloo has no all_to_all for objects: gather per destination
        got = [None] * world
        dist.all_gather_object(got, send[dst])
        if dst == rank:
            out_lists = got
    # local compaction of my range
    runs = [sorted(x, key=lambda kv: (kv[0][:-8], -int.from_bytes(kv[0][-8:], "little"))) for x in out_lists if x]
    all_keys = sorted(k[:-8] for f in range(big.num_files) for k, _ in ssts[f].read_all())
    params = o.CompactionParams(cutoff_ht=o.ht_from_micros(big.base_micros + 2500), largest_user_key=all_keys[-1])
    local = o.compact_runs(runs, params).kv_list()
    pieces = [None] * world
    dist.all_gather_object(pieces, local)
    if rank == 0:
        whole = o.compact(ssts, o.CompactionParams(cutoff_ht=o.ht_from_micros(big.base_micros + 2500))).kv_list()
        assert [kv for p in pieces for kv in p] == whole, "range-sharded output must concatenate to the single-job output"
        # ranges are row-group aligned: no DocKey appears in two pieces
        last_rows = [p[-1][0][:32] for p in pieces if p]
        first_rows = [p[0][0][:32] for p in pieces if p]
        for a, b in zip(last_rows, first_rows[1:]):
            assert a < b
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo(tmp_path):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
