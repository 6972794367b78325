#!/usr/bin/env python
"""TEST / BENCH INFRASTRUCTURE (oracle side): one CPU compaction worker process for bench.py's all-cores figure.

Generates its own sample of the config-2 shape, waits until --start-at (epoch seconds, so that all workers begin
together), then runs one-thread compactions with the oracle until --seconds have passed. Prints one JSON line:
{"done": n, "elapsed": s, "bytes_each": b}. Separate PROCESSES, not threads: concurrent compactions of different
tablets share nothing in the reference either (one RocksDB instance per tablet), and threads of one process would
contend on the allocator / address-space lock, which understates the host."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--seconds", type=float, default=4.0)
    ap.add_argument("--start-at", type=float, default=0.0)
    ap.add_argument("--verify", type=int, default=1)
    ap.add_argument("--seed", type=int, default=2)
    args = ap.parse_args()
    import oracle_py as o
    cfg = o.GenConfig(seed=args.seed, num_rows=args.rows, cols=1, versions=1, num_files=8, value_len=256)
    ssts = o.Sst.generate_all(cfg, o.TableOptions())
    b = sum(s.raw_bytes for s in ssts)
    while time.time() < args.start_at:
        time.sleep(0.005)
    t0 = time.perf_counter()
    done = 0
    while time.perf_counter() - t0 < args.seconds:
        r = o.compact(ssts, o.CompactionParams(), o.TableOptions(filter_policy=1), mode=o.BUILD_SST | o.NO_HASH, verify=bool(args.verify))
        del r
        done += 1
    print(json.dumps({"done": done, "elapsed": time.perf_counter() - t0, "bytes_each": b}))


if __name__ == "__main__":
    main()
