#!/usr/bin/env python
"""profiles/e2e_sweep.py — explores the end-to-end (host files in, host files out) arm of bench.py on a GPU
box: one set of config-2 inputs, then ybgpu_compact_files under several settings (bulk-copy chunk size,
zero-copy status reads, number of key ranges, ranges in flight). Prints one JSON line per setting;
`--trace` adds the per-range timeline of one step (YBGPU_SUB_TRACE) on stderr. Not a bench value source:
bench.py measures the committed defaults.

    python profiles/e2e_sweep.py --rows 100000000 > gpurun_out/e2e_sweep.jsonl 2> gpurun_out/e2e_sweep.err
"""
import argparse
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100_000_000)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--trace", action="store_true")
    args = ap.parse_args()
    import torch
    pkg = importlib.import_module("yugabyte-db_b200")
    torch.cuda.set_device(0)
    cfg = pkg.GenConfig(seed=2, num_rows=args.rows, cols=1, versions=1, num_files=8, value_len=256, hash_rows_total=args.rows)
    ssts = pkg.generate_ssts(cfg, max_threads=8)
    in_bytes = sum(s.raw_bytes for s in ssts)
    file_bytes = sum(s.data_view().size for s in ssts)
    cudart = torch.cuda.cudart()
    for s in ssts:
        v = s.data_view()
        assert int(cudart.cudaHostRegister(v.ctypes.data, v.size, 0)) == 0
    out_data = torch.empty(file_bytes + (64 << 20), dtype=torch.uint8, pin_memory=True).numpy()
    out_meta = torch.empty(max(64 << 20, file_bytes // 100), dtype=torch.uint8, pin_memory=True).numpy()
    files = [(s.meta_view(), s.data_view()) for s in ssts]

    def run(chunk_mb, zc, subs, inflight, steps):
        os.environ["YBGPU_COPY_CHUNK_MB"] = str(chunk_mb)
        os.environ["YBGPU_ZC_STATUS"] = str(zc)
        best, tot = None, 0.0
        for i in range(steps + 1):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = pkg.compact_files(files, max_subcompactions=subs, max_in_flight=inflight, data_arena=out_data,
                                  meta_arena=out_meta, filter_policy=1, verify_checksums=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            assert r.total.num_input_records == args.rows
            if i:
                tot += dt
                best = dt if best is None else min(best, dt)
        print(json.dumps({"chunk_mb": chunk_mb, "zero_copy_status": zc, "subcompactions": subs, "in_flight": inflight,
                          "ranges": len(r.outputs), "ms_per_step": round(tot / steps * 1e3, 1), "best_ms": round(best * 1e3, 1),
                          "gb_per_s": round(in_bytes * steps / tot / 1e9, 2), "gpu_ms_sum": round(r.total.gpu_seconds * 1e3, 1)}), flush=True)

    grid = [(0, 0, 16, 4), (32, 0, 16, 4), (0, 1, 16, 4), (32, 1, 16, 4), (8, 1, 16, 4), (32, 1, 16, 2), (32, 1, 16, 3),
            (32, 1, 16, 6), (32, 1, 8, 3), (32, 1, 32, 4), (32, 1, 32, 6), (32, 1, 1, 1)]
    for g in grid:
        run(*g, steps=args.steps)
    if args.trace:
        os.environ["YBGPU_SUB_TRACE"] = "1"
        run(32, 1, 16, 4, steps=1)


if __name__ == "__main__":
    main()
