#!/usr/bin/env python
"""bench.py — compaction throughput of the B200 engine on BASELINE.json's metric.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one whole compaction job over the workload (SURVEY.md 8d / BASELINE.md §3 config 2:
8-way major compaction, 100 M entries, 32-B DocKey + 256-B value, kNoCompression SSTs). With N>1
every rank compacts its own tablet of that shape (tablets are independent: no data-path
collective, weak scaling).

JSON line (one, rank 0): metric = GB/s of input bytes merged (raw key+value bytes of the input
entries, as rocksdb.raw.key.size + rocksdb.raw.value.size count them).
  value     inputs already resident in HBM; whole job device pipeline, wall clock between syncs.
  e2e       same job through the C ABI with HOST (pinned) input files and HOST output files:
            H2D of every input file and D2H of the result inside the timed region. Headline mode:
            ybgpu_compact_files with --subcompactions key ranges (CompactionJob's subcompaction
            mechanism, compaction_job.cc:409-552), pipelined so that H2D / kernels / D2H of different
            ranges overlap; one output SST per range. e2e.one_table adds ybgpu_sst_concat_meta: the range
            outputs assembled into ONE table (data pieces appended in range order, one rebased index /
            filter index) inside the timed region. e2e.single_job is the same measurement with one
            job and one output file (H2D, run and D2H back to back) — the shape DocDB's single-level
            universal layout produces today (db/compaction.cc:593-604 never forms subcompactions there).
  roofline  dominant kernel, algorithmic bytes / its CUDA-event time.
  cpu_baseline  the oracle (CPU restatement of the reference loop) on a bounded sample, 1 thread
            like the reference (max_subcompactions = 1).
"""
import argparse
import ctypes
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DEFAULT_ROWS = 100_000_000       # config 2: 100 M entries
VALUE_LEN = 256
NUM_FILES = 8
WORKLOAD = "8-way major compaction, 100M keys, 32-B DocKey / 256-B value, 1 GPU"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=DEFAULT_ROWS, help="entries per tablet (debug: smaller)")
    ap.add_argument("--sample-rows", type=int, default=6_000_000, help="entries in the CPU-baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--verify", type=int, default=1, help="verify input block checksums (reference default: on)")
    ap.add_argument("--subcompactions", type=int, default=32,
                    help="e2e arm: key-range subcompactions per job (DBOptions::max_subcompactions; 1 = one job, one output file)")
    ap.add_argument("--in-flight", type=int, default=6, help="e2e arm: subcompactions in flight (host threads / private streams)")
    ap.add_argument("--workload", default="config2", choices=["config2", "mvcc"],
                    help="config2 = BASELINE configs[1] (the bench line); mvcc = configs[3] shape (20 versions/key, "
                         "history cutoff drops 90 %), scaled to --rows entries, for profiles/ only")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 7:
                self.samples.append(parts)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(float(s[0])) for s in self.samples if s[0].replace(".", "").isdigit())
        mx = [int(float(s[1])) for s in self.samples if s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """--impl reference: the reference's CPU loop (oracle port; the reference tree itself cannot be
    compiled in this image) on the host cores. One compaction = one thread, as in the reference
    (rocksdb/util/options.cc:258, db/compaction.cc:593-604), on a bounded sample of the workload.
    Two informational figures ride along: many independent tablets on all cores, and the same
    compaction cut into key ranges with one thread per range (the CPU counterpart of the GPU arm's
    pipelined subcompactions)."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as o
    rows = min(args.rows, args.sample_rows)
    cfg = o.GenConfig(seed=2, num_rows=rows, cols=1, versions=1, num_files=NUM_FILES, value_len=VALUE_LEN)
    ssts = o.Sst.generate_all(cfg, o.TableOptions())
    in_bytes = sum(s.raw_bytes for s in ssts)
    params = o.CompactionParams()
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        r = o.compact(ssts, params, o.TableOptions(filter_policy=1), mode=o.BUILD_SST | o.NO_HASH, verify=bool(args.verify))
        t1 = time.perf_counter()
        n_out = r.stats.num_output_records
        del r
        if i >= args.warmup:
            times.append(t1 - t0)
    total = sum(times)
    gbs = in_bytes * args.steps / total / 1e9
    sample = "%d entries (%0.2f GB raw) of the same 8-way shape, %d output entries" % (rows, in_bytes / 1e9, n_out)
    # Informational: what the host delivers across MANY tablets — T independent compactions of the
    # same sample, one thread each, run concurrently. One job cannot use more than one thread in the reference
    # (max_subcompactions = 1, rocksdb/util/options.cc:258; universal compaction with one level never forms
    # subcompactions, db/compaction.cc:593-604), so the headline value above stays the one-thread figure.
    many = None
    try:
        from concurrent.futures import ThreadPoolExecutor
        T = max(1, min(os.cpu_count() or 1, 32))
        small_rows = min(rows, 1_000_000)
        if small_rows != rows:
            cfg2 = o.GenConfig(seed=2, num_rows=small_rows, cols=1, versions=1, num_files=NUM_FILES, value_len=VALUE_LEN)
            ssts2 = o.Sst.generate_all(cfg2, o.TableOptions())
        else:
            ssts2 = ssts
        b2 = sum(s.raw_bytes for s in ssts2)

        def one(_):
            r = o.compact(ssts2, params, o.TableOptions(filter_policy=1), mode=o.BUILD_SST | o.NO_HASH, verify=bool(args.verify))
            del r
        t0 = time.perf_counter()
        with ThreadPoolExecutor(T) as ex:
            list(ex.map(one, range(T)))
        dt = time.perf_counter() - t0
        many = {"value": round(T * b2 / dt / 1e9, 3), "unit": "GB/s", "cores": T,
                "sample": "%d concurrent one-thread compactions of %d entries each (independent tablets), %.1f s" % (T, small_rows, dt)}
    except Exception as e:   # never fail the arm because of the informational figure
        many = {"error": str(e)}
    # Informational: the SAME compaction cut into key ranges, one CPU-port thread per range — what
    # CompactionJob's subcompactions (compaction_job.cc:409-552) would give the host on a layout that forms them
    # (DocDB's single-level universal layout never does, db/compaction.cc:593-604). This is the CPU counterpart of
    # the GPU arm's pipelined e2e mode; the per-range input slices are cut outside the timed region.
    subs = None
    try:
        import bisect
        from concurrent.futures import ThreadPoolExecutor
        T = max(1, min(os.cpu_count() or 1, 32))
        rows_s = min(rows, 2_000_000)
        cfg3 = o.GenConfig(seed=2, num_rows=rows_s, cols=1, versions=1, num_files=NUM_FILES, value_len=VALUE_LEN)
        ssts3 = o.Sst.generate_all(cfg3, o.TableOptions())
        b3 = sum(s.raw_bytes for s in ssts3)
        kvs = [s.read_all() for s in ssts3]
        uks = [[k[:-8] for k, _ in f] for f in kvs]
        dockeys = sorted(u[:32] for u in uks[0])                      # rows of this workload are 32-byte DocKeys
        splitters = [dockeys[len(dockeys) * i // T] for i in range(1, T)]
        largest = max(u[-1] for u in uks if u)
        range_ssts = []
        for r in range(T):
            lo = splitters[r - 1] if r > 0 else None
            hi = splitters[r] if r < T - 1 else None
            parts = []
            for f, u in zip(kvs, uks):
                a = bisect.bisect_left(u, lo) if lo is not None else 0
                b = bisect.bisect_left(u, hi) if hi is not None else len(u)
                if b > a:
                    parts.append(o.Sst.build(f[a:b], o.TableOptions()))
            range_ssts.append(parts)
        del kvs, uks
        p3 = o.CompactionParams(largest_user_key=largest)

        def one_range(r):
            if range_ssts[r]:
                res = o.compact(range_ssts[r], p3, o.TableOptions(filter_policy=1), mode=o.BUILD_SST | o.NO_HASH, verify=bool(args.verify))
                del res
        t0 = time.perf_counter()
        with ThreadPoolExecutor(T) as ex:
            list(ex.map(one_range, range(T)))
        dt = time.perf_counter() - t0
        subs = {"value": round(b3 / dt / 1e9, 3), "unit": "GB/s", "cores": T,
                "sample": "one compaction of %d entries cut into %d key ranges, one thread per range, %.2f s" % (rows_s, T, dt)}
    except Exception as e:
        subs = {"error": str(e)}
    line = {
        "impl": "reference", "metric": "compaction GB/s (input bytes merged)", "value": round(gbs, 4), "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(total / args.steps * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample, "threads_per_compaction": 1},
        "mkeys_per_s": round(rows * args.steps / total / 1e6, 3),
        "cpu_baseline": {"value": round(gbs, 4), "unit": "GB/s", "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": round(gbs, 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "many_tablets_all_cores": many,
        "subcompactions_all_cores": subs,
    }
    emit_json_line(line)


# ------------------------------------------------------------------------------------------------
def cpu_baseline(args):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as o
    rows = min(args.rows, args.sample_rows)
    cfg = o.GenConfig(seed=2, num_rows=rows, cols=1, versions=1, num_files=NUM_FILES, value_len=VALUE_LEN)
    ssts = o.Sst.generate_all(cfg, o.TableOptions())
    in_bytes = sum(s.raw_bytes for s in ssts)
    t0 = time.perf_counter()
    r = o.compact(ssts, o.CompactionParams(), o.TableOptions(filter_policy=1), mode=o.BUILD_SST | o.NO_HASH, verify=bool(args.verify))
    dt = time.perf_counter() - t0
    del r
    return {"value": round(in_bytes / dt / 1e9, 4), "unit": "GB/s", "cores": 1, "kind": "port",
            "sample": "%d entries (%0.2f GB raw) of the same 8-way shape, %.1f s on one host thread (the reference "
                      "runs one thread per compaction)" % (rows, in_bytes / 1e9, dt),
            "mkeys_per_s": round(rows / dt / 1e6, 3)}


_REAL_STDOUT = None


def quiet_stdout():
    """stdout carries exactly one JSON line: anything libraries print there (NCCL's version banner, torchrun
    notices) is rerouted to stderr; the JSON goes to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit_json_line(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    quiet_stdout()
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("yugabyte-db_b200")
    if not torch.cuda.is_available() or pkg.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the compaction engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner to stdout at NCCL_DEBUG=VERSION
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- inputs: this rank's tablet (distinct key range per rank) ----
    versions = 20 if args.workload == "mvcc" else 1
    nrows = args.rows // versions
    cfg = pkg.GenConfig(seed=2 + rank, num_rows=nrows, cols=1, versions=versions, num_files=NUM_FILES, value_len=VALUE_LEN,
                        row_offset=rank * nrows, hash_rows_total=nrows * world)
    # DocDB tables carry the DocKeyV3 fixed-size bloom filter (docdb_rocksdb_util.cc:761-763): both arms build it
    job_kw = {"filter_policy": 1}
    if args.workload == "mvcc":
        # versions 0..18 are at or below the cutoff (only the newest of them survives), version 19 is above
        job_kw["cutoff_ht"] = ((cfg.base_micros + 18 * 1000 + 500) << 12)
    t0 = time.perf_counter()
    ssts = pkg.generate_ssts(cfg, max_threads=NUM_FILES)
    gen_s = time.perf_counter() - t0
    in_bytes = sum(s.raw_bytes for s in ssts)
    n_entries = sum(s.num_entries for s in ssts)
    file_bytes = sum(s.data_view().size for s in ssts)

    # block handles from each file's own index (host, once; not part of the hot path)
    handles = [read_handles(pkg, s) for s in ssts]

    # ---- HBM-resident arm ----
    dev_files = []
    for s in ssts:
        v = s.data_view()
        t = torch.empty(v.size + 64, dtype=torch.uint8, device="cuda")
        t[16:16 + v.size].copy_(torch.from_numpy(v))
        dev_files.append(t)
    stream_ptr = torch.cuda.current_stream().cuda_stream

    host_ms = {"create": 0.0, "add_inputs": 0.0, "run": 0.0, "close": 0.0}

    def step_resident():
        t0 = time.perf_counter()
        job = pkg.GpuCompactionJob(device=local_rank, verify_checksums=False, cuda_stream=stream_ptr, **job_kw)
        t1 = time.perf_counter()
        for t, s, (off, sz) in zip(dev_files, ssts, handles):
            job.add_input_device(t.data_ptr() + 16, s.data_view().size, off, sz)
        t2 = time.perf_counter()
        st = job.run()
        t3 = time.perf_counter()
        d = st.as_dict()
        job.close()
        t4 = time.perf_counter()
        for k, v in zip(("create", "add_inputs", "run", "close"), (t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
            host_ms[k] += v * 1e3
        return d

    for _ in range(args.warmup):
        step_resident()
    for k in host_ms:
        host_ms[k] = 0.0
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    stats = [step_resident() for _ in range(args.steps)]
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    clock_info = clocks.stop()
    ev_s = e0.elapsed_time(e1) / 1e3
    step_s = max(wall, ev_s)
    tt = torch.tensor([step_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_s = float(tt.item())
    launches = sum(s["gpu_kernel_launches"] for s in stats)
    out_bytes = stats[-1]["total_output_raw_key_bytes"] + stats[-1]["total_output_raw_value_bytes"]
    phases = [sum(s["phase_seconds"][i] for s in stats) / args.steps for i in range(5)]
    enc_kernel_s = sum(s["phase_seconds"][5] for s in stats) / args.steps      # k_encode_smem alone (CUDA events around its launch)
    gpu_s = sum(s["gpu_seconds"] for s in stats) / args.steps

    # ---- e2e arm: host (pinned) files in, host files out ----
    e2e = None
    if not args.no_e2e:
        cudart = torch.cuda.cudart()
        pinned = []
        for s in ssts:
            v = s.data_view()
            rc = cudart.cudaHostRegister(v.ctypes.data, v.size, 0)
            pinned.append((v, int(rc) == 0))

        # pinned host buffers for the output files, reused by every step
        out_data = torch.empty(file_bytes + (64 << 20), dtype=torch.uint8, pin_memory=True).numpy()
        out_meta = torch.empty(max(64 << 20, file_bytes // 100), dtype=torch.uint8, pin_memory=True).numpy()

        e2e_ms = {"add_inputs_h2d": 0.0, "run": 0.0, "fetch_output_d2h": 0.0, "close": 0.0}

        def step_e2e():
            t0 = time.perf_counter()
            job = pkg.GpuCompactionJob(device=local_rank, verify_checksums=bool(args.verify), cuda_stream=stream_ptr, **job_kw)
            for s, (off, sz) in zip(ssts, handles):
                job.add_input(s.data_view(), off, sz)
            t1 = time.perf_counter()
            job.run()
            t2 = time.perf_counter()
            data, meta = job.fetch_output(out_data, out_meta)
            t3 = time.perf_counter()
            st = job.stats().as_dict()
            job.close()
            t4 = time.perf_counter()
            for k, v in zip(e2e_ms, (t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
                e2e_ms[k] += v * 1e3
            return st, data.size + meta.size

        def timed(step_fn):
            for _ in range(min(args.warmup, 2) if args.rows >= 50_000_000 else args.warmup):
                step_fn()
            for k in e2e_ms:
                e2e_ms[k] = 0.0
            barrier()
            t0 = time.perf_counter()
            res = [step_fn() for _ in range(args.steps)]
            barrier()
            dt = time.perf_counter() - t0
            te = torch.tensor([dt], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            return float(te.item()), res

        e2e_s, res = timed(step_e2e)
        single = {"value": round(in_bytes * world * args.steps / e2e_s / 1e9, 4), "unit": "GB/s",
                  "h2d_bytes_per_step": int(res[-1][0]["h2d_bytes"]), "d2h_bytes_per_step": int(res[-1][0]["d2h_bytes"]),
                  "ms_per_step": round(e2e_s / args.steps * 1e3, 2), "output_file_bytes": int(res[-1][1]),
                  "host_ms_per_step": {k: round(v / args.steps, 2) for k, v in e2e_ms.items()}}
        def verify_outputs(file_list, stride=257):
            # host-side CRC32C check of every stride-th output data block (outside the timed regions): the full-size
            # outputs cannot be compared with the oracle, but a wrong or torn device->host copy cannot pass this
            try:
                checked = bad = 0
                for meta_v, data_v in file_list:
                    c, b_ = pkg.sst_verify_blocks(meta_v, data_v, stride)
                    checked += c
                    bad += b_
                return {"blocks_checked": int(checked), "bad_blocks": int(bad), "stride": stride}
            except Exception as ex:
                return {"error": "%s: %s" % (type(ex).__name__, ex)}

        try:
            dlen = int(res[-1][0]["output_data_file_size"])
            mlen = int(res[-1][0]["output_meta_file_size"])
            single["output_check"] = verify_outputs([(out_meta[:mlen], out_data[:dlen])])
        except Exception as ex:
            single["output_check"] = {"error": "%s: %s" % (type(ex).__name__, ex)}
        e2e = dict(single, pinned_inputs=all(ok for _, ok in pinned), verify_checksums=bool(args.verify),
                   mode="one job, one output file")
        if args.subcompactions > 1:
            files = [(s.meta_view(), s.data_view()) for s in ssts]

            def step_sub():
                r = pkg.compact_files(files, max_subcompactions=args.subcompactions, max_in_flight=args.in_flight,
                                      data_arena=out_data, meta_arena=out_meta, device=local_rank,
                                      verify_checksums=bool(args.verify), **job_kw)
                last_sub[0] = r
                return r.total.as_dict(), sum(o_.data_len + o_.meta_len for o_ in r.outputs), len(r.outputs)

            last_sub = [None]
            sub_s, sres = timed(step_sub)
            try:
                sub_check = verify_outputs([(out_meta[o_.meta_offset:o_.meta_offset + o_.meta_len], out_data[o_.data_offset:o_.data_offset + o_.data_len])
                                            for o_ in last_sub[0].outputs if o_.data_len])
            except Exception as ex:
                sub_check = {"error": "%s: %s" % (type(ex).__name__, ex)}
            assert sres[-1][0]["num_input_records"] == n_entries, "subcompactions must see every input entry once"

            # One table out of the range outputs (what a single-level universal layout such as DocDB's needs): the
            # range data files are appended in range order as they are, ybgpu_sst_concat_meta writes the one metadata
            # file (rebased multi-level index, all filter blocks + one filter index, summed properties).
            one_table = None
            try:
                if world > 1:
                    raise RuntimeError("measured at N=1 only (an exception on one rank must not strand the others at a barrier)")
                concat_buf = np.empty(2 * out_meta.size + (1 << 20), np.uint8)
                concat_buf[::4096] = 0                      # touch the pages once, outside the timed region

                def step_one_table():
                    r = pkg.compact_files(files, max_subcompactions=args.subcompactions, max_in_flight=args.in_flight,
                                          data_arena=out_data, meta_arena=out_meta, device=local_rank,
                                          verify_checksums=bool(args.verify), **job_kw)
                    outs = [o_ for o_ in r.outputs if o_.data_len]
                    pieces = [(out_meta[o_.meta_offset:o_.meta_offset + o_.meta_len], o_.data_len, o_.smallest, o_.largest) for o_ in outs]
                    meta = pkg.sst_concat_meta(pieces, out=concat_buf, filter_policy=job_kw.get("filter_policy", 0))
                    return r.total.as_dict(), sum(o_.data_len for o_ in outs), int(meta.size), len(outs), meta
                ot_s, ores = timed(step_one_table)
                st_d, data_bytes, meta_bytes, n_pieces, meta = ores[-1]
                off, sz, _ = pkg.sst_block_handles(meta)     # the product's own reader walks the merged index
                assert len(off) == st_d["num_output_data_blocks"] and int(off[-1] + sz[-1]) + 5 == data_bytes
                assert st_d["num_input_records"] == n_entries
                one_table = {"value": round(in_bytes * world * args.steps / ot_s / 1e9, 4), "unit": "GB/s",
                             "ms_per_step": round(ot_s / args.steps * 1e3, 2), "output_file_bytes": int(data_bytes + meta_bytes),
                             "pieces": int(n_pieces), "data_blocks": int(len(off))}
            except Exception as ex:                          # never lose the bench line to the extra figure
                one_table = {"error": "%s: %s" % (type(ex).__name__, ex)}
            e2e = {"value": round(in_bytes * world * args.steps / sub_s / 1e9, 4), "unit": "GB/s",
                   "h2d_bytes_per_step": int(sres[-1][0]["h2d_bytes"]), "d2h_bytes_per_step": int(sres[-1][0]["d2h_bytes"]),
                   "ms_per_step": round(sub_s / args.steps * 1e3, 2), "pinned_inputs": all(ok for _, ok in pinned),
                   "output_file_bytes": int(sres[-1][1]), "verify_checksums": bool(args.verify),
                   "mode": "ybgpu_compact_files: %d key-range subcompactions (max_subcompactions=%d), %d in flight on private "
                           "streams, one output SST per range" % (sres[-1][2], args.subcompactions, args.in_flight),
                   "output_files": int(sres[-1][2]), "output_check": sub_check,
                   "gpu_ms_per_step": round(sres[-1][0]["gpu_seconds"] * 1e3, 2),
                   "one_table": one_table,
                   "single_job": single}
        for v, ok in pinned:
            if ok:
                cudart.cudaHostUnregister(v.ctypes.data)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_kind = measured_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    # dominant kernel = longest phase; its algorithmic bytes (DESIGN.md "Roofline accounting")
    alg_bytes = {
        "block_scan": in_bytes * 0.0,           # header walk only: not the dominant phase
        "decode": in_bytes,                     # must read every input entry once
        "partition": 0.0,
        "merge_filter": 0.0,
        "encode": out_bytes * 2.0,              # read each survivor once, write it once
    }
    names = pkg.PHASE_NAMES
    # The dominant KERNEL: decode and merge phases are one kernel each; the encode phase is ~25 launches of
    # which the block assembler k_encode_smem is timed separately.
    kernel_s = {"k_decode_all": phases[1], "k_merge_filter": phases[3], "k_encode_smem": enc_kernel_s}
    kernel_alg = {"k_decode_all": float(in_bytes),                  # must read every input entry once
                  "k_merge_filter": float(in_bytes + out_bytes),    # charged the whole path (it moves only keys)
                  "k_encode_smem": out_bytes * 2.0}                 # read each survivor once, write it once
    dom_kernel = max(kernel_s, key=lambda k: kernel_s[k])
    dom = max(range(5), key=lambda i: phases[i])
    dom_bytes = kernel_alg[dom_kernel]
    achieved = dom_bytes / kernel_s[dom_kernel] / 1e9 if kernel_s[dom_kernel] > 0 else 0.0
    # DRAM traffic of that kernel from the committed `ncu --set full` capture of this same command
    traffic = None
    try:
        if args.workload == "config2" and args.rows == DEFAULT_ROWS:
            prof = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_full_100m.json")))["kernels"]
            for name, kd in prof.items():
                if name.startswith(dom_kernel):
                    def gb(x):
                        v = float(x["value"]); return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[x["unit"]]
                    traffic = int(gb(kd["dram__bytes_read.sum"]) + gb(kd["dram__bytes_write.sum"]))
    except Exception:
        traffic = None
    pipeline_achieved = (in_bytes + out_bytes) / gpu_s / 1e9 if gpu_s > 0 else 0.0
    value = in_bytes * world * args.steps / total_s / 1e9
    line = {
        "metric": "compaction GB/s (input bytes merged)", "value": round(value, 3), "unit": "GB/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(total_s / args.steps * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": (WORKLOAD if args.rows == DEFAULT_ROWS else WORKLOAD + " (scaled to %d entries)" % args.rows)
                   if args.workload == "config2" else
                   "MVCC-heavy: 20 versions/key, history_cutoff drops 90%%, %d live keys, 1 GPU (8 input files)" % nrows,
                   "output_entries_per_gpu": int(stats[-1]["num_output_records"]),
                   "entries_per_gpu": int(n_entries), "input_raw_bytes_per_gpu": int(in_bytes),
                   "input_file_bytes_per_gpu": int(file_bytes), "tablets": world,
                   "parallelism": "tablet-per-GPU, no collective",
                   "output": "split SST: data blocks + CRC32C, multi-level index, DocKeyV3 bloom filter blocks (64 KB), properties, footer",
                   "l2": "inputs (%.1f GB) far larger than the 126 MB L2" % (file_bytes / 1e9)},
        "mkeys_per_s": round(n_entries * world * args.steps / total_s / 1e6, 2),
        "gpu_launches": int(launches),
        "clocks": clock_info,
        "roofline": {"bound": "hbm", "kernel": dom_kernel, "achieved": round(achieved, 1), "peak": hbm_peak, "unit": "GB/s",
                     "frac": round(achieved / hbm_peak, 4), "traffic": traffic, "peak_source": peak_kind,
                     "kernel_ms": round(kernel_s[dom_kernel] * 1e3, 3), "algorithmic_bytes_per_launch": int(dom_bytes),
                     "kernels_ms": {k: round(v * 1e3, 3) for k, v in kernel_s.items()},
                     "pipeline": {"algorithmic_bytes": int(in_bytes + out_bytes), "gpu_ms": round(gpu_s * 1e3, 3),
                                  "achieved": round(pipeline_achieved, 1), "frac": round(pipeline_achieved / hbm_peak, 4)},
                     "phase_ms": {names[i]: round(phases[i] * 1e3, 3) for i in range(5)}},
        "setup": {"generate_s": round(gen_s, 1)},
        "host_ms_per_step": {k: round(v / args.steps, 3) for k, v in host_ms.items()},
    }
    if e2e:
        line["e2e"] = e2e
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args)
    emit_json_line(line)
    if world > 1:
        dist.destroy_process_group()


def read_handles(pkg, sst):
    """Data-block handles of a generated SST, via the product's host meta reader."""
    off, sz, _ = pkg.sst_block_handles(sst.meta_view())
    return off, sz


if __name__ == "__main__":
    main()
