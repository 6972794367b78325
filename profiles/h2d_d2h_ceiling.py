#!/usr/bin/env python
"""Host<->device copy ceiling of the box: what the end-to-end compaction path can reach at N = 1, 2, 4, 8 GPUs.

  torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 profiles/h2d_d2h_ceiling.py

One launch measures everything (ranks >= n idle at the barriers while n ranks copy):
  placement "unbound"  pinned buffers allocated wherever the kernel put the process (what bench.py did in round 1)
  placement "bound"    after ybgpu_bind_thread_to_device(local_rank): CPU affinity + preferred memory node = the GPU's
                       NUMA node, then the buffers are allocated / first touched / pinned
  buffer kind          cudaHostAlloc (torch pin_memory) and cudaHostRegister on ordinary first-touched memory (the bench's
                       input files are registered, its output arena is cudaHostAlloc'ed)
  direction            h2d, d2h, both at once (two streams, the e2e path's steady state)
Copies are issued in 32 MB chunks like the engine's ChunkedCopyAsync. Rank 0 prints one JSON line per cell:
aggregate GB/s = n x bytes / max-over-ranks time (CUDA events around the copies, barrier before).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gb", type=float, default=6.0, help="bytes per direction per rank per repetition")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--chunk-mb", type=int, default=32)
    args = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module("yugabyte-db_b200")
    nbytes = int(args.gb * (1 << 30)) & ~0xfffff
    chunk = args.chunk_mb << 20
    dev_in = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dev_out = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dev_out.fill_(7)
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    cudart = torch.cuda.cudart()

    def where():
        try:
            cpus = sorted(os.sched_getaffinity(0))
            return {"cpu_now": int(open("/proc/self/stat").read().split()[38]), "allowed_cpus": len(cpus),
                    "first_cpu": cpus[0], "last_cpu": cpus[-1], "gpu_numa_node": pkg.lib().ybgpu_device_numa_node(local)}
        except Exception as e:
            return {"error": str(e)}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def copy_chunks(dst, src, stream):
        with torch.cuda.stream(stream):
            for off in range(0, nbytes, chunk):
                dst[off:off + chunk].copy_(src[off:off + chunk], non_blocking=True)

    def measure(n, direction, h_src, h_dst):
        active = rank < n
        barrier()
        t0 = time.perf_counter()
        if active:
            for _ in range(args.reps):
                if direction in ("h2d", "both"):
                    copy_chunks(dev_in, h_src, s_in)
                if direction in ("d2h", "both"):
                    copy_chunks(h_dst, dev_out, s_out)
            s_in.synchronize()
            s_out.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt if active else 0.0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        moved = nbytes * args.reps * (2 if direction == "both" else 1)
        return n * moved / float(t.item()) / 1e9

    if rank == 0:
        try:
            topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout
            sys.stderr.write(topo + "\n")
            sys.stderr.write(subprocess.run(["lscpu"], capture_output=True, text=True, timeout=30).stdout + "\n")
        except Exception as e:
            sys.stderr.write("topo: %s\n" % e)
    ns = [n for n in (1, 2, 4, 8) if n <= world]
    for placement in ("unbound", "bound"):
        if placement == "bound":
            node, ncpu = pkg.bind_thread_to_device(local)
        else:
            node, ncpu = -1, 0
        info = where()
        # cudaHostAlloc
        h_src = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        h_dst = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        h_src.fill_(1)
        h_dst.fill_(2)
        # cudaHostRegister on first-touched memory
        r_src_np = np.ones(nbytes, np.uint8)
        r_dst_np = np.ones(nbytes, np.uint8)
        ok1 = int(cudart.cudaHostRegister(r_src_np.ctypes.data, nbytes, 0)) == 0
        ok2 = int(cudart.cudaHostRegister(r_dst_np.ctypes.data, nbytes, 0)) == 0
        r_src, r_dst = torch.from_numpy(r_src_np), torch.from_numpy(r_dst_np)
        infos = [None] * world
        if world > 1:
            dist.all_gather_object(infos, dict(info, rank=rank, bound_node=node, bound_cpus=ncpu))
        else:
            infos = [dict(info, rank=0, bound_node=node, bound_cpus=ncpu)]
        if rank == 0:
            print(json.dumps({"placement": placement, "ranks": infos}), flush=True)
        for kind, (a, b) in (("cudaHostAlloc", (h_src, h_dst)), ("cudaHostRegister", (r_src, r_dst))):
            if kind == "cudaHostRegister" and not (ok1 and ok2):
                continue
            for n in ns:
                row = {"placement": placement, "buffers": kind, "n_gpus": n, "gb_per_dir_per_rank": round(nbytes * args.reps / 1e9, 2)}
                for direction in ("h2d", "d2h", "both"):
                    row[direction + "_gbs"] = round(measure(n, direction, a, b), 1)
                if rank == 0:
                    print(json.dumps(row), flush=True)
        if ok1:
            cudart.cudaHostUnregister(r_src_np.ctypes.data)
        if ok2:
            cudart.cudaHostUnregister(r_dst_np.ctypes.data)
        del h_src, h_dst, r_src, r_dst, r_src_np, r_dst_np
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
