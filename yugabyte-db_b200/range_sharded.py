"""Key-range sharded compaction of one oversized tablet across ranks (BASELINE config 5).

Every rank starts with some of the tablet's input SSTs (files are staged round-robin). Steps:
  1. sample: the index separators of the local files, weighted by block bytes (host, tiny);
  2. plan: all ranks agree on world-1 row-aligned splitter keys (sharding.plan_key_ranges);
  3. exchange: for every (local file, destination rank) the contiguous run of data blocks that can
     hold keys of the destination's range is sent once — ONE all_to_all over NCCL (device tensors);
  4. compact: each rank runs an ordinary GPU job over the slices it received with
     range_lower/range_upper set, so entries of boundary blocks outside its range are invisible;
  5. the per-rank outputs are the compaction's output files, in range order (the reference adds
     every sub-output in order too: rocksdb/db/compaction_job.cc:1128-1131).
The seqno-zeroing exception key (Compaction::GetLargestUserKey) is global and passed to all ranks.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import sharding
from .binding import GpuCompactionJob, sst_block_handles, sst_separators


def _user(k):
    return k[:-8]


def plan(local_metas, world, group=None):
    """Returns (splitters, per-file (offsets, sizes, separators), global largest separator user key)."""
    files = []
    samples = []
    for meta in local_metas:
        off, sz, enc = sst_block_handles(meta)
        seps = sst_separators(meta)
        files.append((off, sz, seps, enc))
        samples += [(_user(k), int(s) + 5) for k, s in zip(seps, sz)]
    gathered = [None] * world
    dist.all_gather_object(gathered, samples, group=group)
    flat = sorted(s for part in gathered for s in part)
    splitters = sharding.plan_key_ranges([k for k, _ in flat], [w for _, w in flat], world)
    return splitters, files


def slices_for(files, splitters, world):
    """slices[dst] = list of (file_idx, block_a, block_b) to send to dst."""
    out = [[] for _ in range(world)]
    for fi, (off, sz, seps, enc) in enumerate(files):
        useps = [_user(k) for k in seps]
        for dst in range(world):
            r = sharding.range_of_rank(splitters, dst)
            if r is None:
                continue
            lo, hi = r
            # block i holds user keys in (useps[i-1], useps[i]] (separator >= last key of the block)
            a = 0
            if lo:
                while a < len(useps) and useps[a] < lo:
                    a += 1
            b = len(useps)
            if hi:
                b = a
                while b < len(useps) and (b == 0 or useps[b - 1] < hi):
                    b += 1
            if b > a:
                out[dst].append((fi, a, b))
    return out


def exchange(local_datas, files, slices, rank, world, device, group=None):
    """One all_to_all of the block byte ranges. local_datas: list of uint8 tensors (on `device`).
    Returns list of (tensor, offsets, sizes, key_encoding) received, one per (source rank, file)."""
    send_chunks, send_meta = [], [[] for _ in range(world)]
    send_sizes = [0] * world
    for dst in range(world):
        for (fi, a, b) in slices[dst]:
            off, sz, seps, enc = files[fi]
            start = int(off[a])
            end = int(off[b - 1] + sz[b - 1]) + 5
            # 16-byte aligned framing so every received slice starts on an aligned address
            pad = (-(end - start)) % 16
            send_chunks.append((dst, local_datas[fi][start:end], pad))
            send_meta[dst].append(((off[a:b] - np.uint64(start)).astype(np.uint64), sz[a:b].copy(), enc, end - start + pad + 32))
            send_sizes[dst] += end - start + pad + 32
    meta_in = [None] * world
    dist.all_gather_object(meta_in, send_meta, group=group)          # meta_in[src][dst] = list of slice metas
    recv_sizes = [sum(m[3] for m in meta_in[src][rank]) for src in range(world)]
    send_buf = torch.zeros(sum(send_sizes), dtype=torch.uint8, device=device)
    pos_by_dst = [sum(send_sizes[:d]) for d in range(world)]
    for dst, chunk, pad in send_chunks:
        p = pos_by_dst[dst] + 16                                    # 16 zero bytes in front of every slice
        send_buf[p:p + chunk.numel()] = chunk
        pos_by_dst[dst] += chunk.numel() + pad + 32
    recv_buf = torch.zeros(sum(recv_sizes) + 64, dtype=torch.uint8, device=device)
    dist.all_to_all_single(recv_buf[:sum(recv_sizes)], send_buf, recv_sizes, send_sizes, group=group)
    received = []
    p = 0
    for src in range(world):
        for (off, sz, enc, framed) in meta_in[src][rank]:
            received.append((recv_buf, p + 16, framed - 32, off, sz, enc))
            p += framed
    return received, int(send_buf.numel()), int(sum(recv_sizes))


def compact(local_ssts, rank, world, device_index, largest_user_key, group=None, **job_kwargs):
    """local_ssts: list of (meta uint8 ndarray, data uint8 ndarray). Returns (job, my_range, stats dict)."""
    device = torch.device("cuda", device_index)
    splitters, files = plan([m for m, _ in local_ssts], world, group)
    slices = slices_for(files, splitters, world)
    datas = [torch.from_numpy(np.ascontiguousarray(d)).to(device) for _, d in local_ssts]
    received, sent_bytes, recv_bytes = exchange(datas, files, slices, rank, world, device, group)
    my = sharding.range_of_rank(splitters, rank)
    job = None
    if my is not None:
        lo, hi = my
        job = GpuCompactionJob(device=device_index, largest_user_key=largest_user_key, range_lower=lo, range_upper=hi,
                               cuda_stream=torch.cuda.current_stream().cuda_stream, **job_kwargs)
        torch.cuda.current_stream().synchronize()
        for (buf, start, length, off, sz, enc) in received:
            job.add_input_device(buf.data_ptr() + start, length, off, sz, key_encoding=enc)
        job._keepalive = (received, datas)
        job.run()
    return job, my, {"sent_bytes": sent_bytes, "recv_bytes": recv_bytes, "splitters": splitters}
