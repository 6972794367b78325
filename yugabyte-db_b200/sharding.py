"""Multi-GPU placement of compactions.

Tablets are independent RocksDB instances (tablet/tablet.cc:1173-1252); the reference runs their
compactions as independent PriorityThreadPool tasks (rocksdb/db/db_impl.cc:3029). Here a tablet's
compaction runs entirely on one GPU, so placement is a bin-packing problem with no data-path
collective: size-balanced greedy (largest first onto the least loaded GPU).

A single oversized tablet is key-range sharded instead (plan_key_ranges): splitter keys on DocKey
boundaries chosen from the inputs' index separators — the GPU analogue of
CompactionJob::GenSubcompactionBoundaries (rocksdb/db/compaction_job.cc:409-519) — after which
each rank owns one key range and input slices are exchanged once (all_to_all over NCCL).
"""
from bisect import bisect_left, bisect_right


def assign_tablets(sizes, n_gpus):
    """sizes[i] = input bytes of tablet i. Returns per-GPU lists of tablet indices (LPT greedy)."""
    if n_gpus < 1:
        raise ValueError("n_gpus must be >= 1")
    order = sorted(range(len(sizes)), key=lambda i: (-sizes[i], i))
    loads = [0] * n_gpus
    out = [[] for _ in range(n_gpus)]
    for i in order:
        g = min(range(n_gpus), key=lambda j: (loads[j], j))
        out[g].append(i)
        loads[g] += sizes[i]
    for lst in out:
        lst.sort()
    return out


def tablets_for_rank(sizes, rank, world):
    return assign_tablets(sizes, world)[rank]


def docdb_row_prefix(user_key):
    """Length of the row-group prefix of a DocDB user key for the simple shapes the splitter
    sampler sees (hash or range DocKeys without cotable ids): up to and including the '!' that ends
    the range group. Falls back to the whole key."""
    k = user_key
    i = 0
    if k[:1] == b"G":
        i = 3
        # hashed group
        i = _skip_group(k, i)
        if i < 0:
            return len(k)
    i = _skip_group(k, i)
    return len(k) if i < 0 else i


def _skip_group(k, i):
    n = len(k)
    while i < n:
        t = k[i]
        if t == 0x21:                       # '!'
            return i + 1
        if t in (0x53, 0x5c, 0x61, 0x5d):   # zero-terminated strings (ascending / descending)
            e = 0x00 if t in (0x53, 0x5c) else 0xff
            j = i + 1
            while True:                     # terminator = e e ; e (e^1) is an escaped e byte
                j = k.find(bytes([e]), j)
                if j < 0 or j + 1 >= n:
                    return -1
                if k[j + 1] == e:
                    break
                j += 2
            i = j + 2
        elif t in (0x48, 0x65, 0x4f, 0x67):  # 32-bit
            i += 5
        elif t in (0x49, 0x62, 0x55, 0x6a, 0x5b, 0x73, 0x63, 0x44, 0x4c):  # 64-bit
            i += 9
        elif t in (0x24, 0x7c, 0x54, 0x46):  # value-less
            i += 1
        else:
            return -1
    return -1


def plan_key_ranges(separator_keys, weights, n_ranges):
    """Choose n_ranges-1 splitter user keys (row-group aligned) so the weighted mass between
    consecutive splitters is balanced. separator_keys: sorted user keys sampled from all inputs
    (e.g. last key of every data block, from the index); weights[i] = bytes the sample stands for.
    Returns the sorted list of splitters (may be shorter if there are few distinct rows)."""
    if n_ranges <= 1 or not separator_keys:
        return []
    total = float(sum(weights))
    target = total / n_ranges
    out = []
    acc = 0.0
    nxt = target
    for k, w in zip(separator_keys, weights):
        acc += w
        if acc >= nxt and len(out) < n_ranges - 1:
            s = k[:docdb_row_prefix(k)]
            if not out or s > out[-1]:
                out.append(s)
                nxt = target * (len(out) + 1)
    return out


def range_of_rank(splitters, rank):
    """[lower, upper) user-key bounds owned by `rank` (b'' = unbounded)."""
    lo = splitters[rank - 1] if rank > 0 and rank - 1 < len(splitters) else b""
    hi = splitters[rank] if rank < len(splitters) else b""
    if rank > len(splitters):
        return None            # fewer ranges than ranks: this rank has nothing to do
    return lo, hi


def blocks_for_range(first_keys, last_keys, lo, hi):
    """Indices [a, b) of the data blocks of one file that can contain user keys in [lo, hi).
    first_keys/last_keys: per-block boundary user keys (sorted)."""
    a = 0 if not lo else bisect_left(last_keys, lo)          # first block whose last key >= lo
    b = len(first_keys) if not hi else bisect_left(first_keys, hi)   # blocks whose first key < hi
    return a, max(a, b)


def run_sharded(items, rank, world, run_fn, sizes=None):
    """Runs run_fn(item_index) for the items placed on this rank; returns {index: result}."""
    sizes = sizes if sizes is not None else [1] * len(items)
    mine = tablets_for_rank(sizes, rank, world)
    return {i: run_fn(i) for i in mine}
