"""Diagnostic for the GPU Snappy output path: compares, block by block, the engine's compressed table with the
oracle's (same encoder) and prints where they part. Usage: python profiles/snappy_out_debug.py"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import oracle_py as o          # noqa: E402
import workloads as w          # noqa: E402
import test_gpu_parity as T    # noqa: E402

pkg = importlib.import_module("yugabyte-db_b200")


def case(seed):
    if seed == 0:
        return T._phrase_runs(900, 3, 300), 1024, 1, w.param_grid()[0]
    if seed == 1:
        return T._phrase_runs(901, 4, 1500, random_every=120), 4096, 2, w.param_grid()[0]
    if seed == 2:
        return T._phrase_runs(902, 3, 4000, vmax=400, random_every=700), 32768, 1, w.param_grid()[2]
    runs = T._phrase_runs(903, 2, 60)
    big = [(k, b"S" + (bytes(range(256)) * 700)[:150000 + 7 * i] if i % 9 == 4 and v[:1] == b"S" else
            (b"S" + b"\0" * (3000 + i) if i % 9 == 7 and v[:1] == b"S" else v)) for i, (k, v) in enumerate(runs[0])]
    return [big, runs[1]], 2048, 1, w.param_grid()[0]


for variant in (0, 1, 2):
    os.environ["YBGPU_SNAPC_VARIANT"] = str(variant)
    print("== encoder variant", variant)
    for seed in range(4):
        runs, bs, enc, kw = case(seed)
        ssts = [o.Sst.build(r, o.TableOptions(block_size=bs)) for r in runs if r]
        exp = o.compact(ssts, o.CompactionParams(**T.okw(kw)), o.TableOptions(block_size=bs, key_encoding=enc, compression=1))
        try:
            job = T.gpu_compact(pkg, ssts, block_size=bs, output_key_encoding=enc, output_compression=1, **kw)
            data, meta = job.fetch_output()
        except Exception as e:                       # noqa: BLE001
            print("seed", seed, "FAILED:", repr(e))
            continue
        data, meta = data.tobytes(), meta.tobytes()
        ed, em = bytes(exp.sst().data), bytes(exp.sst().meta)
        print("seed", seed, "data equal", data == ed, "meta equal", meta == em, "sizes", len(data), len(ed), "flags", job.stats().path_flags)
        if data == ed:
            continue
        eoff, esz = exp.sst().block_handles()
        try:
            goff, gsz, _ = pkg.sst_block_handles(np.frombuffer(meta, np.uint8))
        except Exception as e:                       # noqa: BLE001
            print("  engine metadata unreadable:", repr(e))
            goff, gsz = [], []
        shown = 0
        for b in range(min(len(eoff), len(goff))):
            eo, es, go, gs = int(eoff[b]), int(esz[b]), int(goff[b]), int(gsz[b])
            eb, gb = ed[eo:eo + es + 5], data[go:go + gs + 5]
            if eb != gb or eo != go:
                first = next((i for i in range(min(len(eb), len(gb))) if eb[i] != gb[i]), min(len(eb), len(gb)))
                print("  block", b, "oracle off/size/type", eo, es, eb[es] if len(eb) > es else None, "engine", go, gs, gb[gs] if len(gb) > gs else None,
                      "first differing byte", first, "oracle", eb[max(0, first - 4):first + 12].hex(), "engine", gb[max(0, first - 4):first + 12].hex())
                shown += 1
                if shown >= 6:
                    break
        print("  blocks", len(eoff), len(goff))

# ---- what the encoder costs: device time of k_snappy_compress / k_snappy_gather (stats slots 6 / 7) and of the whole
# output phase, on the bench's synthetic shape (random values: nothing to find, every position is visited), on short values
# (keys dominate) and on tombstone-heavy tables
shapes = ((256, 4000000, 0), (16, 8000000, 0), (64, 6000000, 700))
for value_len, rows, tomb in shapes:
    try:
        cfg = pkg.GenConfig(seed=7, num_rows=rows, cols=1, versions=1, num_files=4, value_len=value_len, tombstone_per_1024=tomb, tombstone_newest=1)
        ssts = pkg.generate_ssts(cfg)
        base = None
        for variant in (None, 0, 1, 2, 0, 1, 2):
            if variant is not None:
                os.environ["YBGPU_SNAPC_VARIANT"] = str(variant)
            job = pkg.GpuCompactionJob(output_compression=0 if variant is None else 1, retain_delete_markers=True)
            for s_ in ssts:
                job.add_input_sst(s_.meta_view(), s_.data_view())
            job.run()
            st = job.stats()
            d, m = job.fetch_output()
            if variant is None:
                base = (st.phase_seconds[4] * 1e3, d.size)
                print("value_len %d tombstones %d/1024: %d bytes of blocks, output phase %.3f ms without compression" % (value_len, tomb, d.size, base[0]))
            else:
                cms, gms = st.phase_seconds[6] * 1e3, st.phase_seconds[7] * 1e3
                print("  variant %d: encoder %.3f ms (%.1f GB/s of block bytes), gather %.3f ms, output phase %.3f ms; stored %d bytes (%.3f)"
                      % (variant, cms, base[1] / max(cms, 1e-6) / 1e6, gms, st.phase_seconds[4] * 1e3, d.size, d.size / base[1]))
            del job, d, m
    except Exception as e:                       # noqa: BLE001
        print("timing value_len", value_len, "FAILED:", repr(e))
