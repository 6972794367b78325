#!/usr/bin/env python
"""Generates tests/golden/compaction_golden.json with the ORACLE (test infrastructure): for a few
seeded compactions of the BASELINE config shapes it records the SHA-256 of the output files, the
KV-stream digest and the CompactionJobStats counters. The fixtures let the GPU parity tests check the
CUDA path against committed expectations (no oracle call in the comparison), and the CPU suite
checks that the oracle still reproduces them (guards the oracle itself against drift).

The oracle is pinned by the reference's own golden vectors (tests/test_oracle_*.py); these fixtures are
derived from it, not from a run of the reference binary (which cannot be built here).

    python tests/golden/make_golden.py          # rewrites compaction_golden.json
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import oracle_py as o  # noqa: E402

# name -> (generator config, table options, compaction parameters relative to cfg.base_micros)
CASES = {
    # BASELINE configs[0] shape: 2 files, minor compaction, cutoff = min
    "config1_two_sst_minor": dict(gen=dict(seed=1, num_rows=5000, cols=2, versions=2, num_files=2, value_len=32),
                                  table=dict(block_size=4096), params=dict(bottommost=False, other_min_ht=0)),
    # configs[1] shape scaled: 8-way major compaction, 256-byte values, bloom filter blocks
    "config2_eight_way_major": dict(gen=dict(seed=2, num_rows=20000, cols=1, versions=1, num_files=8, value_len=256),
                                    table=dict(block_size=32768, filter_policy=1, filter_block_size=65536), params=dict()),
    # configs[3] shape scaled: 20 versions per key, the history cutoff drops 90 %
    "config4_mvcc_heavy": dict(gen=dict(seed=4, num_rows=1500, cols=1, versions=20, num_files=8, value_len=64),
                               table=dict(block_size=8192, filter_policy=1, filter_block_size=4096),
                               params=dict(cutoff_micros=18 * 1000 + 500)),
    # tombstones + three_shared_parts output (YCQL tables)
    "tombstones_three_shared_parts": dict(gen=dict(seed=7, num_rows=4000, cols=3, versions=3, num_files=4, value_len=48, tombstone_per_1024=120),
                                          table=dict(block_size=4096, key_encoding=2), params=dict(cutoff_micros=1500)),
}


def run_case(case):
    cfg = o.GenConfig(**case["gen"])
    ssts = o.Sst.generate_all(cfg, o.TableOptions(block_size=case["table"].get("block_size", 32768)))
    kw = dict(case["params"])
    if "cutoff_micros" in kw:
        kw["cutoff_ht"] = o.ht_from_micros(cfg.base_micros + kw.pop("cutoff_micros"))
    res = o.compact(ssts, o.CompactionParams(**kw), o.TableOptions(**case["table"]))
    sst = res.sst()
    st = res.stats
    return ssts, kw, {
        "input_sha256": [hashlib.sha256(s.data).hexdigest() for s in ssts],
        "data_sha256": hashlib.sha256(sst.data).hexdigest(), "data_len": len(sst.data),
        "meta_sha256": hashlib.sha256(sst.meta).hexdigest(), "meta_len": len(sst.meta),
        "kv_hash": int(st.kv_hash),
        "num_input_records": int(st.num_input_records), "num_output_records": int(st.num_output_records),
        "num_dropped_hidden": int(st.num_dropped_hidden), "num_dropped_obsolete": int(st.num_dropped_obsolete),
        "num_dropped_feed": int(st.num_dropped_feed),
    }


def main():
    out = {name: run_case(case)[2] for name, case in CASES.items()}
    with open(os.path.join(HERE, "compaction_golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")
    print("wrote %d cases" % len(out))


if __name__ == "__main__":
    main()
