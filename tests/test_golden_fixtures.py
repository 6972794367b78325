"""Committed golden fixtures (tests/golden/compaction_golden.json, made by tests/golden/make_golden.py):
the CPU suite checks the oracle still reproduces them; the GPU suite checks the CUDA path against them
directly — inputs from the product's own generator, outputs hashed, no oracle call in the comparison."""
import hashlib
import importlib
import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as mg  # noqa: E402

GOLDEN = json.load(open(os.path.join(HERE, "golden", "compaction_golden.json")))


@pytest.mark.parametrize("name", sorted(mg.CASES))
def test_oracle_reproduces_golden(name):
    assert mg.run_case(mg.CASES[name])[2] == GOLDEN[name]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(mg.CASES))
def test_gpu_matches_golden(name):
    pkg = importlib.import_module("yugabyte-db_b200")
    case, gold = mg.CASES[name], GOLDEN[name]
    cfg = pkg.GenConfig(**case["gen"])
    ssts = pkg.generate_ssts(cfg, block_size=case["table"].get("block_size", 32768))
    assert [hashlib.sha256(s.data_view().tobytes()).hexdigest() for s in ssts] == gold["input_sha256"]
    kw = dict(case["params"])
    if "cutoff_micros" in kw:
        kw["cutoff_ht"] = (cfg.base_micros + kw.pop("cutoff_micros")) << 12
    t = case["table"]
    job = pkg.GpuCompactionJob(block_size=t.get("block_size", 32768), output_key_encoding=t.get("key_encoding", 1),
                               filter_policy=t.get("filter_policy", 0), filter_block_size=t.get("filter_block_size", 65536), **kw)
    for s in ssts:
        job.add_input_sst(s.meta_view(), s.data_view())
    st = job.run()
    data, meta = job.fetch_output()
    assert (len(data), hashlib.sha256(data.tobytes()).hexdigest()) == (gold["data_len"], gold["data_sha256"])
    assert (len(meta), hashlib.sha256(meta.tobytes()).hexdigest()) == (gold["meta_len"], gold["meta_sha256"])
    assert job.digest() == gold["kv_hash"]
    assert (st.num_input_records, st.num_output_records) == (gold["num_input_records"], gold["num_output_records"])
    assert (st.num_record_drop_hidden, st.num_record_drop_obsolete, st.num_record_drop_feed) == (
        gold["num_dropped_hidden"], gold["num_dropped_obsolete"], gold["num_dropped_feed"])
