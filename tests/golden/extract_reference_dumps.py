#!/usr/bin/env python
"""Extracts golden compaction sequences from the reference's own test sources into JSON fixtures.

    python tests/golden/extract_reference_dumps.py [/root/reference]

Source: src/yb/docdb/docdb-ttl-test.cc, TEST_P(DocDBTestWrapper, RedisCollectionTTLCompactionTest): one
initial ASSERT_DOC_DB_DEBUG_DUMP_STR_EQ followed by a chain of FullyCompactHistoryBefore(t[i]) /
ASSERT_DOC_DB_DEBUG_DUMP_STR_EQ pairs (t[i] = 1000 + 1000 i microseconds). The dumps are copied
verbatim; tests/test_reference_dumps.py parses and replays them (the reference tree is not read at
test time)."""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def extract(path, test_name):
    src = open(path).read()
    start = src.index("TEST_P(DocDBTestWrapper, %s)" % test_name)
    end = src.index("\nTEST_P(", start + 10)
    body = src[start:end]
    first_line = src[:start].count("\n") + 1
    events = []
    for m in re.finditer(r'ASSERT_DOC_DB_DEBUG_DUMP_STR_EQ\(\s*R"#\((.*?)\)#"\);|FullyCompactHistoryBefore\(t\[(\d+)\]\);', body, re.S):
        if m.group(2) is not None:
            events.append(("compact", 1000 + 1000 * int(m.group(2))))
        else:
            events.append(("dump", m.group(1)))
    assert events[0][0] == "dump"
    steps = []
    i = 1
    while i < len(events):
        assert events[i][0] == "compact" and events[i + 1][0] == "dump", "not a compaction / dump chain"
        steps.append({"cutoff_us": events[i][1], "expected": events[i + 1][1]})
        i += 2
    return {"source": "src/yb/docdb/docdb-ttl-test.cc:%d-%d %s" % (first_line, first_line + body.count("\n"), test_name),
            "initial": events[0][1], "steps": steps}


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out = extract(os.path.join(ref, "src/yb/docdb/docdb-ttl-test.cc"), "RedisCollectionTTLCompactionTest")
    with open(os.path.join(HERE, "redis_collection_ttl_compaction.json"), "w") as f:
        json.dump(out, f, indent=1)
        f.write("\n")
    print("%s: %d compactions" % (out["source"], len(out["steps"])))


if __name__ == "__main__":
    main()
