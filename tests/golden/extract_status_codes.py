"""Extracts yb::Status::Code numbers from the reference header into tests/golden/status_codes_table.json.
Run in the build container (the GPU box has no /root/reference)."""
import json
import os
import re

SRC = "/root/reference/src/yb/util/status_codes.h"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "status_codes_table.json")

if __name__ == "__main__":
    tab = {m.group(1): int(m.group(3)) for m in re.finditer(r"YB_STATUS_CODE\((\w+),\s*(\w+),\s*(\d+),", open(SRC).read())}
    json.dump({"source": "src/yb/util/status_codes.h (reference 2.31.0.0-b0), extracted by tests/golden/extract_status_codes.py",
               "codes": tab}, open(OUT, "w"), indent=1)
    print(len(tab), "codes")
