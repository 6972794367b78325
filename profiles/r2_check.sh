#!/bin/bash
# Round-2 GPU check (one B200): parity suite, then the bench at 8 M and 100 M entries with per-phase timings, then —
# if asked — the A/B switches that select the round-1 kernels (same binary).
#   gpurun --timeout 1500 -- 'bash profiles/r2_check.sh [tests|small|full|ab]'
mkdir -p gpurun_out
what=${1:-full}
if [ "$what" = "tests" ] || [ "$what" = "full" ] || [ "$what" = "res" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
  tail -15 gpurun_out/r2_pytest.log
fi
if [ "$what" = "small" ] || [ "$what" = "full" ] || [ "$what" = "res" ]; then
  timeout 300 python bench.py --rows 8000000 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-configs > gpurun_out/r2_bench_8m.json 2> gpurun_out/r2_bench_8m.err
  echo "bench 8m rc=$?"; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_bench_8m.json"))
    print("8M:", d["value"], "GB/s", d["ms_per_step"], "ms", d["roofline"]["phase_ms"], d["roofline"]["kernels_ms"], "no-verify", d["value_no_verify"]["ms_per_step"])
except Exception as e:
    print("8M bench unreadable:", e)
PY
fi
if [ "$what" = "ab" ]; then
  for env in "" "YBGPU_NO_INGEST=1" "YBGPU_ENC_V3=1" "YBGPU_NO_INGEST=1 YBGPU_ENC_V3=1"; do
    env $env timeout 300 python bench.py --rows 8000000 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-configs > gpurun_out/r2_ab.json 2> gpurun_out/r2_ab.err
    python - "$env" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/r2_ab.json"))
    print("[%s]" % sys.argv[1], d["ms_per_step"], "ms", d["roofline"]["phase_ms"], d["roofline"]["kernels_ms"])
except Exception as e:
    print("[%s] unreadable: %s" % (sys.argv[1], e))
PY
  done
fi
if [ "$what" = "res" ]; then
  timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extra-configs > gpurun_out/r2_bench_100m_res.json 2> gpurun_out/r2_bench_100m_res.err
  echo "bench 100m (resident only) rc=$?"; tail -c 300 gpurun_out/r2_bench_100m_res.err
  python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_bench_100m_res.json"))
    print("100M:", d["value"], "GB/s", d["ms_per_step"], "ms/step; no-verify", d["value_no_verify"]["ms_per_step"], "ms")
    print(" phases", d["roofline"]["phase_ms"], "kernels", d["roofline"]["kernels_ms"], "pipeline frac", d["roofline"]["pipeline"]["frac"])
except Exception as e:
    print("100M bench unreadable:", e)
PY
fi
if [ "$what" = "full" ]; then
  timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_100m.json 2> gpurun_out/r2_bench_100m.err
  echo "bench 100m rc=$?"; tail -c 600 gpurun_out/r2_bench_100m.err
  python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_bench_100m.json"))
    print("100M:", d["value"], "GB/s", d["ms_per_step"], "ms/step; no-verify", d["value_no_verify"]["ms_per_step"], "ms")
    print(" phases", d["roofline"]["phase_ms"], "kernels", d["roofline"]["kernels_ms"], "pipeline frac", d["roofline"]["pipeline"]["frac"])
    e = d.get("e2e", {})
    print(" e2e one table", e.get("value"), "GB/s", e.get("ms_per_step"), "ms; range files", e.get("range_files", {}).get("value"), "single", e.get("single_job", {}).get("value"), "ceiling", e.get("pcie_ceiling_gbs"))
    print(" configs", {k: (v.get("value"), v.get("ms_per_step"), v.get("error")) for k, v in d.get("configs", {}).items()})
    print(" parity", d.get("parity_check"))
except Exception as e:
    print("100M bench unreadable:", e)
PY
fi
