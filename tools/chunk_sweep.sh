#!/bin/bash
# forward sweep time per step for different internal chunk sizes (candidates per kernel launch)
for c in 16384 32768 65536; do
  python bench.py --steps 3 --warmup 3 --chunk $c --no-cpu-baseline --no-e2e --no-grad-leg 2>/dev/null > gpurun_out/chunk_$c.json
  python - "$c" <<'PY'
import json, sys
c = sys.argv[1]
d = json.loads(open(f"gpurun_out/chunk_{c}.json").readline())
print("chunk", c, round(d["ms_per_step"], 2), [round(k["ms_per_step"], 2) for k in d["roofline"]["kernels"]])
PY
done
