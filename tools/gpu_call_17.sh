#!/bin/bash
# Round 2, call 17 (1 GPU, 5.1 GPU-minutes left): the final tree — whole device suite (incl. the FSR fixtures and the G-buffer against
# the rasterised prepass.wgsl), smoke(), and the default bench line.
mkdir -p gpurun_out
O=gpurun_out
T=r2c17
echo "== device suite"
timeout -s INT 150 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/${T}_pytest.txt 2>&1; tail -6 $O/${T}_pytest.txt
echo "== smoke"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== default bench line"
timeout 120 python bench.py > $O/${T}_default.json 2> $O/${T}_default.err; tail -c 300 $O/${T}_default.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c17_default.json").read().strip().splitlines()[-1])
print("ms/frame", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], "roofline", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], "cpu", d["cpu_baseline"]["value"])
PY
