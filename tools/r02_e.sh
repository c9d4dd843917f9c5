#!/bin/bash
T=r02e
mkdir -p gpurun_out
python tools/diag_vs_reference.py config4_sibling > gpurun_out/${T}_diag_default.log 2>&1; grep -E "links|lightmap" gpurun_out/${T}_diag_default.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -5 gpurun_out/${T}_pytest.log
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02e_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','bake_wall_s','stage_ms')}); print(d['e2e']); print(d['counters'])
PY
