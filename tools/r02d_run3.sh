#!/bin/bash
# round 2 (fourth session), GPU call 3: where the end-to-end step loses its 0.5 ms against the resident one
mkdir -p gpurun_out
for g in 0 1; do
echo "RS_PREFETCH_GATE=$g"
RS_PREFETCH_GATE=$g RS_BENCH_E2E_PHASES=1 timeout 400 python bench.py --no-cpu-baseline --steps 10 --warmup 3 2>&1 >/dev/null | grep "e2e phases"
done | tee gpurun_out/r02d_e2e_phases.txt
python - <<'PY' | tee -a gpurun_out/r02d_e2e_phases.txt
import json
PY
