#!/bin/bash
# round 2 (fourth session), GPU call 4: does a 20 MB pinned H2D copy per step slow the resident step down?
mkdir -p gpurun_out
for g in 0 1 2 0 1 2; do
echo -n "RS_DIAG_H2D=$g: "
RS_DIAG_H2D=$g timeout 400 python bench.py --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); f=d['roofline']['families']
g=lambda p: [v for k,v in f.items() if k.startswith(p)][0]['ms_per_step']
print('value %.2f ms/step (fwd %.2f bwd %.2f)  e2e %.2f ms/step' % (d['ms_per_step'], g('lstm_stack_forward'), g('lstm_stack_backward'), d['e2e']['ms_per_step']))"
done | tee gpurun_out/r02d_diag_h2d.txt
