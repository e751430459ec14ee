#!/bin/bash
# dev: head kernels in the step for the row-rotation variants (LLC_HEAD_ROT)
for r in 0 1 2 3 1 0; do
  LLC_HEAD_ROT=$r python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('rot $r', round(d['ms_per_step'],3), round(d['kernel_breakdown']['head']['ms_per_step'],3))"
done
