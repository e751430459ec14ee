#!/bin/bash
# dev: the bench at other per-GPU batch sizes (tile tails, fewer / more waves)
for b in "$@"; do
  python bench.py --batch $b --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('batch', d['config']['batch_per_gpu'], round(d['value']), round(d['ms_per_step'],2), d['e2e']['last_loss_acc'])"
done
