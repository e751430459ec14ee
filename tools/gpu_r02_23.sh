#!/bin/bash
# rows-per-warp variants of the LayerNorm + rank-r row product kernel (were selected by bits 3-4 of
# llc_set_traversal; measured, rejected and removed again - DESIGN.md section 8 has the numbers)
cd /root/repo; mkdir -p gpurun_out
for mode in 0 8 16 0 8; do
  LLC_TRAVERSAL=$mode timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_breakdown']['ln_fwd']; print('mode $mode', round(d['value']), round(d['ms_per_step'],3), 'ln_fwd ms', round(k['ms_per_step'],3), 'GB/s', round(k['gbs']))"
done
LLC_TRAVERSAL=8 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_e2e_gpu.py -q -k "ln or golden or step" 2>&1 | tail -2
