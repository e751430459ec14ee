#!/bin/bash
# adapter-clip: bench lines (both towers / image only) + per-launch records
cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py --method adapter --peft both --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_adapter_both.json 2> gpurun_out/bench_adapter_both.err; tail -c 4000 gpurun_out/bench_adapter_both.json; tail -5 gpurun_out/bench_adapter_both.err
timeout 600 python bench.py --method adapter --peft image --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_adapter_image.json 2> gpurun_out/bench_adapter_image.err; tail -c 600 gpurun_out/bench_adapter_image.json; tail -5 gpurun_out/bench_adapter_image.err
timeout 600 python bench.py --method adapter --peft both --batch 32 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_adapter_b32.json 2> gpurun_out/bench_adapter_b32.err; tail -c 400 gpurun_out/bench_adapter_b32.json
