# round-2 GPU pass 3: full suite again (MaPLe, calibrated two-tower tolerances), true kernel
# durations at 32 images per GPU (ncu launch list), stream-K A/B at 32 and 128
python -m pytest tests -m gpu -q > gpurun_out/r02_t3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t3.log
for b in 32 128; do
  timeout 300 python bench.py --scaling weak --batch $b --steps 30 --warmup 5 --no-cpu-baseline --dump-prof gpurun_out/r02_prof3_b$b.json > gpurun_out/r02_bench3_b$b.json 2> gpurun_out/r02_bench3_b$b.err
  LLC_STREAM_K=0 timeout 300 python bench.py --scaling weak --batch $b --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench3_nosk_b$b.json 2> gpurun_out/r02_bench3_nosk_b$b.err
done
timeout 300 python bench.py --scaling weak --batch 32 --steps 30 --warmup 5 --no-cpu-baseline --no-graph > gpurun_out/r02_bench3_nograph_b32.json 2> gpurun_out/r02_bench3_nograph_b32.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 330 -c 700 --csv --log-file gpurun_out/r02_ncu_launches_b32.csv python bench.py --scaling weak --batch 32 --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r02_ncu_launches_b32.log 2>&1
timeout 600 python bench.py --model ViT-L/14 --scaling weak --batch 64 --classes 200 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench3_vitl14_b64.json 2> gpurun_out/r02_bench3_vitl14_b64.err
