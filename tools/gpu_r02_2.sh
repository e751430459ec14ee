# round-2 GPU pass 2: full GPU suite (stream-K on), benches at the strong-scaling shard sizes,
# ncu launch list and gemm2 full capture of the N=1 bench
python -m pytest tests -m gpu -q > gpurun_out/r02_t2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t2.log
for b in 256 128 64 32; do timeout 300 python bench.py --scaling weak --batch $b --steps 20 --warmup 3 --no-cpu-baseline --dump-prof gpurun_out/r02_prof2_b$b.json > gpurun_out/r02_bench2_b$b.json 2> gpurun_out/r02_bench2_b$b.err; done
timeout 600 python bench.py > gpurun_out/r02_bench2_default.json 2> gpurun_out/r02_bench2_default.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 330 -c 700 --csv --log-file gpurun_out/r02_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r02_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel --launch-skip 110 -c 10 -o gpurun_out/r02_gemm2_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r02_ncu_gemm2.log 2>&1
