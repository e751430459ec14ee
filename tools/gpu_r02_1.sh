# round-2 GPU pass 1: everything with the whole-tile GEMM schedule first (baseline), then the
# stream-K schedule
LLC_STREAM_K=0 python -m pytest tests -m gpu -q > gpurun_out/r02_t1_nosk.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t1_nosk.log
for b in 256 32; do LLC_STREAM_K=0 python bench.py --scaling weak --batch $b --steps 20 --warmup 3 --no-cpu-baseline --dump-prof gpurun_out/r02_prof_nosk_b$b.json > gpurun_out/r02_bench_nosk_b$b.json 2> gpurun_out/r02_bench_nosk_b$b.err; done
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r02_t1_sk.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t1_sk.log
for b in 256 128 64 32; do timeout 300 python bench.py --scaling weak --batch $b --steps 20 --warmup 3 --no-cpu-baseline --dump-prof gpurun_out/r02_prof_b$b.json > gpurun_out/r02_bench_b$b.json 2> gpurun_out/r02_bench_b$b.err; done
timeout 300 python bench.py --mode eval --steps 5 --warmup 3 > gpurun_out/r02_bench_eval.json 2> gpurun_out/r02_bench_eval.err
timeout 300 python bench.py --peft both --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_both.json 2> gpurun_out/r02_bench_both.err
