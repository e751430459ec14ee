# final validation of the committed tree: full GPU suite, smoke, default bench + reference arm
python -m pytest tests -m gpu -q > gpurun_out/r02_tfinal.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tfinal.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_1gpu_ref.json 2> gpurun_out/r02_bench_1gpu_ref.err
timeout 600 python bench.py --mode eval --steps 5 --warmup 3 > gpurun_out/r02_bench_eval.json 2> gpurun_out/r02_bench_eval.err
timeout 600 python bench.py --peft both --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_both.json 2> gpurun_out/r02_bench_both.err
