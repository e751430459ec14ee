# round-2 GPU pass 7: long-sequence attention (ViT-L/14), full suite, ViT-L/14 bench
python -m pytest tests -m gpu -q > gpurun_out/r02_t7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t7.log
timeout 600 python bench.py --model ViT-L/14 --scaling weak --batch 64 --classes 200 --steps 10 --warmup 3 --no-cpu-baseline --dump-prof gpurun_out/r02_prof_vitl14_b64.json > gpurun_out/r02_bench_vitl14_b64.json 2> gpurun_out/r02_bench_vitl14_b64.err
timeout 600 python bench.py > gpurun_out/r02_bench7_default.json 2> gpurun_out/r02_bench7_default.err
