# round-2 GPU pass 8: long-attention side kernels after the in-flight-loads change
python -m pytest tests/test_round2_gpu.py tests/test_kernels_gpu.py tests/test_e2e_gpu.py -m gpu -q > gpurun_out/r02_t8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t8.log
timeout 600 python bench.py --model ViT-L/14 --scaling weak --batch 64 --classes 200 --steps 10 --warmup 3 --no-cpu-baseline --dump-prof gpurun_out/r02_prof_vitl14_b64.json > gpurun_out/r02_bench_vitl14_b64.json 2> gpurun_out/r02_bench_vitl14_b64.err
