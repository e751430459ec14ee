python tools/attn_long_bench.py > gpurun_out/r02_attn_long_v2.txt 2>&1
python -m pytest tests/test_round2_gpu.py -m gpu -q -k "vitl14 or longer" > gpurun_out/r02_t15.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t15.log
timeout 600 python bench.py --model ViT-L/14 --scaling weak --batch 64 --classes 200 --steps 10 --warmup 3 --no-cpu-baseline --dump-prof gpurun_out/r02_prof_vitl14_b64.json > gpurun_out/r02_bench_vitl14_b64.json 2> gpurun_out/r02_bench_vitl14_b64.err
timeout 300 python bench.py --scaling weak --batch 32 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench15_b32.json 2> gpurun_out/r02_bench15_b32.err
