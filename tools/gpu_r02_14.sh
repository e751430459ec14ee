python tools/attn_long_bench.py > gpurun_out/r02_attn_long.txt 2>&1
L=256 python tools/attn_long_bench.py >> gpurun_out/r02_attn_long.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_long --launch-skip 6 -c 2 -o gpurun_out/r02_attn_long -f python tools/attn_long_bench.py > gpurun_out/r02_attn_long_ncu.log 2>&1
python -m pytest tests/test_round2_gpu.py -m gpu -q -k "vitl14 or longer" > gpurun_out/r02_t14.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t14.log
