# multi-GPU pass: usage  bash tools/gpu_r02_multi.sh N
N=$1
if [ "$N" = "2" ]; then
  python -m pytest tests/test_multigpu_gpu.py -m gpu -q > gpurun_out/r02_tmulti.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tmulti.log
fi
PORT=$((29500 + N))
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+20)) bench.py --gpus $N --steps 20 --warmup 5 --scaling weak > gpurun_out/r02_bench_${N}gpu_weak.json 2> gpurun_out/r02_bench_${N}gpu_weak.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+40)) bench.py --gpus $N --steps 1 --warmup 1 --impl reference --cpu-batch 8 > gpurun_out/r02_bench_${N}gpu_ref.json 2> gpurun_out/r02_bench_${N}gpu_ref.err
