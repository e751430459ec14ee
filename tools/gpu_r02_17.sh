for rep in 1 2; do
  for mode in 0 4 7; do
    LLC_TRAVERSAL=$mode timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_trav${mode}_$rep.json 2> gpurun_out/r02_bench_trav${mode}_$rep.err
  done
done
LLC_TRAVERSAL=4 python -m pytest tests/test_e2e_gpu.py tests/test_kernels_gpu.py -m gpu -q -x > gpurun_out/r02_t17_trav.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t17_trav.log
