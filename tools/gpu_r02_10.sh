# round-2 GPU pass 10: traversal order of the LayerNorm kernels (A/B, interleaved to cancel drift)
for rep in 1 2; do
  for mode in 0 3 1 2; do
    LLC_TRAVERSAL=$mode timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --dump-prof gpurun_out/r02_prof_trav${mode}_$rep.json > gpurun_out/r02_bench_trav${mode}_$rep.json 2> gpurun_out/r02_bench_trav${mode}_$rep.err
  done
done
LLC_TRAVERSAL=3 python -m pytest tests/test_e2e_gpu.py tests/test_kernels_gpu.py -m gpu -q -x > gpurun_out/r02_t10_trav.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t10_trav.log
