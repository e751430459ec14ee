# round-2 GPU pass 4: stream-K A/B per shape, early PDL trigger A/B on the step
python tools/sk_bench.py > gpurun_out/r02_streamk_ab.txt 2> gpurun_out/r02_streamk_ab.err
for b in 32 256; do
  for trig in 0 1; do
    LLC_PDL_TRIGGER=$trig timeout 300 python bench.py --scaling weak --batch $b --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench4_trig${trig}_b$b.json 2> gpurun_out/r02_bench4_trig${trig}_b$b.err
  done
done
LLC_PDL_TRIGGER=1 python -m pytest tests -m gpu -q -x > gpurun_out/r02_t4_trig.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t4_trig.log
