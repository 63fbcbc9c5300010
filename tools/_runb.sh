python -m pytest tests/test_gpu_extractor.py -m gpu -x -q 2>&1 | tail -2
for g in 0 8 12 16 20 24; do
ORBCUDA_FAST_FMA=$g python bench.py --steps 10 --warmup 3 --no-cpu --no-matching --configs tum > gpurun_out/r2_bench_fa$g.json 2> gpurun_out/r2_bench_g.err; tail -c 100 gpurun_out/r2_bench_g.err; python -c "
import json
j=json.loads(open('gpurun_out/r2_bench_fa$g.json').read().strip().splitlines()[-1])
print($g, j['roofline']['stage_ms_per_batch']['fast_score'], j['roofline']['stage_ms_per_batch']['total'])"
done
ORBCUDA_FAST_FMA=0 python -m pytest tests/test_gpu_extractor.py -m gpu -x -q 2>&1 | tail -2
