python -m pytest tests/test_gpu_extractor.py -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 --no-cpu --no-matching --configs tum > gpurun_out/r2_bench_g.json 2> gpurun_out/r2_bench_g.err; tail -c 100 gpurun_out/r2_bench_g.err; python -c "
import json
j=json.loads(open('gpurun_out/r2_bench_g.json').read().strip().splitlines()[-1])
print(j['roofline']['stage_ms_per_batch'], j['single_frame_latency_ms'])"
