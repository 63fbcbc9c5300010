for st in 4 6 8; do
python bench.py --steps 20 --warmup 5 --no-cpu --no-matching --configs tum --streams $st > gpurun_out/r2_bench_s$st.json 2> gpurun_out/r2_bench_g.err; tail -c 100 gpurun_out/r2_bench_g.err
done
python bench.py --steps 20 --warmup 5 --no-cpu --no-matching --configs tum --streams 8 --batch 64 --batches-per-step 128 > gpurun_out/r2_bench_s8b64.json 2> gpurun_out/r2_bench_g.err; tail -c 100 gpurun_out/r2_bench_g.err
python bench.py --steps 20 --warmup 5 --no-cpu --no-matching --configs tum --streams 4 --batch 256 --batches-per-step 32 > gpurun_out/r2_bench_s4b256.json 2> gpurun_out/r2_bench_g.err; tail -c 100 gpurun_out/r2_bench_g.err
