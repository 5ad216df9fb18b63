mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py > gpurun_out/bench_r1g.json 2> gpurun_out/bench_r1g.err; tail -2 gpurun_out/bench_r1g.err; cat gpurun_out/bench_r1g.json
python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_r1g.json 2> gpurun_out/bench_c4_r1g.err; tail -3 gpurun_out/bench_c4_r1g.err; cat gpurun_out/bench_c4_r1g.json
python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline --gate-impl 0 --chunks 1 > gpurun_out/bench_c4_g0.json 2> gpurun_out/bench_c4_g0.err; tail -3 gpurun_out/bench_c4_g0.err; cat gpurun_out/bench_c4_g0.json
