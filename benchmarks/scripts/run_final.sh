set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r3b_pytest.log 2>&1; tail -3 gpurun_out/r3b_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r3b_smoke.log 2>&1; tail -2 gpurun_out/r3b_smoke.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r3b_bench.json 2> gpurun_out/r3b_bench.err; tail -c 600 gpurun_out/r3b_bench.json
