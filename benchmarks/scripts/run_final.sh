# Round-end validation on one B200: the GPU test suite, smoke(), the driver's bench command, then one ncu capture of
# k_match (source-level, for profiles/ncu_hotspots.py) -- ncu last, after the un-profiled commands exited.
set -x
tag=${1:-r3e}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${tag}_smoke.log 2>&1; tail -2 gpurun_out/${tag}_smoke.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 300 gpurun_out/${tag}_bench.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'^k_match$' -s 112 -c 1 -f -o gpurun_out/prof_${tag}_match \
  python bench.py --steps 3 --warmup 3 --chunks 1 --no-cpu-baseline --no-configs > gpurun_out/${tag}_ncu_match.log 2>&1
ls -l gpurun_out/
