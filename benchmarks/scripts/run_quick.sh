# quick A/B on one GPU: the C3 line without the CPU arm and the other configurations, then the tracker parity tests
tag=${1:-q}
mkdir -p gpurun_out
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-configs > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
for l in open("gpurun_out/${tag}_bench.json"):
    if l.startswith("{"):
        d=json.loads(l); print("${tag}", "tick", round(d["ms_per_step"],4), round(d["value"]), "e2e", round(d["e2e"]["value"]), d["stage_ms"])
PY
python -m pytest tests/test_gpu_tracker.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
