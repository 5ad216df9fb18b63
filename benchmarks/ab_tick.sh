mkdir -p gpurun_out
python -m pytest tests/test_gpu_tracker.py -x -q -k "chunks" 2>&1 | tail -2
for cfg in "--chunks 1" "--chunks 2" "--chunks 4"; do
  n=$(echo $cfg | tr -d ' -')
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline $cfg > gpurun_out/b_$n.json 2> gpurun_out/b_$n.err
  python - <<PY
import json
for l in open("gpurun_out/b_$n.json"):
    if l.startswith("{"):
        d=json.loads(l); s=d["stage_ms"]; print("$cfg", round(d["value"]), "enq", d["host_enqueue_ms_per_step"], "e2e", d["e2e"])
PY
  tail -2 gpurun_out/b_$n.err
done
