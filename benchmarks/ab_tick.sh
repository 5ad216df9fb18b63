mkdir -p gpurun_out
python -m pytest tests/test_gpu_gallery.py tests/test_gpu_tracker.py tests/test_gpu_golden.py -x -q 2>&1 | tail -2
for cfg in "--chunks 1" "--chunks 2" "--workload c4 --steps 10 --chunks 2"; do
  n=$(echo $cfg | tr -d ' -')
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline $cfg > gpurun_out/b_$n.json 2> gpurun_out/b_$n.err
  python - <<PY
import json
for l in open("gpurun_out/b_$n.json"):
    if l.startswith("{"):
        d=json.loads(l); s=d["stage_ms"]; print("$cfg", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(v,4) for k,v in s.items() if k!="pass"})
PY
  tail -2 gpurun_out/b_$n.err
done
