mkdir -p gpurun_out
python -m pytest tests/test_gpu_tracker.py tests/test_gpu_golden.py -x -q 2>&1 | tail -3
for cfg in "--chunks 1" "--chunks 1 --cosine-ctas 5" "--chunks 2" "--chunks 3" "--chunks 4"; do
  n=$(echo $cfg | tr -d ' -')
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline $cfg > gpurun_out/b_$n.json 2> gpurun_out/b_$n.err
  python - <<PY
import json
for l in open("gpurun_out/b_$n.json"):
    if l.startswith("{"):
        d=json.loads(l); s=d["stage_ms"]; print("$cfg", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["e2e"]["h2d_bytes_per_step"], {k: round(v,4) for k,v in s.items() if k!="pass"}, round(d["roofline"]["frac"],3))
PY
  tail -2 gpurun_out/b_$n.err
done
