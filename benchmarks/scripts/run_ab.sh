# A/B of bench.py knobs on one GPU: each argument is one quoted option string, "" = defaults
mkdir -p gpurun_out
tag=$1; shift
i=0
for cfg in "$@"; do
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-configs $cfg > gpurun_out/${tag}_$i.json 2> gpurun_out/${tag}_$i.err
  python - <<PY
import json
for l in open("gpurun_out/${tag}_$i.json"):
    if l.startswith("{"):
        d=json.loads(l); s=d["stage_ms"]; print("[$cfg]", "tick", round(d["ms_per_step"],4), round(d["value"]), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],4), "gal", round(s["cosine"],4), "match", round(s["match"],4), "host", round(d["host_enqueue_ms_per_step"],4))
PY
  i=$((i+1))
done
