# A/B of prebuilt library variants (benchmarks/ab/lib_<name>.so, built beforehand with DD_NVCC_EXTRA): each is copied
# over the package's library on the box's scratch copy and the C3 line is measured.  usage: run_ab_lib.sh TAG name...
mkdir -p gpurun_out
tag=$1; shift
for v in "$@"; do
  cp benchmarks/ab/lib_$v.so deepdish_b200/libdeepdish_b200.so
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-configs > gpurun_out/${tag}_$v.json 2> gpurun_out/${tag}_$v.err
  python - <<PY
import json
for l in open("gpurun_out/${tag}_$v.json"):
    if l.startswith("{"):
        d=json.loads(l); s=d["stage_ms"]; print("[$v]", "tick", round(d["ms_per_step"],4), round(d["value"]), "e2e", round(d["e2e"]["value"]), "gal", round(s["cosine"],4), "match", round(s["match"],4))
PY
done
