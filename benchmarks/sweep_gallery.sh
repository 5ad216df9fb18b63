# A/B sweep of the gallery kernel shape: warp triples per CTA, ring stages per triple, stream chunks (run on the GPU box)
for cfg in "5 4 2" "5 4 4" "4 7 2" "4 7 4" "3 9 2" "3 9 4" "6 3 4" "5 5 3" "6 4 3"; do
  set -- $cfg
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-configs --cosine-ctas $1 --gallery-stages $2 --chunks $3 > gpurun_out/r2o_$1_$2_$3.json 2> gpurun_out/r2o.err
done
