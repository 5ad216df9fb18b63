# A/B sweep of the gallery kernel shape (run on the GPU box): warp triples per SM, ring stages per triple, stream
# chunks, waves (0 = one persistent CTA per SM, W = one triple per CTA in W waves), extra bench.py flags.
# usage: bash benchmarks/sweep_gallery.sh TAG "T S P W [flags]" ...
tag=$1; shift
for cfg in "$@"; do
  set -- $cfg
  t=$1; s=$2; p=$3; w=$4; shift 4
  name=$(echo "${t}_${s}_${p}_${w}$*" | tr -d ' -')
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-configs --cosine-ctas $t --gallery-stages $s --chunks $p \
      --gallery-waves $w $* > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}.err
done
