# A/B of compile-time variants on the GPU box: rebuild the library with extra nvcc flags, run the default bench line.
# usage: bash benchmarks/ab_build.sh TAG "name:-DFOO=1 -DBAR=2" ... ; the last build stays in place (rebuild after)
tag=$1; shift
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  DD_NVCC_EXTRA="$flags" python -m deepdish_b200.build --force > /dev/null 2> gpurun_out/${tag}_${name}.build.err
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-configs > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
done
python -m deepdish_b200.build --force > /dev/null 2>&1
