# ncu captures of the C3 tick (single stream chunk): launch list of every kernel + full-set capture of the first timed tick
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --chunks 1"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 760 --csv --log-file gpurun_out/launches_r1h.csv $CMD > gpurun_out/ncu_l.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_cosine_h|k_match|k_apply|k_gate" -s 452 -c 4 -o gpurun_out/prof_r1h $CMD > gpurun_out/ncu_f.log 2>&1
tail -n 2 gpurun_out/ncu_f.log | cut -c1-300
