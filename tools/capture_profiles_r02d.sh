#!/bin/bash
# End-of-round-2 evidence (run on the GPU box): bench line, launch list of the bench command, ncu pages of the
# kernels this part of the round changed (slots kernel fed in conditioning order, the score kernel, the rotation
# LM with wide turns), stage / latency timings.
#   gpurun --timeout 1500 -- 'bash tools/capture_profiles_r02d.sh'
TAG=r02d
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --set full --clock-control none --import-source on"
python bench.py > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $OUT/launches_${TAG}_bench.csv \
    python bench.py --steps 20 --warmup 3 > /dev/null 2>&1
run_cap() {  # name regex skip script [env assignments...]
  local name=$1 regex=$2 skip=$3 script=$4; shift 4
  env "$@" $NCU -k regex:$regex -s $skip -c 1 -f -o /tmp/prof_$name python $script > $OUT/prof_${name}.log 2>&1
  ncu -i /tmp/prof_$name.ncu-rep --page raw --csv > $OUT/${name}_${TAG}_ncu_raw.csv 2>/dev/null
  python tools/ncu_src_summary.py /tmp/prof_$name.ncu-rep 25 > $OUT/${name}_${TAG}_source_summary.txt 2>&1
}
run_cap solve_slots_kernel solve_slots_kernel 2 tools/profile_run.py PROF_MODE=solve
run_cap solve_score_kernel solve_score_kernel 2 tools/profile_run.py PROF_MODE=solve
run_cap es_lm_kernel es_lm_kernel 1 tools/profile_frame.py
python tools/solve_order_timing.py > $OUT/solve_order_timing_${TAG}.jsonl 2>&1
python tools/frame_timing.py > $OUT/frame_stages_${TAG}.jsonl 2>&1
python tools/es_lm_tail.py > $OUT/es_lm_tail_${TAG}.json 2>&1
python tools/frame_tail.py > $OUT/frame_tail_${TAG}.json 2>&1
python tools/ransac_stage_timing.py 2>&1 | tail -1 > $OUT/ransac_stage_timing_${TAG}.json
python tools/single_pair_latency.py > $OUT/single_pair_latency_${TAG}.jsonl 2>&1
python tools/single_pair_breakdown.py 2>&1 | head -6 > $OUT/single_pair_breakdown_${TAG}.log
python tools/aux_timing.py > $OUT/aux_timing_${TAG}.jsonl 2>&1
CONFIGS_OUT=$OUT/configs_${TAG}.json python tools/config_sweep.py > $OUT/configs_${TAG}.log 2>&1
ls -la $OUT | tail -30
