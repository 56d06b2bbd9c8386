#!/bin/bash
# Evidence of a round, run on the GPU box: bench line, launch list of the bench command, one `ncu --set full`
# capture per hot kernel (raw page + source summary; the .ncu-rep files stay in /tmp).
#   gpurun --timeout 1500 -- 'bash tools/capture_profiles.sh r02'
# Results land in gpurun_out/; copy what should be judged into profiles/.
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --set full --clock-control none --import-source on"
python bench.py > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/launches_${TAG}_bench.csv \
    python bench.py --steps 20 --warmup 3 > /dev/null 2>&1
cap() {  # name, kernel regex, skip, env..., script
  local name=$1 regex=$2 skip=$3; shift 3
  env "$@" > /dev/null 2>&1 || true
}
run_cap() {  # name regex skip script [env assignments...]
  local name=$1 regex=$2 skip=$3 script=$4; shift 4
  env "$@" $NCU -k regex:$regex -s $skip -c 1 -f -o /tmp/prof_$name python $script > $OUT/prof_${name}.log 2>&1
  ncu -i /tmp/prof_$name.ncu-rep --page raw --csv > $OUT/${name}_${TAG}_ncu_raw.csv 2>/dev/null
  python tools/ncu_src_summary.py /tmp/prof_$name.ncu-rep 25 > $OUT/${name}_${TAG}_source_summary.txt 2>&1
}
if [ -z "$CAPTURE_RANSAC_ONLY" ]; then
run_cap k1_eval_warp eval_warp_kernel 2 tools/profile_run.py PROF_MODE=eval
run_cap solve_slots_kernel solve_slots_kernel 2 tools/profile_run.py PROF_MODE=solve
run_cap solve_kernel 'solve_kernel' 2 tools/profile_run.py PROF_MODE=solve PNEC_B200_SOLVE_SLOTS=0
run_cap es_lm_kernel es_lm_kernel 1 tools/profile_frame.py
run_cap scf_kernel 'scf_kernel' 1 tools/profile_frame.py
fi
if [ -z "$CAPTURE_RANSAC_ONLY" ]; then
run_cap ransac_kernel 'ransac_kernel' 0 tools/ransac_once.py PNEC_B200_RANSAC_SPLIT=0
run_cap ransac_hyp_kernel 'ransac_hyp_kernel' 1 tools/ransac_once.py
fi
run_cap ransac_lm_kernel 'ransac_lm_kernel' 0 tools/ransac_once.py OUTLIERS=0
run_cap ransac_post_kernel 'ransac_post_kernel' 0 tools/ransac_once.py OUTLIERS=0
ls -la $OUT | tail -30
