#!/bin/bash
# ncu captures for the kernels that change in round 2 (one GPU; run AFTER scratch/round2_first_call.sh is green):
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash scratch/prof_round2.sh'
# then, here:  python bench/summarize_ncu.py gpurun_out/<name>.ncu-rep profiles/<name>.ncu.txt
mkdir -p gpurun_out
cap() {   # cap <name> <kernel regex> <skip> <count> <command...>
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$rx" -s "$skip" -c "$cnt" -o "gpurun_out/$name" -f "$@" \
      > "gpurun_out/$name.log" 2>&1
  echo "$name: exit $? $(tail -n 1 gpurun_out/$name.log | cut -c1-120)"
}
cap stem_fwd      psb_stem_fwd      1 1 python bench/stem_fused_check.py --only timing
cap stem_wgrad    psb_stem_wgrad    1 1 python bench/stem_fused_check.py --only wgrad_implicit
cap bnpool        "psb_bnrelu_pool|psb_bnpool_bwd" 6 3 python bench/bnpool_check.py
cap gemm_epi      psb_bcast_gemm2x  0 4 python bench/gemm_variants.py
PSB200_STEM=fused PSB200_STEM_WGRAD=implicit PSB200_BNPOOL=fused timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none \
    -c 1600 --csv --log-file gpurun_out/launches_fused.csv python bench.py --steps 1 --warmup 3 --no-e2e > gpurun_out/launches_fused_run.log 2>&1
echo "launch list: exit $?"
ls -la gpurun_out | tail -12
