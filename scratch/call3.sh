#!/bin/bash
# round 2, GPU call 3 (one B200): whole GPU suite with the device-resident async server + snapshot kernels, launch list,
# ncu --set full captures of the kernels that changed, GEMM bench with the promoted epilogue
mkdir -p gpurun_out
export PSB200_NO_AUTOBUILD=1
run() { local name=$1 secs=$2; shift 2; echo "=== $name" | tee -a gpurun_out/call3.log; timeout "$secs" "$@" > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(tail -n 1 gpurun_out/$name.log | cut -c1-200))" | tee -a gpurun_out/call3.log; }
cap() { local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$rx" -s "$skip" -c "$cnt" -o "gpurun_out/$name" -f "$@" > "gpurun_out/$name.log" 2>&1
  echo "cap $name: exit $? $(tail -n 1 gpurun_out/$name.log | cut -c1-120)" | tee -a gpurun_out/call3.log; }
: > gpurun_out/call3.log
run c3_gpu_tests 1200 python -m pytest tests -m gpu -q
run c3_gemm_bench 300 python bench/gemm_bench.py
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/c3_launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-comparators > gpurun_out/c3_launches_run.log 2>&1
echo "launch list: exit $?" | tee -a gpurun_out/call3.log
cap c3_prof_update  psb_update_kernel 12 5 python bench.py --steps 1 --warmup 3 --no-e2e --no-comparators
cap c3_prof_stem    "psb_stem_fwd|psb_stem_wgrad" 4 2 python bench.py --steps 1 --warmup 3 --no-e2e --no-comparators
cap c3_prof_gemm2   psb_bcast_gemm2 2 1 python bench/gemm_one.py
cap c3_prof_bnbwd   "psb_bn_bwd_reduce|psb_maxpool_bwd" 30 3 python bench.py --steps 1 --warmup 3 --no-e2e --no-comparators
cat gpurun_out/call3.log
ls -la gpurun_out | tail -12
