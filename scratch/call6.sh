#!/bin/bash
# round 2, GPU call 6 (one B200): everything since call 4 — whole GPU suite (fail fast), bnpool quad backward, benches
mkdir -p gpurun_out
export PSB200_NO_AUTOBUILD=1
run() { local name=$1 secs=$2; shift 2; echo "=== $name" | tee -a gpurun_out/call6.log; timeout "$secs" "$@" > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(tail -n 1 gpurun_out/$name.log | cut -c1-220))" | tee -a gpurun_out/call6.log; }
: > gpurun_out/call6.log
run c6_gpu_tests 600 python -m pytest tests -m gpu -q -x
run c6_bnpool 200 python bench/bnpool_check.py
run c6_bench 300 python bench.py --steps 30 --warmup 5 --no-comparators
PSB200_BNPOOL=fused run c6_bench_bnpool 300 python bench.py --steps 30 --warmup 5 --no-comparators
cat gpurun_out/call6.log
