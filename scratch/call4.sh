#!/bin/bash
mkdir -p gpurun_out
export PSB200_NO_AUTOBUILD=1
run() { local name=$1 secs=$2; shift 2; echo "=== $name" | tee -a gpurun_out/call4.log; timeout "$secs" "$@" > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(tail -n 1 gpurun_out/$name.log | cut -c1-220))" | tee -a gpurun_out/call4.log; }
: > gpurun_out/call4.log
run c4_tests 900 python -m pytest tests/test_gpu_fused_ops.py tests/test_gpu_engine.py tests/test_gpu_bn.py -m gpu -q
run c4_bnpool 300 python bench/bnpool_check.py
run c4_stemcheck 300 python bench/stem_fused_check.py --only wgrad_implicit
run c4_bench 400 python bench.py --steps 30 --warmup 5 --no-comparators
PSB200_BNPOOL=fused run c4_bench_bnpool 400 python bench.py --steps 30 --warmup 5 --no-comparators
PSB200_DIRECT_GRAD=0 run c4_bench_nodirect 400 python bench.py --steps 30 --warmup 5 --no-comparators
cat gpurun_out/call4.log
