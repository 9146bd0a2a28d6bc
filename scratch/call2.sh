#!/bin/bash
# round 2, GPU call 2 (one B200): the pipelined engine + promoted kernels through the whole GPU test-suite, then the bench
mkdir -p gpurun_out
export PSB200_NO_AUTOBUILD=1
run() { local name=$1 secs=$2; shift 2; echo "=== $name" | tee -a gpurun_out/call2.log; timeout "$secs" "$@" > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(tail -n 1 gpurun_out/$name.log | cut -c1-200))" | tee -a gpurun_out/call2.log; }
: > gpurun_out/call2.log
run c2_smoke 240 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
run c2_gpu_tests 900 python -m pytest tests -m gpu -x -q
run c2_bench 400 python bench.py --steps 20 --warmup 5
run c2_bench_prof 300 python bench.py --steps 20 --warmup 5 --no-comparators --profile
PSB200_STEM=im2col run c2_bench_im2col 300 python bench.py --steps 20 --warmup 5 --no-comparators --bcast-gemm off
cat gpurun_out/call2.log
