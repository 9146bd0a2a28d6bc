#!/bin/bash
# First GPU call of round 2 (one B200, ~12 min of box time):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scratch/round2_first_call.sh'
# Order: cheapest sanity first, every experimental kernel in its own process under its own timeout (a device-side trap
# kills only that process), everything tee'd into gpurun_out/ so a cut-off call still leaves evidence.
mkdir -p gpurun_out
export PSB200_NO_AUTOBUILD=1
run() {   # run <name> <seconds> <command...>
  local name=$1 secs=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/round2_first_call.log
  timeout "$secs" "$@" > "gpurun_out/$name.log" 2>&1
  echo "    exit $? ($(tail -n 1 gpurun_out/$name.log | cut -c1-160))" | tee -a gpurun_out/round2_first_call.log
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | tee gpurun_out/round2_first_call.log
run smoke            240 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
run gpu_tests        600 python -m pytest tests -m gpu -x -q
run gemm_variants    300 python bench/gemm_variants.py
for s in numerics wgrad wgrad_implicit model timing; do
  run "stem_$s"      300 python bench/stem_fused_check.py --only "$s"
done
run bnpool           300 python bench/bnpool_check.py
PSB200_TEST_EXPERIMENTAL=1 run gpu_tests_experimental 600 python -m pytest tests/test_gpu_experimental.py -m gpu -q
run bench_default    300 python bench.py --steps 20 --warmup 5
PSB200_STEM=fused PSB200_BNPOOL=fused run bench_fused 300 python bench.py --steps 20 --warmup 5
PSB200_STEM=fused PSB200_STEM_WGRAD=implicit PSB200_BNPOOL=fused run bench_fused_wgrad 300 python bench.py --steps 20 --warmup 5
cat gpurun_out/round2_first_call.log
