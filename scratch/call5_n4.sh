#!/bin/bash
# round 2, GPU call 5 (4 x B200): multi-GPU correctness (NVLS bf16 default, per-chunk pipeline, big arena, device-resident
# async), bench at N = 2 / 4, bandwidth sweep at N = 2 / 4 (config 5), + the single-GPU re-checks
mkdir -p gpurun_out
export PSB200_NO_AUTOBUILD=1
run() { local name=$1 secs=$2; shift 2; echo "=== $name" | tee -a gpurun_out/call5.log; timeout "$secs" "$@" > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(tail -n 1 gpurun_out/$name.log | cut -c1-220))" | tee -a gpurun_out/call5.log; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
: > gpurun_out/call5.log
nvidia-smi -L | tee -a gpurun_out/call5.log
run c5_multigpu_tests 1200 python -m pytest tests -m "gpu and multigpu" -q -x
run c5_single_rechecks 600 python -m pytest tests/test_gpu_engine.py::test_direct_gradient_placement_matches_encode_path tests/test_gpu_fused_ops.py tests/test_gpu_bn.py -m gpu -q
run c5_bnpool 300 python bench/bnpool_check.py
run c5_bench_n4 600 $TR --nproc-per-node 4 --master-port 29611 bench.py --gpus 4 --steps 30 --warmup 5 --profile
run c5_bench_n2 600 $TR --nproc-per-node 2 --master-port 29612 bench.py --gpus 2 --steps 30 --warmup 5 --no-comparators
run c5_bench_n1 600 python bench.py --steps 30 --warmup 5 --no-comparators
run c5_sweep_n4 900 $TR --nproc-per-node 4 --master-port 29613 bench/bandwidth_sweep.py --max-mb 256 --peer-copy --out gpurun_out/bw_sweep_n4.json
run c5_sweep_n2 900 $TR --nproc-per-node 2 --master-port 29614 bench/bandwidth_sweep.py --max-mb 256 --out gpurun_out/bw_sweep_n2.json
cat gpurun_out/call5.log
