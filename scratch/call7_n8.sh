#!/bin/bash
# round 2, GPU call 7 (8 x B200, ~5 min): NVLS-pipeline correctness at N = 8, headline bench at N = 8 and N = 1 on the same
# box, config-5 sweep at N = 8, rank-0 ncu of the update kernel with NVLink counters
mkdir -p gpurun_out
export PSB200_NO_AUTOBUILD=1
run() { local name=$1 secs=$2; shift 2; echo "=== $name" | tee -a gpurun_out/call7.log; timeout "$secs" "$@" > "gpurun_out/$name.log" 2>&1; echo "    exit $? ($(tail -n 1 gpurun_out/$name.log | cut -c1-200))" | tee -a gpurun_out/call7.log; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
: > gpurun_out/call7.log
run c7_tests 240 python -m pytest "tests/test_gpu_engine.py::test_big_bf16_arena_multi_gpu" "tests/test_gpu_engine.py::test_engine_multi_gpu" -q -x
run c7_bench_n8 150 $TR --nproc-per-node 8 --master-port 29711 bench.py --gpus 8 --steps 30 --warmup 5 --no-comparators --profile
run c7_sweep_n8 240 $TR --nproc-per-node 8 --master-port 29712 bench/bandwidth_sweep.py --max-mb 1024 --peer-copy --out gpurun_out/bw_sweep_n8.json
PORT=29713 timeout 120 bash scratch/ncu_rank0.sh 8 gpurun_out/update_nvls_n8 -- bench/update_probe.py --mb 64 --reduce auto | tee -a gpurun_out/call7.log
PORT=29714 timeout 120 bash scratch/ncu_rank0.sh 8 gpurun_out/update_p2p_n8 -- bench/update_probe.py --mb 64 --reduce p2p | tee -a gpurun_out/call7.log
run c7_bench_n1 120 python bench.py --steps 30 --warmup 5 --no-comparators
run c7_bench_n8_cmp 200 $TR --nproc-per-node 8 --master-port 29715 bench.py --gpus 8 --steps 20 --warmup 5
cat gpurun_out/call7.log
