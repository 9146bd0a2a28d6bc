set -x
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/launches3.csv python bench.py --steps 1 --warmup 3 --no-e2e > gpurun_out/launches3_run.log 2>&1; tail -1 gpurun_out/launches3_run.log | cut -c1-160
timeout 300 ncu --set full --clock-control none --import-source on -k regex:psb_update -c 1 -s 3 -o gpurun_out/prof_update2 python bench.py --steps 1 --warmup 3 --no-e2e > gpurun_out/prof_update2.log 2>&1; tail -1 gpurun_out/prof_update2.log | cut -c1-120
timeout 300 ncu --set full --clock-control none --import-source on -k regex:psb_bcast_gemm2 -s 2 -c 1 -o gpurun_out/prof_gemm2 python bench/gemm_one.py > gpurun_out/prof_gemm2.log 2>&1; tail -1 gpurun_out/prof_gemm2.log | cut -c1-120
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"psb_bn_bwd_apply|psb_bn_stats|psb_encode" -s 40 -c 3 -o gpurun_out/prof_bn python bench.py --steps 1 --warmup 3 --no-e2e > gpurun_out/prof_bn.log 2>&1; tail -1 gpurun_out/prof_bn.log | cut -c1-120
timeout 200 python bench/gemm_bench.py > gpurun_out/gemm_bench_final.jsonl 2>&1; cut -c1-220 gpurun_out/gemm_bench_final.jsonl
ls -la gpurun_out | tail -8
