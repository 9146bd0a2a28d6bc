#!/bin/bash
# usage: ncu_rank0.sh <nranks> <out-prefix> -- <script.py> [args…]
# Starts <nranks> ranks of one node (torchrun-compatible env); ONLY rank 0 runs under ncu, with a handful of metrics.
N=$1; OUT=$2; shift 3
PORT=${PORT:-29733}
METRICS=gpu__time_duration.sum,nvlrx__bytes.sum,nvltx__bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active
pids=()
for r in $(seq 1 $((N-1))); do
  RANK=$r LOCAL_RANK=$r WORLD_SIZE=$N LOCAL_WORLD_SIZE=$N MASTER_ADDR=127.0.0.1 MASTER_PORT=$PORT python "$@" > "$OUT.rank$r.log" 2>&1 &
  pids+=($!)
done
RANK=0 LOCAL_RANK=0 WORLD_SIZE=$N LOCAL_WORLD_SIZE=$N MASTER_ADDR=127.0.0.1 MASTER_PORT=$PORT \
  timeout 600 ncu --metrics $METRICS --clock-control none -k regex:psb_update_kernel --csv --log-file "$OUT.csv" python "$@" > "$OUT.rank0.log" 2>&1
rc=$?
for p in "${pids[@]}"; do wait $p || rc=$?; done
echo "ncu_rank0 $OUT: exit $rc $(tail -n 1 $OUT.rank0.log | cut -c1-160)"
