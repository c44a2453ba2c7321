#!/bin/bash
# Round-2 evidence run (GPU box): bench line, ncu launch list, ncu --set full captures per kernel group.
# Usage: bash scripts/profile_r2.sh <tag> [parts]   parts: any of b (bench) l (launch list) 2 3 5 (ncu --set full of that
# config's kernels); outputs under gpurun_out/<tag>_* (raw-page CSVs are exported on the box; gpurun brings back <= 64 MiB)
tag=${1:-r2}
parts=${2:-bl235}
out=gpurun_out
mkdir -p $out
NCU="ncu --clock-control none"
full() {   # name, kernel regex, skip, count, bench args...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 900 $NCU --set full --import-source on -k regex:"$rx" --launch-skip $skip -c $cnt -o $out/${tag}_$name -f \
    python bench.py --no-cpu-baseline "$@" > $out/${tag}_ncu_$name.log 2>&1
  echo "$name full rc=$?"
  ncu -i $out/${tag}_$name.ncu-rep --page raw --csv > $out/${tag}_$name.raw.csv 2>/dev/null
}
if [[ $parts == *b* ]]; then
  python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
fi
if [[ $parts == *l* ]]; then
  timeout 600 $NCU --metrics gpu__time_duration.sum -c 6000 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline ${LAUNCH_CONFIGS:+--configs $LAUNCH_CONFIGS} > $out/${tag}_launches.log 2>&1
  echo "launch list rc=$?"
fi
[[ $parts == *2* ]] && full cfg2 'sqdist_tc|cost_finalize_tiled|sinkhorn_fwd_small|sinkhorn_bwd_small|grad_tc|martingale_bwd|build_w_image' 35 7 --steps 2 --warmup 3 --configs ''
[[ $parts == *3* ]] && full cfg3 'axis_col|axis_tile|tie_sums' 8 8 --steps 2 --warmup 3 --configs cfg3_bair
if [[ $parts == *5* ]]; then
  full cfg5gemm 'gemm_f16x3' 0 3 --steps 1 --warmup 3 --configs cfg5_large
  full cfg5sk 'sk_persist_fwd|sk_persist_bwd' 0 2 --steps 1 --warmup 3 --configs cfg5_large
fi
du -sh $out; ls -la $out/${tag}_*
