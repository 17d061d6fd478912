#!/bin/bash
# ncu launch list (per-launch duration, warp instructions, DRAM bytes) of a few accumulation steps -> gpurun_out/$1.csv
set -e
out=${1:-launches}
C="python bench.py --steps 2 --warmup 3 --no-shia --no-cpu-baseline --e2e-steps 0 --no-secondary"
$C > gpurun_out/${out}_bench.json 2> gpurun_out/${out}_bench.err
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -c 120 --csv --log-file gpurun_out/${out}.csv $C > gpurun_out/${out}_ncu.log 2>&1
python tools/launches.py gpurun_out/${out}.csv
cat gpurun_out/${out}_bench.json
