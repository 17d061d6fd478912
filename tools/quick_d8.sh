#!/bin/bash
# GPU tests of the accumulation + a short bench + the ncu launch list -> gpurun_out/$1*
out=${1:-q}
python -m pytest tests/test_gpu_accum.py tests/test_gpu_stripes.py -m gpu -x -q 2>&1 | tail -3
python bench.py --no-shia --no-cpu-baseline --e2e-steps 0 > gpurun_out/${out}_bench.json 2> gpurun_out/${out}_bench.err; tail -3 gpurun_out/${out}_bench.err
python -c "
import json; l=json.load(open('gpurun_out/${out}_bench.json')); print('ms', l['ms_per_step'], 'frac', l['roofline']['frac']); print({k:round(v['ms_per_step'],3) for k,v in l['secondary'].items()})"
bash tools/prof_step.sh ${out}_launches 2>&1 | tail -12 | cut -c1-200
