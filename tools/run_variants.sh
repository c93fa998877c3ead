#!/bin/bash
# run bench.py against each experimental build of the library (kernel tuning)
for so in build/var_*.so; do
  DSF_B200_LIB=$PWD/$so python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/bench_$(basename $so .so).json 2>/dev/null
  python -c "
import json,sys; d=json.load(open('gpurun_out/bench_$(basename $so .so).json')); print('$(basename $so .so)', round(d['ms_per_step'],4), round(d['roofline']['stage_ms']['raster_fwd_kernel'],4), round(d['roofline']['stage_ms']['raster_bwd_kernel'],4))"
done
