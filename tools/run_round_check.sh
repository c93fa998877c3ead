#!/bin/bash
# one gpurun call: GPU test-suite, bench line, ncu launch list of the bench command, ncu --set full of the step's kernels
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2s2_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r2s2_tests.log
python bench.py > gpurun_out/r2s2_bench.json 2> gpurun_out/r2s2_bench.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2s2_launches_B4096.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r2s2_bench_under_ncu.json 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:"raster_fwd|mano_|tf32x3" -s 14 -c 7 -o gpurun_out/r2s2_full -f \
    python tools/prof_fused.py 4096 4 > gpurun_out/r2s2_prof.log 2>&1
tail -2 gpurun_out/r2s2_prof.log
