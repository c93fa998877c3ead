#!/bin/bash
for c in 1 2 3 4 8; do
  python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-other-configs --chunks $c > gpurun_out/bench_chunks$c.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/bench_chunks$c.json')); print('chunks $c', round(d['ms_per_step'],4), round(d['value']), d['launches_per_step'], round(d['loss'],6))"
done
