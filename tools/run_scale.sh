#!/bin/bash
# usage: bash tools/run_scale.sh N   (under gpurun --gpus N): strong-scaling bench line + host->device ceiling
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu-baseline --no-other-configs > gpurun_out/r2s2_bench_n$N.json 2> gpurun_out/r2s2_bench_n$N.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2s2_bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/h2d_ceiling.py 128 > gpurun_out/r2s2_h2d_n$N.json 2>/dev/null
cat gpurun_out/r2s2_h2d_n$N.json | tail -1
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2s2_bench_n$N.json') if l.startswith('{')][0])
print('N=$N value %.3e ms %.4f e2e %.3e weak %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], (d.get('weak_scaling') or {}).get('value')))
PY
