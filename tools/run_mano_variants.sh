for so in build/var_skin*.so; do
  DSF_B200_LIB=$PWD/$so python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/bench_$(basename $so .so).json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/bench_$(basename $so .so).json')); s=d['roofline']['stage_ms']; print('$(basename $so .so)', round(d['ms_per_step'],4), round(s['mano_forward(3 kernels)'],4), round(s['mano_backward(3 kernels)'],4))"
done
