cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in 0 1 3 0 1 3 0 1 3; do
  FSAR_LN_VARIANT=$v timeout 300 python bench.py --no-extras --no-cpu-baseline --no-parity > gpurun_out/r2_c13_bench_ln$v.json 2> gpurun_out/r2_c13.err
  python -c "
import json
d=json.load(open('gpurun_out/r2_c13_bench_ln$v.json'))
k=d['kernels']
print('ln_variant=$v value %.1f clk %s ln %.4f ms/ep' % (d['value'], d['clocks']['sm_mhz'], k['layernorm']['ms_per_episode']))
"
done
