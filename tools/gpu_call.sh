cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_episode.py -m gpu -q ) > gpurun_out/r2_c11_pytest.log 2>&1
tail -6 gpurun_out/r2_c11_pytest.log
for d in 1 0 1 0; do
  FSAR_NO_LN_STREAM=$d timeout 300 python bench.py --no-extras --no-cpu-baseline --no-parity > gpurun_out/r2_c11_bench_nostream$d.json 2> gpurun_out/r2_c11.err
  python -c "
import json
d=json.load(open('gpurun_out/r2_c11_bench_nostream$d.json'))
k=d['kernels']
print('nostream=$d value %.1f clk %s ln %.4f ms/ep (%.0f GB/s, %.1f launches)' % (d['value'], d['clocks']['sm_mhz'], k['layernorm']['ms_per_episode'], k['layernorm']['gbs'], k['layernorm']['launches_per_episode']))
"
done
