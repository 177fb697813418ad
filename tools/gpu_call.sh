cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_episode.py tests/test_text_tower.py -m gpu -q ) > gpurun_out/r2_c10_pytest.log 2>&1
tail -6 gpurun_out/r2_c10_pytest.log
for d in 1 0 1 0; do
  FSAR_NO_ATT_SPLIT=$d timeout 300 python bench.py --no-extras --no-cpu-baseline --no-parity > gpurun_out/r2_c10_bench_nosplit$d.json 2> gpurun_out/r2_c10.err
  python -c "
import json
d=json.load(open('gpurun_out/r2_c10_bench_nosplit$d.json'))
k=d['kernels']
print('nosplit=$d value %.1f clk %s att %.4f ms/ep (%s TF) ln %.4f (%.1f launches)' % (d['value'], d['clocks']['sm_mhz'], k['attention']['ms_per_episode'], round(k['attention']['tflops']), k['layernorm']['ms_per_episode'], k['layernorm']['launches_per_episode']))
"
done
