cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_episode.py tests/test_preprocess.py -m gpu -q -x ) > gpurun_out/r2_c20_pytest.log 2>&1
tail -3 gpurun_out/r2_c20_pytest.log
for d in 1 0 1 0 1 0; do
  if [ $d = 1 ]; then export FSAR_NO_PRE=1; else unset FSAR_NO_PRE; fi
  timeout 300 python bench.py --no-extras --no-cpu-baseline --no-parity > gpurun_out/r2_c20_bench_nopre$d.json 2> gpurun_out/r2_c20.err
  python -c "
import json
d=json.load(open('gpurun_out/r2_c20_bench_nopre$d.json'))
print('nopre=$d value %.1f clk %s' % (d['value'], d['clocks']['sm_mhz']))
"
done
