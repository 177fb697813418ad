cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for d in 1 0 1 0 1 0; do
  FSAR_NO_MOD_FUSED=$d timeout 300 python bench.py --batch 1 --no-extras --no-cpu-baseline --no-parity --min-seconds 1.0 > gpurun_out/r2_c17_bench_nofused$d.json 2> gpurun_out/r2_c17.err
  python -c "
import json
d=json.load(open('gpurun_out/r2_c17_bench_nofused$d.json'))
k=d['kernels']
print('nofused=$d one-episode-per-call value %.1f clk %s modulator %.4f ms/ep (%.1f launches)' % (d['value'], d['clocks']['sm_mhz'], k['modulator']['ms_per_episode'], k['modulator']['launches_per_episode']))
"
done
