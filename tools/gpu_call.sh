cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -14 > gpurun_out/r2_topo_8gpu.txt
lscpu | grep -i "numa\|socket\|model name\|^CPU(s)" >> gpurun_out/r2_topo_8gpu.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
( time timeout 600 $TR --master-port 29531 bench.py --gpus 8 --no-cpu-baseline ) > gpurun_out/r2_final_bench_8gpu.json 2> gpurun_out/r2_final_bench_8gpu.err
tail -3 gpurun_out/r2_final_bench_8gpu.err
python -c "
import json
d=json.load(open('gpurun_out/r2_final_bench_8gpu.json'))
print({k:d[k] for k in ('value','n_gpus','steps','counters','numa')}); print('e2e', d['e2e']['value'], 'e2e_u8', d['e2e_u8']['value'], 'module', d['module_path']['value'])
"
cat gpurun_out/r2_topo_8gpu.txt | tail -8
