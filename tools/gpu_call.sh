cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | wc -l
( time timeout 600 python -m pytest tests/test_reference_runner.py -m gpu -q -x -k "8" ) > gpurun_out/r2_c9_pytest.log 2>&1
tail -4 gpurun_out/r2_c9_pytest.log
cat gpurun_out/r2_reference_runner_8gpu.json
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
( time timeout 600 $TR --master-port 29521 bench.py --gpus 8 --workload l14_t16 --no-cpu-baseline ) > gpurun_out/r2_c9_l14_8gpu.json 2> gpurun_out/r2_c9_l14_8gpu.err
tail -4 gpurun_out/r2_c9_l14_8gpu.err
( time timeout 900 $TR --master-port 29522 bench.py --gpus 8 --workload sweep --no-parity --sweep-seconds 0.4 ) > gpurun_out/r2_c9_sweep_8gpu.json 2> gpurun_out/r2_c9_sweep_8gpu.err
grep "sweep\|real" gpurun_out/r2_c9_sweep_8gpu.err | tail -20
python -c "
import json
d=json.load(open('gpurun_out/r2_c9_l14_8gpu.json'))
print({k:d[k] for k in ('value','n_gpus','steps','counters')}); print(d['e2e']['value'], d['e2e_u8']['value'], d['module_path']['value'], d['parity'])
"
