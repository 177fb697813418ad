set -x
nvidia-smi -L
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_c1_pytest.log 2>&1
tail -5 gpurun_out/r2_c1_pytest.log
( time timeout 300 python bench.py ) > gpurun_out/r2_c1_bench.json 2> gpurun_out/r2_c1_bench.err
tail -c 600 gpurun_out/r2_c1_bench.err
( time timeout 300 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r2_c1_ref.json 2> gpurun_out/r2_c1_ref.err
( time timeout 300 python bench.py --workload 5w5s --no-cpu-baseline ) > gpurun_out/r2_c1_5w5s.json 2> gpurun_out/r2_c1_5w5s.err
( time timeout 400 python bench.py --workload l14_t16 --no-cpu-baseline ) > gpurun_out/r2_c1_l14.json 2> gpurun_out/r2_c1_l14.err
ls -la gpurun_out
