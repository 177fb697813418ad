# The round's final single-GPU validation, as run through `gpurun --timeout 3300 -- 'bash tools/gpu_call.sh'`
# (outputs under gpurun_out/, summarised into profiles/ by tools/launch_shares.py, ncu_summary.py, ncu_traffic.py,
# ncu_tensor_pipe.py and sass_histogram.sh). Multi-GPU lines: the same bench.py under
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r2_final_pytest.log 2>&1
tail -4 gpurun_out/r2_final_pytest.log
( time timeout 120 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2_final_smoke.log 2>&1
tail -3 gpurun_out/r2_final_smoke.log
( time timeout 400 python bench.py ) > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
( time timeout 300 python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err
( time timeout 300 python bench.py --workload 5w5s ) > gpurun_out/r2_final_5w5s.json 2> gpurun_out/r2_final_5w5s.err
( time timeout 400 python bench.py --workload l14_t16 --no-cpu-baseline ) > gpurun_out/r2_final_l14.json 2> gpurun_out/r2_final_l14.err
( time timeout 900 python bench.py --workload sweep ) > gpurun_out/r2_final_sweep.json 2> gpurun_out/r2_final_sweep.err
BARGS="--exact-steps --steps 12 --warmup 6 --no-extras --no-cpu-baseline --no-parity"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py $BARGS > gpurun_out/r2_final_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tn|attention_tcgen05|layernorm" -s 200 -c 40 -o gpurun_out/r2_prof -f python bench.py $BARGS > gpurun_out/r2_final_ncu2.log 2>&1
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_memcheck_smoke.txt 2>&1; echo "memcheck rc $?"
