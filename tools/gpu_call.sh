cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r2_c6_pytest.log 2>&1
tail -4 gpurun_out/r2_c6_pytest.log
( time timeout 400 python bench.py ) > gpurun_out/r2_c6_bench.json 2> gpurun_out/r2_c6_bench.err
( time timeout 900 python bench.py --workload sweep ) > gpurun_out/r2_c6_sweep.json 2> gpurun_out/r2_c6_sweep.err
tail -22 gpurun_out/r2_c6_sweep.err
BARGS="--exact-steps --steps 12 --warmup 6 --no-extras --no-cpu-baseline --no-parity"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py $BARGS > gpurun_out/r2_c6_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tn|attention_tcgen05|layernorm_kernel" -s 200 -c 40 -o gpurun_out/r2_prof -f python bench.py $BARGS > gpurun_out/r2_c6_ncu2.log 2>&1
ls -la gpurun_out | tail -12
