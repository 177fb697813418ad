cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool racecheck --racecheck-report analysis python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_racecheck_smoke.txt 2>&1; echo "racecheck rc $?"
grep -E "RACECHECK SUMMARY|hazard|smoke:" gpurun_out/r2_sanitizer_racecheck_smoke.txt | sort | uniq -c | head -12
timeout 300 compute-sanitizer --tool synccheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_synccheck_smoke.txt 2>&1; echo "synccheck rc $?"
tail -2 gpurun_out/r2_sanitizer_synccheck_smoke.txt
