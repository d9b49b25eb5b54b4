cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider > gpurun_out/r2_pytest_final.log 2>&1; tail -4 gpurun_out/r2_pytest_final.log
python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -c 600 gpurun_out/r2_bench_final.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
