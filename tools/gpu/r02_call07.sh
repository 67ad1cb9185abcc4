cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1300 python -u -m pytest tests -x -q -m gpu -s > gpurun_out/r02_c7_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c7_pytest.log
grep -E "full-size|passed|failed|Error|error|assert|rc=|agree|err " gpurun_out/r02_c7_pytest.log | cut -c1-300 | tail -40
timeout 600 python bench.py --steps 6 --warmup 3 > gpurun_out/r02_c7_bench.json 2> gpurun_out/r02_c7_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r02_c7_bench.err; cat gpurun_out/r02_c7_bench.json
