cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_c11_smi.txt
timeout 900 python -u -m pytest tests -q -m gpu > gpurun_out/r02_c11_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_c11_pytest.log
timeout 300 python tests/kernel_bench.py > gpurun_out/r02_c11_kb_all.jsonl 2>&1; echo "kb all rc=$?"; cat gpurun_out/r02_c11_kb_all.jsonl | cut -c1-400
timeout 900 python bench.py > gpurun_out/r02_c11_bench.json 2> gpurun_out/r02_c11_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r02_c11_bench.err; cat gpurun_out/r02_c11_bench.json
