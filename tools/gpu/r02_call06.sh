cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nproc; free -g | head -2; cat /sys/fs/cgroup/cpu.max 2>/dev/null
timeout 300 python tests/gpu_diag.py --many elementwise ff_geglu self_group5_xl32 self_1024 self_group5_d40 d80 d160 cross_basic cross_fs_cap > gpurun_out/r02_c6_diag.log 2>&1
echo "diag rc=$?" >> gpurun_out/r02_c6_diag.log
grep -E "FAIL|MANY|rc=|Error|error" gpurun_out/r02_c6_diag.log | cut -c1-200 | tail -30
RTTI_KBENCH_ONLY=self timeout 120 python tests/kernel_bench.py > gpurun_out/r02_c6_kb.jsonl 2>&1; echo "kb rc=$?"
RTTI_KBENCH_ONLY=geglu timeout 120 python tests/kernel_bench.py > gpurun_out/r02_c6_kb_geglu.jsonl 2>&1; echo "kb geglu rc=$?"
cat gpurun_out/r02_c6_kb.jsonl gpurun_out/r02_c6_kb_geglu.jsonl
timeout 420 python -u -m pytest tests/test_parity_gpu.py -x -q -m gpu -s -k "full_size_sd15_step or processor or color_guidance or segment_labels" > gpurun_out/r02_c6_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c6_pytest.log
grep -E "full-size|passed|failed|Error|error|assert|rc=|differ" gpurun_out/r02_c6_pytest.log | cut -c1-300 | tail -40
