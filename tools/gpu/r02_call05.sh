cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 240 python tests/gpu_diag.py --many self_1tile self_small self_inject self_ragged self_1024 self_lse_pm d40 d32 self_group2_lse self_group3 self_group5 self_group7_split self_group_scattered self_group5_d40 self_rescale self_4096 self_xl32 self_group5_xl32 > gpurun_out/r02_c5_diag.log 2>&1
echo "diag rc=$?" >> gpurun_out/r02_c5_diag.log
RTTI_KBENCH_ONLY=self timeout 200 python tests/kernel_bench.py > gpurun_out/r02_c5_kb.jsonl 2>&1
grep -E "FAIL|MANY|rc=|Error|error" gpurun_out/r02_c5_diag.log | cut -c1-200 | tail -30
cat gpurun_out/r02_c5_kb.jsonl
timeout 1500 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -s -k "full_size_sd15_step or full_size_sdxl_step or processor or color_guidance or segment_labels" > gpurun_out/r02_c5_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c5_pytest.log
grep -E "full-size|passed|failed|Error|error|assert|rc=" gpurun_out/r02_c5_pytest.log | cut -c1-260 | tail -60
