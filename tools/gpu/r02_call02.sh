cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 240 python tests/gpu_diag.py --many self_1tile self_small self_inject self_ragged self_1024 self_lse_pm d40 d32 self_group2_lse self_group3 self_group5 self_group7_split self_group_scattered self_group5_d40 self_rescale self_4096 self_xl32 self_group5_xl32 > gpurun_out/r02_c2_diag.log 2>&1
echo "diag rc=$?" >> gpurun_out/r02_c2_diag.log
timeout 120 python tests/gpu_diag.py g1:self_group5 >> gpurun_out/r02_c2_diag.log 2>&1
RTTI_KBENCH_ONLY=self timeout 200 python tests/kernel_bench.py > gpurun_out/r02_c2_kb.jsonl 2>&1
RTTI_ATTN_MAX_GROUP=1 RTTI_KBENCH_ONLY=self timeout 200 python tests/kernel_bench.py > gpurun_out/r02_c2_kb_g1.jsonl 2>&1
RTTI_ATTN_MAX_GROUP=3 RTTI_KBENCH_ONLY=self timeout 200 python tests/kernel_bench.py > gpurun_out/r02_c2_kb_g3.jsonl 2>&1
grep -E "FAIL|PASS|MANY|rc=|Error|error" gpurun_out/r02_c2_diag.log | cut -c1-200 | tail -60
cat gpurun_out/r02_c2_kb.jsonl gpurun_out/r02_c2_kb_g1.jsonl gpurun_out/r02_c2_kb_g3.jsonl
