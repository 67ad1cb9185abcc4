cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tests/gpu_diag.py --many self_1tile self_small self_inject self_ragged self_1024 self_lse_pm d40 d32 self_group2_lse self_group3 self_group5 self_group7_split self_group_scattered self_group5_d40 self_rescale self_4096 self_xl32 self_group5_xl32 cross_fs sanitizer_small > gpurun_out/r02_c9_diag.log 2>&1
echo "diag rc=$?" >> gpurun_out/r02_c9_diag.log
grep -E "FAIL|MANY|rc=|Error|error" gpurun_out/r02_c9_diag.log | cut -c1-200 | tail -30
for P in 0 4 2; do RTTI_ATTN_POLY=$P RTTI_KBENCH_ONLY=self timeout 120 python tests/kernel_bench.py > gpurun_out/r02_c9_kb_poly$P.jsonl 2>&1; echo "poly $P rc=$?"; cat gpurun_out/r02_c9_kb_poly$P.jsonl; done
RTTI_ATTN_POLY=4 timeout 200 python tests/gpu_diag.py --many self_1024 self_ragged self_lse_pm d40 self_rescale > gpurun_out/r02_c9_diag_poly4.log 2>&1; echo "poly4 diag rc=$?"; grep -E "FAIL|MANY" gpurun_out/r02_c9_diag_poly4.log | cut -c1-200
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tests/gpu_diag.py sanitizer_small > gpurun_out/r02_c9_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/r02_c9_memcheck.log | cut -c1-200
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tests/gpu_diag.py sanitizer_small > gpurun_out/r02_c9_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/r02_c9_racecheck.log | cut -c1-200
