cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tests/gpu_diag.py --many self_1tile self_small self_inject self_ragged self_1024 self_lse_pm d40 d32 self_group2_lse self_group3 self_group5 self_group7_split self_group_scattered self_group5_d40 self_rescale self_4096 self_xl32 self_group5_xl32 cross_fs sanitizer_small > gpurun_out/r02_c8_diag.log 2>&1
echo "diag rc=$?" >> gpurun_out/r02_c8_diag.log
grep -E "FAIL|MANY|rc=|Error|error" gpurun_out/r02_c8_diag.log | cut -c1-200 | tail -30
RTTI_KBENCH_ONLY=self timeout 120 python tests/kernel_bench.py > gpurun_out/r02_c8_kb.jsonl 2>&1; echo "kb rc=$?"
cat gpurun_out/r02_c8_kb.jsonl
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c8_bench.json 2> gpurun_out/r02_c8_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r02_c8_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r02_c8_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','clocks','breakdown_ms')}); print(d['e2e']); print({k:d['roofline'][k] for k in ('achieved','frac','ms_per_step_in_kernel','launches_timed')})"
