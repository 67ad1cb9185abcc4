cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tests/gpu_diag.py --many cross_small cross_basic cross_fs cross_cap cross_fs_cap d8 d32 d40_cross d160_cross cross_xl64 self_1024 self_group5 sanitizer_small > gpurun_out/r02_c12_diag.log 2>&1
echo "diag rc=$?"; grep -E "FAIL|MANY|rc=|Error|error" gpurun_out/r02_c12_diag.log | cut -c1-200 | tail -20
timeout 300 python tests/kernel_bench.py > gpurun_out/r02_c12_kb_all.jsonl 2>&1; echo "kb all rc=$?"; cut -c1-330 gpurun_out/r02_c12_kb_all.jsonl
timeout 900 python -u -m pytest tests -q -m gpu -x -k "kernel_case or attention_vs or unet_vs or xl_loops or sd_loops or processor or batched_equals or color_guidance or labels or properties" > gpurun_out/r02_c12_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_c12_pytest.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02_c12_bench.json 2> gpurun_out/r02_c12_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r02_c12_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r02_c12_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','clocks','breakdown_ms','gpu_launches','consistency','sampling_loop')}); print(d['e2e']); print({k:d['roofline'][k] for k in ('achieved','frac','ms_per_step_in_kernel','launches_timed')}); print({k:d['roofline_cross_attention'][k] for k in ('achieved','frac','ms_per_step_in_kernel')})"
