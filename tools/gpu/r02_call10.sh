cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tests/gpu_diag.py --many cross_small cross_basic cross_fs cross_cap cross_fs_cap d8 d40_cross d160_cross cross_xl64 self_1024 self_group5 sanitizer_small > gpurun_out/r02_c10_diag.log 2>&1
echo "diag rc=$?" >> gpurun_out/r02_c10_diag.log
grep -E "FAIL|MANY|rc=|Error|error" gpurun_out/r02_c10_diag.log | cut -c1-200 | tail -30
for P in 4 3; do RTTI_ATTN_POLY=$P RTTI_KBENCH_ONLY=self timeout 120 python tests/kernel_bench.py > gpurun_out/r02_c10_kb_poly$P.jsonl 2>&1; echo "poly $P rc=$?"; grep plain gpurun_out/r02_c10_kb_poly$P.jsonl; done
timeout 200 python tests/kernel_bench.py > gpurun_out/r02_c10_kb_all.jsonl 2>&1; echo "kb all rc=$?"; grep -E "cross|groupnorm|layernorm|geglu|color|blend" gpurun_out/r02_c10_kb_all.jsonl
timeout 600 python -u -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "attention_vs_reference or xl_loops or sd_loops or processor or attention_properties" > gpurun_out/r02_c10_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_c10_pytest.log
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c10_bench.json 2> gpurun_out/r02_c10_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r02_c10_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r02_c10_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','clocks','breakdown_ms')}); print(d['e2e']); print({k:d['roofline'][k] for k in ('achieved','frac','ms_per_step_in_kernel','launches_timed')}); print({k:d['roofline_cross_attention'][k] for k in ('achieved','frac','ms_per_step_in_kernel')})"
