cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tests/kernel_bench.py > gpurun_out/r02_c18_kb_all.jsonl 2>&1; echo "kb all rc=$?"; grep -E "groupnorm|layernorm|geglu rows|blend|color" gpurun_out/r02_c18_kb_all.jsonl | cut -c1-260
timeout 900 python -u -m pytest tests -q -m gpu -x > gpurun_out/r02_c18_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_c18_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02_c18_bench.json 2> gpurun_out/r02_c18_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r02_c18_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r02_c18_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','clocks','breakdown_ms','consistency','sampling_loop')}); print(d['e2e']); print({k:d['roofline'][k] for k in ('achieved','frac','ms_per_step_in_kernel')}); print({k:d['roofline_cross_attention'][k] for k in ('achieved','frac','ms_per_step_in_kernel')})"
