cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 150 python -u -m pytest tests/test_parity_gpu.py -q -m gpu -x -k "vae_explicit or gn32 or color_guidance or xl_loops" > gpurun_out/r02_c20_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_c20_pytest.log
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r02_c20_bench.json 2> gpurun_out/r02_c20_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r02_c20_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r02_c20_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','clocks','breakdown_ms','consistency')}); print(d['e2e'])"
