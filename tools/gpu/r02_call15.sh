cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_c15_smi.txt
timeout 900 python -u -m pytest tests -q -m gpu > gpurun_out/r02_c15_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_c15_pytest.log
timeout 900 python bench.py > gpurun_out/r02_c15_bench.json 2> gpurun_out/r02_c15_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r02_c15_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r02_c15_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','clocks','breakdown_ms','gpu_launches','consistency','sampling_loop','cpu_baseline','gpu_eager_baseline')}); print(d['e2e']); print({k:d['roofline'][k] for k in ('achieved','frac','ms_per_step_in_kernel','launches_timed')}); print({k:d['roofline_cross_attention'][k] for k in ('achieved','frac','ms_per_step_in_kernel')})"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" --csv --log-file gpurun_out/r02_c15_launches_n1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c15_ncu_bench.log 2>&1; echo "ncu launch list rc=$?"; wc -l gpurun_out/r02_c15_launches_n1.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"attn_self_kernel|attn_cross_kernel|ff_geglu_kernel|add_bias_layernorm" -o gpurun_out/r02_hot_kernels -f python tools/gpu/ncu_cases.py > gpurun_out/r02_c15_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/r02_c15_ncu_full.log; ls -la gpurun_out/*.ncu-rep
timeout 200 python tests/unet_graph_diag.py 1 > gpurun_out/r02_c15_unet_b1.log 2>&1; tail -3 gpurun_out/r02_c15_unet_b1.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1500 --csv --log-file gpurun_out/r02_c15_launches_unet_b1.csv python tests/unet_graph_diag.py 1 > /dev/null 2>&1; wc -l gpurun_out/r02_c15_launches_unet_b1.csv
timeout 300 python bench.py --config 2 --no-cpu-baseline > gpurun_out/r02_c15_bench_cfg2.json 2> gpurun_out/r02_c15_bench_cfg2.err; echo "cfg2 rc=$?"; cut -c1-1500 gpurun_out/r02_c15_bench_cfg2.json
timeout 300 python bench.py --config 1 --no-cpu-baseline > gpurun_out/r02_c15_bench_cfg1.json 2> gpurun_out/r02_c15_bench_cfg1.err; echo "cfg1 rc=$?"; cut -c1-900 gpurun_out/r02_c15_bench_cfg1.json
