cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
show() { python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l)
        print({k: d.get(k) for k in ('n_gpus', 'value', 'ms_per_step', 'breakdown_ms', 'gpu_launches', 'consistency', 'sampling_loop', 'ranks_bit_identical', 'single_gpu_check', 'fallback', 'clocks')})
        print(d['config']['workload'][:60], d['e2e'])
PY
}
timeout 300 $TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 --check > gpurun_out/r02_c14_bench_n8_cfg3.json 2> gpurun_out/r02_c14_bench_n8_cfg3.err
echo "n8 cfg3 rc=$?"; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r02_c14_bench_n8_cfg3.err | tail -4 | cut -c1-300; show gpurun_out/r02_c14_bench_n8_cfg3.json
( CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 300 $TR --nproc-per-node 4 --master-port 29542 bench.py --gpus 4 --steps 10 --warmup 3 --check > gpurun_out/r02_c14_bench_n4_cfg3.json 2> gpurun_out/r02_c14_bench_n4_cfg3.err; echo "n4 cfg3 rc=$?" ) &
( CUDA_VISIBLE_DEVICES=4,5,6,7 timeout 300 $TR --nproc-per-node 4 --master-port 29543 tests/multigpu_check.py > gpurun_out/r02_c14_multigpu_check_n4.log 2>&1; echo "multigpu_check n4 rc=$?" ) &
wait
show gpurun_out/r02_c14_bench_n4_cfg3.json; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r02_c14_bench_n4_cfg3.err | tail -3 | cut -c1-300
grep "world=\|guidance\|MULTIGPU\|Error\|error" gpurun_out/r02_c14_multigpu_check_n4.log | cut -c1-260 | tail -30
timeout 300 $TR --nproc-per-node 8 --master-port 29544 bench.py --gpus 8 --config 4 --steps 10 --warmup 3 --check > gpurun_out/r02_c14_bench_n8_cfg4.json 2> gpurun_out/r02_c14_bench_n8_cfg4.err
echo "n8 cfg4 rc=$?"; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r02_c14_bench_n8_cfg4.err | tail -4 | cut -c1-300; show gpurun_out/r02_c14_bench_n8_cfg4.json
timeout 300 $TR --nproc-per-node 8 --master-port 29545 bench.py --gpus 8 --config 5 --steps 6 --warmup 3 > gpurun_out/r02_c14_bench_n8_cfg5.json 2> gpurun_out/r02_c14_bench_n8_cfg5.err
echo "n8 cfg5 rc=$?"; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r02_c14_bench_n8_cfg5.err | tail -4 | cut -c1-300; show gpurun_out/r02_c14_bench_n8_cfg5.json
