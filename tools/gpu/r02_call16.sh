cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
CUDA_VISIBLE_DEVICES=0 timeout 120 rich-text-to-image_b200/build/tmem_bw > gpurun_out/r02_c16_tmem_bw.txt 2>&1; echo "tmem_bw rc=$?"; cat gpurun_out/r02_c16_tmem_bw.txt
timeout 400 $TR --nproc-per-node 2 --master-port 29551 tests/multigpu_check.py > gpurun_out/r02_c16_multigpu_check_n2.log 2>&1
echo "multigpu_check rc=$?"; grep "world=\|guidance\|MULTIGPU\|Error\|error" gpurun_out/r02_c16_multigpu_check_n2.log | cut -c1-260 | tail -24
timeout 400 $TR --nproc-per-node 2 --master-port 29552 bench.py --gpus 2 --steps 10 --warmup 3 --check > gpurun_out/r02_c16_bench_n2.json 2> gpurun_out/r02_c16_bench_n2.err
echo "bench n2 rc=$?"; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r02_c16_bench_n2.err | tail -3 | cut -c1-300; python -c "
import json
for l in open('gpurun_out/r02_c16_bench_n2.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print({k:d.get(k) for k in ('value','ms_per_step','breakdown_ms','consistency','ranks_bit_identical','single_gpu_check','fallback')}); print(d['e2e'])"
