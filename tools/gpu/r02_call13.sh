cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES=0,1
( RTTI_ATTN_POLY=0 CUDA_VISIBLE_DEVICES=0 timeout 200 python tests/debug_sd_rowsums.py > gpurun_out/r02_c13_sd_poly0.log 2>&1 ) &
( CUDA_VISIBLE_DEVICES=1 timeout 200 python tests/debug_sd_rowsums.py > gpurun_out/r02_c13_sd_poly4.log 2>&1 ) &
wait
grep -h "POLY" gpurun_out/r02_c13_sd_poly0.log gpurun_out/r02_c13_sd_poly4.log | cut -c1-260; tail -3 gpurun_out/r02_c13_sd_poly4.log | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multigpu_check.py > gpurun_out/r02_c13_multigpu_check_n2.log 2>&1
echo "multigpu_check rc=$?"; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r02_c13_multigpu_check_n2.log | tail -40 | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 6 --warmup 3 --check > gpurun_out/r02_c13_bench_n2.json 2> gpurun_out/r02_c13_bench_n2.err
echo "bench n2 rc=$?"; tail -5 gpurun_out/r02_c13_bench_n2.err | cut -c1-300; python -c "
import json
for l in open('gpurun_out/r02_c13_bench_n2.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print({k:d.get(k) for k in ('value','ms_per_step','parallelism','breakdown_ms','gpu_launches','consistency','sampling_loop','ranks_bit_identical','single_gpu_check','fallback')}); print(d['e2e'])"
