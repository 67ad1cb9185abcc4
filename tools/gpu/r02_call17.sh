cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
CUDA_VISIBLE_DEVICES=0 timeout 120 rich-text-to-image_b200/build/tmem_bw > gpurun_out/r02_c17_tmem_bw.txt 2>&1; echo "tmem_bw rc=$?"; cat gpurun_out/r02_c17_tmem_bw.txt
timeout 400 $TR --nproc-per-node 8 --master-port 29561 bench.py --gpus 8 --steps 20 --warmup 5 --check > gpurun_out/r02_c17_bench_n8.json 2> gpurun_out/r02_c17_bench_n8.err
echo "bench n8 rc=$?"; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r02_c17_bench_n8.err | tail -3 | cut -c1-300; python -c "
import json
for l in open('gpurun_out/r02_c17_bench_n8.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print({k:d.get(k) for k in ('value','ms_per_step','breakdown_ms','consistency','sampling_loop','ranks_bit_identical','single_gpu_check','fallback','clocks')}); print(d['e2e'])"
