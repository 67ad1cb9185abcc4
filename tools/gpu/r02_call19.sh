cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 2 --master-port 29571 bench.py --gpus 2 --steps 20 --warmup 5 --check > gpurun_out/r02_c19_bench_n2.json 2> gpurun_out/r02_c19_bench_n2.err
echo "bench n2 rc=$?"; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r02_c19_bench_n2.err | tail -3 | cut -c1-300; python -c "
import json
for l in open('gpurun_out/r02_c19_bench_n2.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print({k:d.get(k) for k in ('value','ms_per_step','breakdown_ms','consistency','sampling_loop','ranks_bit_identical','single_gpu_check','fallback')}); print(d['e2e'])"
