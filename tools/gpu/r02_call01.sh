cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_c1_smi.txt
( for c in v4:self_1tile v4:self_small v4:self_1024 v4:self_ragged v4:odd_tiles v4:self_inject v4:d40 v4:self_4096 x2:self_1024 x2:self_ragged x2:d40; do echo "== $c"; timeout 120 python tests/gpu_diag.py $c 2>&1 | tail -8; done ) > gpurun_out/r02_v4_x2_diag.log 2>&1
RTTI_KBENCH_ONLY=self timeout 200 python tests/kernel_bench.py > gpurun_out/r02_kb_v3.jsonl 2>&1
RTTI_ATTN_V4=1 RTTI_KBENCH_ONLY=self timeout 200 python tests/kernel_bench.py > gpurun_out/r02_kb_v4.jsonl 2>&1
RTTI_ATTN_X2=1 RTTI_KBENCH_ONLY=self timeout 200 python tests/kernel_bench.py > gpurun_out/r02_kb_x2.jsonl 2>&1
tail -30 gpurun_out/r02_v4_x2_diag.log; cat gpurun_out/r02_kb_v3.jsonl gpurun_out/r02_kb_v4.jsonl gpurun_out/r02_kb_x2.jsonl
