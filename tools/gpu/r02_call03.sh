cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 240 python tests/gpu_diag.py --many self_1tile self_small self_inject self_ragged self_1024 self_lse_pm d40 d32 self_group2_lse self_group3 self_group5 self_group7_split self_group_scattered self_group5_d40 self_rescale self_4096 self_xl32 self_group5_xl32 > gpurun_out/r02_c4_diag.log 2>&1
echo "diag rc=$?" >> gpurun_out/r02_c4_diag.log
nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader -lms 100 > gpurun_out/r02_c4_clocks.csv &
SMI=$!
RTTI_KBENCH_ONLY=self timeout 200 python tests/kernel_bench.py > gpurun_out/r02_c4_kb.jsonl 2>&1
RTTI_ATTN_MAX_GROUP=1 RTTI_KBENCH_ONLY=self timeout 200 python tests/kernel_bench.py > gpurun_out/r02_c4_kb_g1.jsonl 2>&1
kill $SMI
grep -E "FAIL|MANY|rc=|Error|error" gpurun_out/r02_c4_diag.log | cut -c1-200 | tail -30
cat gpurun_out/r02_c4_kb.jsonl gpurun_out/r02_c4_kb_g1.jsonl
sort -t, -k1 -n gpurun_out/r02_c4_clocks.csv | awk -F, '{print $1}' | sort -n | uniq -c | sort -rn | head -8
# ncu full capture of the plain and the group kernel at the XL-32 shape (small launch list)
cat > /tmp/ncu_case.py <<'PY'
import torch, sys
sys.path.insert(0, '.')
from rtti_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
B, H, T = 8, 20, 1024
C = H * 64
qkv = torch.randn(B, T, 3 * C, device="cuda", generator=g).half()
q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
o = torch.empty(B, T, C, device="cuda", dtype=torch.float16)
for _ in range(2):
    ops.attention(q, k, v, H, out=o)
    ops.attention(q, k, v, H, out=o, qk_src=[0, 1, 2, 3, 3, 3, 3, 3])
torch.cuda.synchronize()
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_self_kernel -s 3 -c 3 -o gpurun_out/r02_attn_self_tmemP -f python /tmp/ncu_case.py > gpurun_out/r02_c4_ncu.log 2>&1
tail -3 gpurun_out/r02_c4_ncu.log
