"""Extract per-launch DRAM traffic of the attention kernels from an `ncu --set full` capture into
profiles/r02_ncu_traffic.json (read by bench.py for `roofline.traffic`).

    python tools/ncu_traffic.py gpurun_out/<capture>.ncu-rep [shape note]

traffic = dram__bytes_read.sum + dram__bytes_write.sum of ONE launch (B200_PROFILING.md); when the capture holds several
launches of a kernel the largest-grid one is taken (the batch-8 shape of the benchmark step)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep = sys.argv[1]
    note = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    best = {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        key = "attn_self_kernel" if "attn_self_kernel" in name else ("attn_fwd_kernel_cross" if "attn_fwd_kernel<80" in name.replace("(int)", "") else None)
        if key is None:
            continue
        rd = float(r[col["dram__bytes_read.sum"]]) * UNIT[units[col["dram__bytes_read.sum"]]]
        wr = float(r[col["dram__bytes_write.sum"]]) * UNIT[units[col["dram__bytes_write.sum"]]]
        grid = int(float(r[col["launch__grid_size"]]))
        if key not in best or grid > best[key]["grid"]:
            best[key] = {"bytes_per_launch": rd + wr, "grid": grid, "kernel": name[:80],
                         "note": f"dram__bytes_read.sum + dram__bytes_write.sum = {rd / 1e6:.2f} + {wr / 1e6:.2f} MB for one launch "
                                 f"(grid {grid} CTAs) in the ncu --set full capture {os.path.basename(rep)} {note}".strip()}
    p = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    old = json.load(open(p)) if os.path.exists(p) else {}
    old.update(best)
    json.dump(old, open(p, "w"), indent=1)
    print(json.dumps(best, indent=1))


if __name__ == "__main__":
    main()
