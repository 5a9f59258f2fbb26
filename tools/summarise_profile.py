"""Developer tool: turn gpurun_out/<tag>_launches.csv and <tag>_prof.ncu-rep into the committed summaries under profiles/."""
import collections
import csv
import json
import os
import subprocess
import sys

tag = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, dst = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")
rows = [r for r in csv.reader(open(os.path.join(src, tag + "_launches.csv"))) if len(r) > 5]
hdr = rows[0]
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
out = [f"# {tag}: ncu --metrics gpu__time_duration.sum --clock-control none, tools/profile_frame.py 3 (C3 stand-in 3840x2160 depth 3, SAH BVH, one batch per frame)",
       "# per-launch device time in us (cold-cache, serialised: compare SHARES)", "id,kernel,grid,block,us"]
agg, tot = collections.OrderedDict(), 0.0
FRAME_KERNELS = ("k_level_reset", "k_extend", "k_shade", "k_shadow", "k_sphere_finalize", "k_plane_finalize", "k_resolve", "k_pack_rgb", "k_post")
for r in rows[1:]:
    n = r[ki].split("(")[0].replace("void ", "").replace("rtb::", "")
    if not n.startswith(FRAME_KERNELS):   # scene upload / BVH build / tie-key kernels run once, before the frames
        continue
    v = float(r[vi].replace(",", "")) / 1e3
    out.append(f"{r[0]},{n},{r[gi].strip()},{r[bi].strip()},{v:.1f}")
    agg.setdefault(n, [0.0, 0])
    agg[n][0] += v
    agg[n][1] += 1
    tot += v
out.append("# shares: " + ", ".join(f"{n} {a[0] / tot * 100:.1f}% ({a[1]} launches)" for n, a in agg.items()))
open(os.path.join(dst, tag + "_launches.csv"), "w").write("\n".join(out) + "\n")
print(out[-1])
raw = subprocess.run(["ncu", "-i", os.path.join(src, tag + "_prof.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(os.path.join(dst, tag + "_extend_shade_shadow_raw.csv"), "w").write(raw)
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"]
vals = {}
for w in want:
    if w in hdr:
        i = hdr.index(w)
        vals[w] = [r[i] for r in rows[2:]]
        print(f"{w:70s} {units[i]:10s}", [r[i][:16] for r in rows[2:]])
mul = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}
ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
per = {}
for n, r, w in zip(vals["Kernel Name"], vals["dram__bytes_read.sum"], vals["dram__bytes_write.sum"]):
    k = "extend" if "k_extend" in n else "shadow_point" if "k_shadow_point" in n else "shade"
    per.setdefault(k, []).append(float(r) * mul[ur] + float(w) * mul[uw])
tr = {k: sum(v) / len(v) for k, v in per.items()}
tr["_note"] = f"dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the captured launches (all bounce levels of frame 2), ncu --set full, profiles/{tag}_extend_shade_shadow_raw.csv"
json.dump(tr, open(os.path.join(dst, "traffic.json"), "w"), indent=1)
print(tr)
