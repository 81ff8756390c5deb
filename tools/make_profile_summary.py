"""Turn the raw evidence of a gpurun (`gpurun_out/`) into the tracked summaries under profiles/.

    python tools/make_profile_summary.py <tag> <round-name>
Reads  gpurun_out/launches_<tag>.csv (ncu --metrics gpu__time_duration.sum,dram__bytes_* at the bench size),
       gpurun_out/full_<tag>_raw.csv (ncu --set full, 512 frames), gpurun_out/bench_<tag>.json
Writes profiles/<round>_launches.csv, profiles/<round>_ncu_full_summary.txt, profiles/<round>_bench.json,
       profiles/<round>_traffic.json (per kernel class: launches, ms, DRAM bytes — read by bench.py for roofline.traffic)
"""
import csv, hashlib, json, os, re, subprocess, sys
tag, rnd = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def source_digest():                       # = bench.py's: the traffic record is only quoted for the sources it was captured on
    h = hashlib.sha1()
    d = os.path.join(ROOT, "deepfake_video_detection_b200", "csrc")
    for f in sorted(os.listdir(d)):
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:12]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

def klass(name):
    if "mbconv_fused" in name: return "expand_dwconv_fused"   # bench.py's class name for the fused producer + depthwise kernel
    if "dwconv" in name: return "dwconv_se_squeeze"
    if "se_kernel" in name or "se_wide" in name: return "se_gate"
    if "head_pool_tc" in name: return "gemm_head_pool"
    if "stem" in name: return "stem"
    if "pool_head" in name: return "attn_pool_head"
    if "preprocess" in name: return "preprocess"
    if "scale_weights" in name: return "gemm_project"      # per-frame weight scaling belongs to the project step (bench.py counts it there)
    if "gemm_tc" in name:
        m = re.search(r"<[^,]+, *(\w+), *(\w+), *(\w+), *(\w+)", name)
        g, a, r, p = [x in ("1", "true") for x in m.groups()]
        return "gemm_head_pool" if p else ("gemm_expand" if a else "gemm_project")
    return "other"

rows = list(csv.reader(open(os.path.join(G, f"launches_{tag}.csv"))))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
per = {}
for r in rows[hi + 1:]:
    if len(r) <= vi: continue
    d = per.setdefault(r[ii], {"name": re.sub(r"^void dfd::|^dfd::", "", r[ki])})
    d[r[mi]] = float(r[vi].replace(",", ""))
    d["unit_" + r[mi]] = r[h.index("Metric Unit")]
def to_bytes(v, u): return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
def to_us(v, u): return v * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1e-3)
out, cls = [], {}
for k, d in per.items():
    t = to_us(d["gpu__time_duration.sum"], d["unit_gpu__time_duration.sum"])
    rd = to_bytes(d["dram__bytes_read.sum"], d["unit_dram__bytes_read.sum"]); wr = to_bytes(d["dram__bytes_write.sum"], d["unit_dram__bytes_write.sum"])
    c = klass(d["name"])
    out.append([k, c, d["name"][:80], f"{t:.1f}", f"{rd/1e6:.1f}", f"{wr/1e6:.1f}", f"{(rd+wr)/t/1e3:.0f}"])
    a = cls.setdefault(c, {"launches": 0, "us": 0.0, "dram_bytes": 0.0}); a["launches"] += 1; a["us"] += t; a["dram_bytes"] += rd + wr
with open(os.path.join(P, f"{rnd}_launches.csv"), "w", newline="") as f:
    w = csv.writer(f); w.writerow(["id", "class", "kernel", "time_us", "dram_read_MB", "dram_write_MB", "dram_GBps"]); w.writerows(out)
tot = sum(a["us"] for a in cls.values())
for a in cls.values(): a["share"] = round(a["us"] / tot, 4); a["dram_bytes_per_launch"] = a["dram_bytes"] / a["launches"]
json.dump({"source": f"ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, python tools/prof_step.py --videos 64 --frames 32 (2048 frames, one forward)",
           "source_digest": source_digest(), "total_us": tot, "classes": cls}, open(os.path.join(P, f"{rnd}_traffic.json"), "w"), indent=1)
txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_table.py"), os.path.join(G, f"full_{tag}_raw.csv")], capture_output=True, text=True).stdout
open(os.path.join(P, f"{rnd}_ncu_full_summary.txt"), "w").write(
    "ncu --set full --clock-control none, python tools/prof_step.py --videos 16 --frames 32 (512 frames), second forward pass, one line per launch\n"
    "(times are serialised / cold-cache: compare shares, not absolutes)\n\n" + txt)
b = json.loads(open(os.path.join(G, f"bench_{tag}.json")).read().strip().splitlines()[-1])
json.dump(b, open(os.path.join(P, f"{rnd}_bench.json"), "w"), indent=1)
print(json.dumps({k: {"us": round(v["us"], 1), "share": v["share"], "dramMB/launch": round(v["dram_bytes_per_launch"] / 1e6, 1)} for k, v in cls.items()}, indent=1))
print("bench kernels:", {k: v["ms"] for k, v in b["kernels"].items()})
