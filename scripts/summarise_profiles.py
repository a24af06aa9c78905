"""Turns the ncu outputs gpurun brought back (gpurun_out/<round>_*) into the committed summaries under profiles/:
  <round>_launches_<workload>.csv        per-kernel launch list (cold-cache, serialised: compare SHARES)
  <round>_ncu_full_<workload>.csv        key metrics of the full captures, one row per launch
  <round>_sass_<workload>.txt            opcode mix / stall reasons of the hot kernels
  traffic.json                           dram bytes (read+write) per launch per stage, read by bench.py
Usage: python scripts/summarise_profiles.py r2"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r2"
OUT = os.path.join(ROOT, "profiles")
GO = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
STAGE = {"k_mm2meters": "preprocess", "k_alloc_sdf": "alloc", "k_alloc_ofusion": "alloc", "k_alloc_first_key_chain": "alloc",
         "k_filter_blocks": "alloc", "k_integrate_sdf": "fuse", "k_integrate_ofusion": "fuse", "k_raycast": "raycast",
         "k_render_shade": "render", "k_render_volume": "render"}
WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__inst_executed_pipe_uniform.sum"]


def short(name):
    n = name.split("(")[0].replace("void ", "").replace("se_b200::", "")
    return n.split("<")[0]


def launches(workload):
    src = os.path.join(GO, f"{R}_launches_{workload}.csv")
    if not os.path.exists(src):
        return
    rows = list(csv.reader(open(src)))
    hdr = next(r for r in rows if "Kernel Name" in r)
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows:
        if len(r) == len(hdr) and r[ix["Metric Name"]] == "gpu__time_duration.sum":
            agg.setdefault(short(r[ix["Kernel Name"]]), []).append(float(r[ix["Metric Value"]].replace(",", "")))
    ours = {k: v for k, v in agg.items() if k.startswith("k_")}
    tot = sum(sum(v) for v in ours.values())
    with open(os.path.join(OUT, f"{R}_launches_{workload}.csv"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --workload %s --steps 4 --warmup 3 --no-cpu-baseline --no-extra\n" % workload)
        f.write("# per-launch times are cold-cache and serialised under ncu: compare the SHARES with bench.py's stage times, not the absolutes\n")
        f.write("kernel,launches,mean_us,total_us,share_of_our_kernels_pct\n")
        for k, v in sorted(ours.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{k},{len(v)},{sum(v)/len(v)/1000:.3f},{sum(v)/1000:.2f},{100*sum(v)/tot:.1f}\n")
        other = {k: v for k, v in agg.items() if not k.startswith("k_")}
        f.write("# kernels that are not ours in the same process (torch: L2 flush fill, copies)\n")
        for k, v in sorted(other.items(), key=lambda kv: -sum(kv[1]))[:6]:
            f.write(f"# {k[:70]},{len(v)},{sum(v)/len(v)/1000:.3f}\n")


def full(workload, traffic):
    rep = os.path.join(GO, f"{R}_full_{workload}.ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    idx = [hdr.index(w) if w in hdr else -1 for w in WANT]
    per_stage = collections.defaultdict(list)
    with open(os.path.join(OUT, f"{R}_ncu_full_{workload}.csv"), "w") as f:
        w = csv.writer(f)
        f.write(f"# ncu --set full --clock-control none (cache control: flush before each replay) on scripts/profile_frames.py {workload}: one frame's kernels at the bench's operating point (scripts/make_profiles.sh)\n")
        w.writerow(WANT)
        w.writerow([rows[1][i] if i >= 0 else "" for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] if i >= 0 else "" for i in idx])
            name = short(r[hdr.index("Kernel Name")])
            unit_r, unit_w = rows[1][hdr.index("dram__bytes_read.sum")], rows[1][hdr.index("dram__bytes_write.sum")]
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            b = float(r[hdr.index("dram__bytes_read.sum")]) * scale[unit_r] + float(r[hdr.index("dram__bytes_write.sum")]) * scale[unit_w]
            per_stage[(STAGE.get(name, name), name)].append(b)
    t = collections.defaultdict(float)
    for (stage, name), v in per_stage.items():
        t[stage] += sum(v) / len(v)                 # mean per launch, summed over the kernels of the stage
    traffic[workload] = {k: int(v) for k, v in t.items()}
    with open(os.path.join(OUT, f"{R}_sass_{workload}.txt"), "w") as f:
        for kern in ("k_raycast", "k_alloc_sdf", "k_alloc_ofusion", "k_filter_blocks", "k_integrate_sdf", "k_integrate_ofusion"):
            src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{kern}"], capture_output=True, text=True).stdout
            if "Address" not in src:
                continue
            out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_sass_summary.py"), "8"], input=src, capture_output=True, text=True).stdout
            f.write(out + "\n")


traffic = {}
tp = os.path.join(OUT, "traffic.json")
if os.path.exists(tp):
    traffic = json.load(open(tp))
for wl in ("planar_sweep_sdf512", "box_room_sdf2048", "box_room_ofusion1024"):
    launches(wl)
    full(wl, traffic)
traffic["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over captured launches), summed over the kernels of a stage; from profiles/%s_ncu_full_*.csv "
                    "(one frame in the middle of the frames bench.py times: scripts/make_profiles.sh)" % R)
json.dump(traffic, open(tp, "w"), indent=1, sort_keys=True)
print(json.dumps(traffic, indent=1))
