#!/usr/bin/env python
"""Turn the .ncu-rep files of tools/r02_p.sh into the tracked summaries:
   profiles/<tag>_ncu_full_config4.md   (tools/ncu_summary.py + the 45 hottest source lines of tools/ncu_hot_lines.py per kernel)
   profiles/<tag>_traffic.json          (DRAM bytes per average launch + lanes / IPC / hit rates; read by bench.py)
usage: python tools/make_profiles.py gpurun_out/r02p r02 <vis_ms_per_step> <pairs_ms_per_step> [launches_per_step=7]"""
import csv
import io
import json
import subprocess
import sys

src, tag, vis_ms, pairs_ms = sys.argv[1], sys.argv[2], float(sys.argv[3]), float(sys.argv[4])
launches = int(sys.argv[5]) if len(sys.argv) > 5 else 7
kernels = (("rad_visibility_kernel", "rad_visibility"), ("rad_candidates_kernel", "rad_candidates"), ("direct_march_kernel", "direct_march"))
md = []
for k, rep in kernels:
    md.append(subprocess.run([sys.executable, "tools/ncu_summary.py", f"{src}_{rep}.ncu-rep"], capture_output=True, text=True).stdout)
    hot = subprocess.run([sys.executable, "tools/ncu_hot_lines.py", f"{src}_{rep}.ncu-rep"], capture_output=True, text=True).stdout.splitlines()
    md.append("```\n" + "\n".join(l[:200] for l in hot[:47]) + "\n```\n")
open(f"profiles/{tag}_ncu_full_config4.md", "w").write("\n".join(md))

out = {"config4": {}, "ncu": {"config4": {}}, "captured": [],
       "how": "dram__bytes_read.sum + dram__bytes_write.sum of ONE ncu --set full capture per kernel on config 4 (tools/r02_p.sh); the radiosity kernels run "
              f"{launches} launches of different sizes per bake, so their figure is (captured bytes / captured ms) x the average launch duration of the bench run "
              f"({vis_ms} / {launches} and {pairs_ms} / {launches} ms), i.e. per AVERAGE launch like roofline.achieved"}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
tscale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
avg = {"rad_visibility_kernel": vis_ms / launches, "rad_candidates_kernel": pairs_ms / launches, "direct_march_kernel": None}
side_metrics = (("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes_per_warp_instruction"), ("sm__inst_executed.avg.per_cycle_elapsed", "ipc_per_sm"),
                ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occupancy_pct"),
                ("lts__t_sector_hit_rate.pct", "l2_hit_pct"), ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"), ("launch__registers_per_thread", "registers"))
for k, rep in kernels:
    raw = subprocess.run(["ncu", "-i", f"{src}_{rep}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, r = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    val = lambda name: (float(r[col[name]].replace(",", "")), units[col[name]])
    rd, u1 = val("dram__bytes_read.sum"); wr, u2 = val("dram__bytes_write.sum"); t, tu = val("gpu__time_duration.sum")
    nbytes, ms = rd * scale[u1] + wr * scale[u2], t * tscale[tu]
    out["captured"].append({"kernel": k, "ms_under_ncu": ms, "dram_bytes": nbytes})
    out["config4"][k] = nbytes * (avg[k] / ms) if avg[k] else nbytes
    out["ncu"]["config4"][k] = {label: val(name)[0] for name, label in side_metrics if name in col and r[col[name]]}
json.dump(out, open(f"profiles/{tag}_traffic.json", "w"), indent=1)
print(json.dumps(out["config4"]))
