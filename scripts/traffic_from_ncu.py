#!/usr/bin/env python
"""profiles/traffic.json from the `ncu --set full` captures of one GPU pass (scripts/gpu_pass.sh), read on the CPU box:
per kernel dram__bytes_read.sum + dram__bytes_write.sum and smsp__inst_executed.sum of ONE launch — what bench.py reports as
roofline.traffic and uses for the issue roofline.      usage: python scripts/traffic_from_ncu.py <tag, e.g. r02b>"""
import csv, io, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
out = {}
used = []
for rep in ("prof_step_random", "prof_mcts_search", "prof_net_acc", "prof_tree", "prof_net"):
    path = os.path.join(ROOT, "gpurun_out", rep + ".ncu-rep")
    if not os.path.exists(path):
        continue
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = rows[0]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = re.sub(r"[<(].*", "", d["Kernel Name"]).replace("void ", "").strip()
        num = lambda k: float(d[k].replace(",", "")) if d.get(k) else 0.0
        unit = dict(zip(hdr, rows[1]))
        scale = lambda k: {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(unit.get(k, "byte"), 1.0)
        dram = num("dram__bytes_read.sum") * scale("dram__bytes_read.sum") + num("dram__bytes_write.sum") * scale("dram__bytes_write.sum")
        out[name] = int(dram)
        out[name + "_warp_insts"] = int(num("smsp__inst_executed.sum"))
        out[name + "_smem_wavefronts"] = int(num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"))
        out[name + "_smem_conflicts"] = int(num("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"))
        out[name + "_launch_us"] = num("gpu__time_duration.sum") * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit.get("gpu__time_duration.sum"), 1.0)
    used.append(rep + ".ncu-rep")
# bench.py looks the step kernel up under this key whatever variant is the default
for k in list(out):
    if k.startswith("k_step_random") and not k.endswith(("_warp_insts", "_launch_us", "_smem_wavefronts", "_smem_conflicts")):
        out["k_step_random_flat"] = out[k]
        out["k_step_random_flat_warp_insts"] = out[k + "_warp_insts"]
        out["k_step_random_flat_smem_wavefronts"] = out[k + "_smem_wavefronts"]
        out["k_step_random_flat_smem_conflicts"] = out[k + "_smem_conflicts"]
        out["step_kernel"] = k
out["source"] = "ncu --set full captures of the %s GPU pass (%s); per launch: dram__bytes_read.sum + dram__bytes_write.sum, smsp__inst_executed.sum, l1tex__data_pipe_lsu_wavefronts_mem_shared.sum" % (tag, ", ".join(used))
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
