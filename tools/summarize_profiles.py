#!/usr/bin/env python
"""Turn a tools/prof_run.sh capture (gpurun_out/<tag>_*) into the committed summaries under profiles/:
   profiles/<tag>_launches.md   launch-list shares (ncu --metrics gpu__time_duration.sum) + bench category timing
   profiles/<tag>_ncu_full.md   highlights of the --set full captures of the hot kernels
   profiles/ncu_traffic.json    measured DRAM bytes per launch per bench.py kernel category (roofline.traffic)
Runs here (no GPU): `ncu -i` only reads the reports."""
import csv, io, json, os, subprocess, sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def launches():
    rows = []
    with open(os.path.join(G, tag + "_launches.csv")) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rd:
        if r[im] != "gpu__time_duration.sum":
            continue
        name = r[ik].split("(")[0].replace("void ", "")
        agg[name][0] += 1
        agg[name][1] += float(r[iv].replace(",", "")) / 1e6  # ns -> ms
    tot = sum(v[1] for v in agg.values())
    out = ["| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| `%s` | %d | %.2f | %.1f%% |" % (k[:70], n, ms, 100 * ms / tot))
    return out, tot, sum(v[0] for v in agg.values())


def raw(rep):
    txt = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


WANT = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %"), ("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe %"),
        ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue %"), ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem (tensor) %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1/smem %"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs")]
CAT = {"k_layer_fwd_p_umma": "layer_fwd", "k_layer_bwd_fused_umma": "layer_bwd", "k_post_fwd_umma": "post_fwd_loss",
       "k_post_bwd_umma": "post_bwd", "k_wgrad_umma": "wgrad"}


def main():
    os.makedirs(P, exist_ok=True)
    lt, tot, n = launches()
    bench = json.load(open(os.path.join(G, tag + "_bench.json")))
    with open(os.path.join(P, tag + "_bench.json"), "w") as f:
        json.dump(bench, f, indent=1)
    with open(os.path.join(P, tag + "_launches.md"), "w") as f:
        f.write("# %s: ncu launch list of `python bench.py --steps 1 --warmup 3 --no-gen --no-cpu`\n\n" % tag)
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv` (4 training steps of BASELINE configs[1]; "
                "per-launch times are cold-cache and serialised: compare SHARES).\n\n")
        f.write("\n".join(lt) + "\n\nTotal %.2f ms over %d launches.\n\n" % (tot, n))
        f.write("CUDA-event category timing of the un-profiled bench run of the same build (ms per step, `kernel_shares`):\n\n")
        for k, v in bench["kernel_shares"].items():
            f.write("* %s: %.3f ms, %d launches\n" % (k, v["ms_per_step"], v["launches_per_step"]))
        f.write("\nstep %.3f ms, %.4g timesteps/s (e2e %.4g)\n" % (bench["ms_per_step"], bench["value"], bench["e2e"]["value"]))
    traffic = {}
    with open(os.path.join(P, tag + "_ncu_full.md"), "w") as f:
        f.write("# %s: `ncu --set full --clock-control none --import-source on` highlights\n\n" % tag)
        f.write("| kernel | " + " | ".join(w[1] for w in WANT) + " |\n|---|" + "---:|" * len(WANT) + "\n")
        for rep in ("_post", "_lfwd", "_lbwd"):
            path = tag + rep + ".ncu-rep"
            if not os.path.exists(os.path.join(G, path)):
                continue
            hdr, units, rows = raw(path)
            ik = hdr.index("Kernel Name")
            for r in rows:
                name = r[ik].split("(")[0].replace("void ", "").replace("wn::", "")
                cells = []
                for m, _ in WANT:
                    if m in hdr:
                        i = hdr.index(m)
                        cells.append("%s %s" % (r[i], units[i]) if units[i] not in ("", "%") else r[i])
                    else:
                        cells.append("-")
                f.write("| `%s` | " % name[:40] + " | ".join(cells) + " |\n")
                base = name.split("<")[0]
                if base in CAT and CAT[base] not in traffic:
                    def val(m):
                        i = hdr.index(m)
                        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
                        return float(r[i]) * mult
                    traffic[CAT[base]] = {"dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                                          "dram_read_bytes": val("dram__bytes_read.sum"), "dram_write_bytes": val("dram__bytes_write.sum"),
                                          "kernel": name, "capture": path, "note": "one launch, steady state (4th training step)"}
    with open(os.path.join(P, "ncu_traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1)
    print(open(os.path.join(P, tag + "_ncu_full.md")).read())
    print(json.dumps(traffic, indent=1))


main()
