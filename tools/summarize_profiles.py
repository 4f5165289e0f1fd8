"""Turns the raw output of tools/measure_round.sh (gpurun_out/<tag>_*) into the tracked summaries under profiles/.

    python tools/summarize_profiles.py r1f r1

Reads the .ncu-rep files with `ncu -i ... --page raw --csv` (no GPU needed). Numbers taken under ncu are
never bench values: the bench lines are copied from the un-profiled bench.py runs of the same pass.
"""
import collections
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out")
DST = os.path.join(ROOT, "profiles")

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_static",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def ncu_raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {}
    for h, u, v in zip(hdr, units, vals):
        if h in WANT or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
            try:
                m[h] = {"value": float(v.replace(",", "")), "unit": u}
            except ValueError:
                pass
    return m, vals[hdr.index("Kernel Name")]


def ncu_raw_all(rep):
    """[(metrics, kernel name)] for every launch in the report (or in its exported raw page, <name>.raw.csv)"""
    csv_path = rep[:-len(".ncu-rep")] + ".raw.csv"
    if os.path.exists(csv_path):
        out = open(csv_path).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        if len(vals) != len(hdr):
            continue
        m = {}
        for h, u, v in zip(hdr, units, vals):
            if h in WANT or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
                try:
                    m[h] = {"value": float(v.replace(",", "")), "unit": u}
                except ValueError:
                    pass
        res.append((m, vals[hdr.index("Kernel Name")]))
    return res


def have(rep):
    return os.path.exists(rep) or os.path.exists(rep[:-len(".ncu-rep")] + ".raw.csv")


def mbytes(m, key):
    v = m[key]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[v["unit"]]
    return v["value"] * scale


def main():
    tag, out_tag = sys.argv[1], sys.argv[2]
    # bench lines
    for f in sorted(os.listdir(SRC)):
        if f.startswith(tag + "_bench_") and f.endswith(".json"):
            lines = [l for l in open(os.path.join(SRC, f)) if l.startswith("{")]  # torchrun / NCCL banners share stdout
            if lines:
                open(os.path.join(DST, out_tag + f[len(tag):]), "w").write(lines[-1])
    for f in ("compute_sanitizer_memcheck.txt", "compute_sanitizer_racecheck.txt", "launches_c3_b64.csv"):
        p = os.path.join(SRC, "%s_%s" % (tag, f))
        if os.path.exists(p):
            shutil.copy(p, os.path.join(DST, "%s_%s" % (out_tag, f)))
    # launch list
    p = os.path.join(SRC, tag + "_launches_c3_b64.csv")
    if os.path.exists(p):
        rows = list(csv.reader(open(p)))
        hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
        hdr = rows[hi]
        kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        agg = collections.defaultdict(list)
        for r in rows[hi + 1:]:
            if len(r) <= mv:
                continue
            v = float(r[mv].replace(",", ""))
            v = v / 1000 if r[mu] == "ns" else (v * 1000 if r[mu] == "ms" else v)
            agg[re.sub(r"\(.*", "", r[kn]).replace("jpgb::<unnamed>::", "").replace("void ", "")[:48]].append(v)
        tot = sum(sum(v) for v in agg.values())
        with open(os.path.join(DST, out_tag + "_launches_c3_b64_summary.txt"), "w") as f:
            f.write("ncu --metrics gpu__time_duration.sum --clock-control none --csv python bench.py --workload c3 --batch 64 --steps 2 --warmup 1 --no-cpu\n")
            f.write("(all launches of the process: parity gate, warm-up, timed steps and the chunked e2e calls; cold-cache and serialised,\n"
                    " so only the SHARES are comparable with the live stage timers of the bench line)\n")
            for k, v in sorted(agg.items(), key=lambda x: -sum(x[1])):
                f.write("%-50s n=%3d %10.1f us %5.1f%%  largest launch %8.1f us\n" % (k, len(v), sum(v), 100 * sum(v) / tot, max(v)))
    # full captures: one report with one launch of every dominant kernel (C3, 64 frames per launch)
    frames = 64
    alg = 1920 * 1080 * 3 + 128 * 48960
    rep = os.path.join(SRC, tag + "_full.ncu-rep")
    if have(rep):
        cmd = ("ncu --set full --clock-control none --import-source on -k 'regex:stage_a_warp|encode_chunks|place_chunks|stuff_scatter|count_ff' "
               "-s 10 -c 5 python bench.py --workload c3 --batch 64 --steps 2 --warmup 1 --no-cpu --no-c5")
        for m, name in ncu_raw_all(rep):
            short = re.sub(r"\(.*", "", name).replace("void ", "").split("::")[-1].split("<")[0].strip()
            rec = {"kernel": name, "command": cmd, "frames_per_launch": frames, "metrics": m,
                   "note": "cold-cache single launch under ncu; live times are in the bench lines"}
            if "dram__bytes_read.sum" in m:
                tr = mbytes(m, "dram__bytes_read.sum") + mbytes(m, "dram__bytes_write.sum")
                rec.update(dram_bytes_per_launch=tr, dram_bytes_per_frame=tr / frames)
            if short.startswith("stage_a"):
                rec["algorithmic_bytes_per_frame"] = alg
                rec["sass"] = ("UTMALDG.3D (TMA), SYNCS.ARRIVE.TRANS64 / SYNCS.PHASECHK.TRANS64.TRYWAIT (mbarrier), IDP.2A (colour), "
                               "STG.E.ENL2.256 (stores): cuobjdump -sass jpeg_encoder_b200/build/stage_a.cu.o")
                json.dump({"workload": "c3", "dram_bytes_per_frame": rec["dram_bytes_per_frame"],
                           "source": "profiles/%s_stage_a_warp_kernel_ncu_summary.json" % out_tag}, open(os.path.join(DST, "stage_a_traffic.json"), "w"), indent=1)
            if short.startswith("encode_chunks"):
                visits = frames * 48960
                rec["visits_per_launch"] = visits
                rec["warp_instructions_per_32_visits"] = m["smsp__inst_executed.sum"]["value"] / (visits / 32)
                rec["algorithmic_bytes_per_frame"] = 128 * 48960
            json.dump(rec, open(os.path.join(DST, "%s_%s_ncu_summary.json" % (out_tag, short)), "w"), indent=1)
    rep = os.path.join(SRC, tag + "_coder_c5.ncu-rep")
    if have(rep):
        m, name = ncu_raw_all(rep)[0]
        blocks = (16384 // 8) ** 2 + 2 * (16384 // 16) ** 2  # 16384x16384 4:2:0: true grids of Y, Cb, Cr
        json.dump({"kernel": name, "workload": "c5", "metrics": m, "blocks": blocks,
                   "dram_bytes_read_per_block": mbytes(m, "dram__bytes_read.sum") / blocks,
                   "note": "one 16384x16384 progressive image per launch (DC scan + AC bands of every component from one staging of each block), cold cache, under ncu"},
                  open(os.path.join(DST, "%s_encode_chunks_c5_ncu_summary.json" % out_tag), "w"), indent=1)
    rep = os.path.join(SRC, tag + "_histogram_c3o.ncu-rep")
    if have(rep):
        m, name = ncu_raw_all(rep)[0]
        json.dump({"kernel": name, "workload": "c3o, 64 frames per launch", "metrics": m, "note": "cold cache, under ncu"},
                  open(os.path.join(DST, "%s_histogram_kernel_ncu_summary.json" % out_tag), "w"), indent=1)
    for f in ("planar_stage_a.txt", "compute_sanitizer_initcheck.txt"):
        p = os.path.join(SRC, "%s_%s" % (tag, f))
        if os.path.exists(p):
            shutil.copy(p, os.path.join(DST, "%s_%s" % (out_tag, f)))
    for w in ("c4a", "c4b", "x_rgb444"):
        rep = os.path.join(SRC, "%s_stage_a_%s.ncu-rep" % (tag, w))
        if have(rep):
            m, name = ncu_raw_all(rep)[0]
            json.dump({"kernel": name, "workload": w, "metrics": m,
                       "note": ("one 8192x8192 image per launch" if w.startswith("c4") else "64 frames of 1920x1080 per launch") + ", cold cache, under ncu"},
                      open(os.path.join(DST, "%s_stage_a_%s_ncu_summary.json" % (out_tag, w)), "w"), indent=1)


if __name__ == "__main__":
    main()
