"""Rewrites the measured tables of BASELINE.md section 3 from the bench lines under profiles/ (python tools/update_baseline_md.py)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles", "r2_bench_%s.json")


def L(f):
    return json.loads(open(P % f).read())


def num(x):
    if isinstance(x, dict):
        x = x.get("value")
    return "%.0f" % x if x else "—"


def row(f):
    d = L(f)
    e = d.get("e2e") or {}
    di = e.get("drop_in_call") or {}
    cpu = d.get("cpu_baseline") or {}
    r = d["roofline"]
    drop = "%s / %s" % (num(di.get("pageable_input")), num(di.get("pinned_input"))) if di.get("pageable_input") else "—"
    return dict(v="%.0f" % d["value"], ms="%.3f" % d["ms_per_step"], e2e="%.0f" % e.get("value", 0), drop=drop,
                c1=("%.0f" % cpu["single_thread_value"]) if cpu.get("single_thread_value") else "—",
                ca=("%.0f (%s)" % (cpu["value"], cpu["cores"])) if cpu.get("value") else "—", gb="%.0f" % r["achieved"], fr="%.1f %%" % (100 * r["frac"]))


def main():
    r = {k: row(k + "_1gpu") for k in ("c1", "c2", "c3", "c4a", "c4a_x16", "c4b", "c5", "c5o", "c3o", "x_ycbcr", "x_ycck", "x_cmyk", "x_rgb411", "x_rgb444") if os.path.exists(P % (k + "_1gpu"))}
    T = ("| Config (BASELINE.json) | GPU MP/s, inputs in HBM, 1 B200 | ms / step | round 1 ms | e2e MP/s (pinned host pixels → host JPEG) | "
         "drop-in call MP/s, pageable / pinned input | CPU 1 thread MP/s | CPU all cores MP/s (cores) | stage-A GB/s | stage-A % of measured HBM peak (6541.1 GB/s) |\n"
         "|---|---|---|---|---|---|---|---|---|---|\n")

    def line(name, k, r1, note="", drop_note=""):
        x = r[k]
        return "| %s | %s | %s | %s | %s | %s%s | %s | %s | %s | %s%s |\n" % (name, x["v"], x["ms"], r1, x["e2e"], x["drop"], drop_note, x["c1"], x["ca"], x["gb"], x["fr"], note)
    x16 = r["c4a_x16"]
    T += line("C1 1920×1080 RGB q90 `F_2_2` baseline", "c1", "0.168", " (a 17 µs launch)")
    T += line("C2 4096² RGB q85 4:2:0 optimized, restart 64", "c2", "0.315", " (a 37 µs launch)")
    T += line("C3 1024 × C1, fixed batch sharded by image", "c3", "7.392", "", " (one frame per call)")
    T += line("C4a 8192² Luma q95 custom tables", "c4a", "0.304", " (a 58 µs launch; **%s** for 16 such images: %s MP/s)" % (x16["fr"], x16["v"]))
    T += line("C4b 8192² `CmykAsYcck` q95 4:4:4 custom tables", "c4b", "0.693", " (four components of arithmetic per pixel: issue-bound)")
    T += line("C5 16384² RGB progressive 4:2:0, restart 2048 (strips)", "c5", "1.927")
    T += line("C5o 16384² RGB 4:2:0 optimized tables, restart 2048 (strips + histogram exchange; not a BASELINE config)", "c5o", "1.993")
    T += line("C3o 1024 × 1080p with per-image optimized tables (histogram + K.2 on the device; not a BASELINE config)", "c3o", "—")
    T += line("YCbCr 4:2:0 verbatim, 256 × 1080p", "x_ycbcr", "(generic kernel)")
    T += line("YCCK 4:4:4 verbatim, 16 × 4096²", "x_ycck", "(generic kernel)")
    T += line("CMYK (inverted), K at 2×2, 16 × 4096²", "x_cmyk", "(generic kernel)")
    T += line("RGB 4:1:1 (factor 4, sequential scans), 256 × 1080p", "x_rgb411", "(generic kernel: 7.320)")
    if "x_rgb444" in r:
        T += line("RGB 4:4:4, 256 × 1080p", "x_rgb444", "—")
    path = os.path.join(ROOT, "BASELINE.md")
    s = open(path).read()
    a = s.index("| Config (BASELINE.json) | GPU MP/s")
    b = s.index("Stage times of the default line")
    s = s[:a] + T + "\n" + s[b:]
    c3 = L("c3_1gpu")
    st = c3["stage_ms_per_step"]
    a = s.index("Stage times of the default line")
    b = s.index("(round 1: 2.90, 2.79, 0.86, 0.72)")
    s = s[:a] + "Stage times of the default line (C3, ms): colour+DCT+quant %.2f, coding + prefix sums %.2f, placement %.2f, stuffing\n%.2f " % (
        st["colour_dct_quant"], st["code_chunks_scans"], st["place_chunks"], st["stuff_scatter"]) + s[b:]
    open(path, "w").write(s)


if __name__ == "__main__":
    main()
