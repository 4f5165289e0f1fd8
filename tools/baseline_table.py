"""Prints the rows of BASELINE.md section 3 from the bench lines under profiles/ (python tools/baseline_table.py r2)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROWS = [("c1", "C1"), ("c2", "C2"), ("c3", "C3"), ("c4a", "C4a"), ("c4a_x16", "C4a x16"), ("c4b", "C4b"), ("c5", "C5"), ("c5o", "C5o"), ("c3o", "C3o"),
        ("x_ycbcr", "YCbCr 4:2:0"), ("x_ycck", "YCCK 4:4:4"), ("x_cmyk", "CMYK"), ("x_rgb411", "RGB 4:1:1"), ("x_rgb444", "RGB 4:4:4")]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    print("| id | GPU MP/s | ms/step | e2e MP/s | drop-in pageable / pinned | CPU 1 thr | CPU all (cores) | stage-A GB/s | % of HBM peak | stage ms |")
    for wid, name in ROWS:
        p = os.path.join(ROOT, "profiles", "%s_bench_%s_1gpu.json" % (tag, wid))
        if not os.path.exists(p):
            continue
        d = json.loads(open(p).read())
        e = d.get("e2e") or {}
        di = e.get("drop_in_call") or {}
        cpu = d.get("cpu_baseline") or {}
        r = d.get("roofline") or {}
        st = {k: round(v, 3) for k, v in (d.get("stage_ms_per_step") or {}).items() if v}

        def mp(x):
            return "%.0f" % x["value"] if isinstance(x, dict) and x.get("value") else ("%.0f" % x if isinstance(x, (int, float)) and x else "-")
        print("| %s | %.0f | %.4f | %s | %s / %s | %s | %s (%s) | %.0f | %.1f %% | %s |" % (
            name, d["value"], d["ms_per_step"], mp(e), mp(di.get("pageable_input")), mp(di.get("pinned_input")),
            ("%.0f" % cpu["single_thread_value"]) if cpu.get("single_thread_value") else "-", mp(cpu), cpu.get("cores", "-"),
            r.get("achieved") or 0, 100 * (r.get("frac") or 0), st))
    for n in (2, 4, 8):
        p = os.path.join(ROOT, "profiles", "%s_bench_c3_%dgpu.json" % (tag, n))
        if os.path.exists(p):
            d = json.loads(open(p).read())
            print("N=%d: strong %.0f MP/s (%.3f ms) weak %s e2e %.0f c5 %s rank_ms %s" % (
                n, d["value"], d["ms_per_step"], (d.get("weak") or {}).get("value"), (d.get("e2e") or {}).get("value", 0),
                {k: (d.get("c5") or {}).get(k) for k in ("value", "ms_per_step", "gather_ms", "e2e")}, d.get("rank_ms_per_step")))


if __name__ == "__main__":
    main()
