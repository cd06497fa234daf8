#!/usr/bin/env python
"""Prints the tables of BENCH.md sections 1 and 2 from a bench line (bench.py's JSON) and the full-size parity records
(tests/test_gpu_fullsize.py writes gpurun_out/fullsize_parity.jsonl), so that the document is copied from the
measurement files and not typed.
usage: python tools/bench_tables.py profiles/<bench line>.json profiles/<fullsize parity>.jsonl [previous bench line.json]"""
import json
import sys


def main():
    line = json.load(open(sys.argv[1]))
    par = {}
    if len(sys.argv) > 2:
        for l in open(sys.argv[2]):
            l = l.strip()
            if l:
                r = json.loads(l)
                par[r["config"]] = r
    prev = json.load(open(sys.argv[3])) if len(sys.argv) > 3 else None
    c, e, rf = line["config"], line["e2e"], line["roofline"]
    rays = c["rays_per_frame"]
    total = rays["primary"] + rays["shadow"] + rays["reflection"]
    print("## headline (%s, %dx%d x %d spp)" % (c["workload"], c["width"], c["height"], c["spp"]))
    print("rays: %d primary + %d shadow + %d reflection = %.0f Mrays" % (rays["primary"], rays["shadow"], rays["reflection"], total / 1e6))
    print("value: %.2f ms/frame = %.0f Mrays/s%s" % (line["ms_per_step"], line["value"],
                                                       "  (previous: %.2f ms = %.0f Mrays/s)" % (prev["ms_per_step"], prev["value"]) if prev else ""))
    print("e2e (%s, %s): %.2f ms/frame = %.0f Mrays/s, e2e/value = %.3f, d2h %.1f MB" % (e["host_buffer"], e["out_format"].split(" ")[0], e["ms_per_step"], e["value"],
                                                                                        e["value"] / line["value"], e["d2h_bytes_per_step"] / 1e6))
    print("e2e f64: %.2f ms/frame (%.0f MB)" % (e["f64"]["ms_per_step"], e["f64"]["d2h_bytes_per_step"] / 1e6))
    if "cpu_baseline" in line:
        cb = line["cpu_baseline"]
        print("cpu_baseline: %.2f Mrays/s on %d threads (%s)" % (cb["value"], cb["cores"], cb["sample"]))
    if "parity" in line:
        p = line["parity"]
        print("parity on the CPU sample: %.5f %% of %d pixels within 1/255, max err %.4f; prim-id mismatches %d (unexplained %d) of %d samples"
              % (100 * p["frac_within_1_255"], p["pixels"], p["max_err"], p["prim_id_mismatches"], p["prim_id_unexplained"], p["primary_samples"]))
    print("kernel: %.2f ms; %.1f GFLOP/launch => %.2f TFLOP/s = %.1f %% of %.1f TFLOP/s; traffic %s" % (
        rf["kernel_ms"], rf["algorithmic_flops_per_launch"] / 1e9, rf["achieved"], 100 * rf["frac"], rf["peak"], rf.get("traffic")))
    if "reference_algorithm" in rf:
        print("reference algorithm: %d flops/ray -> %.1f TFLOP/s equivalent" % (rf["reference_algorithm"]["intersection_flops_per_ray"], rf["reference_algorithm"]["equivalent_tflops"]))
    print("clocks:", line.get("clocks"))
    print()
    print("| config | W×H×spp | ms/frame (before) | Mrays/s | e2e ms | roofline frac | within 1/255 | max err | prim-id mismatches | unexplained |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    per = dict(line.get("per_config", {}))
    per[c["workload"]] = dict(width=c["width"], height=c["height"], spp=c["spp"], ms_per_step=line["ms_per_step"], value=line["value"], e2e_ms_per_step=e["ms_per_step"],
                              roofline_frac=rf["frac"])
    pper = dict(prev.get("per_config", {})) if prev else {}
    if prev:
        pper[prev["config"]["workload"]] = dict(ms_per_step=prev["ms_per_step"])
    for name in ["cfg1-sample", "cfg2-hollow-sphere", "cfg3-house", "cfg3-night-house", "cfg4-bunny", "cfg4-bunny-d12", "cfg4-bunny-full-d14", "cfg5-moon", "cfg5-repeat"]:
        r = per.get(name)
        if not r or "see" in r:
            continue
        q = par.get(name, {})
        before = " (%.3f)" % pper[name]["ms_per_step"] if name in pper and "ms_per_step" in pper[name] else ""
        print("| %s | %d×%d×%d | %.3f%s | %.0f | %.3f | %.3f | %s | %s | %s | %s |" % (
            name, r["width"], r["height"], r["spp"], r["ms_per_step"], before, r["value"], r["e2e_ms_per_step"], r["roofline_frac"],
            "%.6f" % q["frac_within_1_255"] if q else "-", "%.3g" % q["max_err"] if q else "-",
            "%d / %.1f M" % (q["prim_id_mismatches"], q["primary_samples"] / 1e6) if q else "-",
            ("%d" % q["prim_id_unexplained"]) + (" (%d edge leak)" % q["prim_id_edge_leaks"] if q.get("prim_id_edge_leaks") else "") if q else "-"))
    cr = [(n, r.get("scene_create_ms")) for n, r in line.get("per_config", {}).items() if isinstance(r, dict) and "scene_create_ms" in r]
    if cr:
        print("\nscene_create_ms: " + ", ".join("%s %.1f" % (n, v) for n, v in cr) + "; %s %.1f" % (c["workload"], e["scene_create_ms"]))


if __name__ == "__main__":
    main()
