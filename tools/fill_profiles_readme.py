"""profiles/README.md from tools/profiles_readme.tmpl.md + the committed bench line and launch list of the round."""
import collections, csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = json.loads([l for l in open(os.path.join(ROOT, "profiles/r02_bench_B4096.json")) if l.startswith("{")][0])
o, r = d["other_configs"], d["roofline"]
M = lambda v: "%.2f" % (v / 1e6)
sub = {
    "VALUE": M(d["value"]), "MS": "%.3f" % d["ms_per_step"], "E2E": M(d["e2e"]["value"]), "E2EMS": "%.3f" % d["e2e"]["ms_per_step"],
    "E2EPCT": "%.0f" % (100 * d["e2e"]["value"] / d["value"]),
    "E2EUNPACK": M(d["e2e_u16rows_unpack_target"]["value"]), "E2EU16": M(d["e2e_u16_target"]["value"]),
    "E2EF32": M(d["e2e_f32_target"]["value"]), "CPU": "%.0f" % d["cpu_baseline"]["value"],
    "RATIO": "%d" % (round(d["value"] / d["cpu_baseline"]["value"], -2)),
    "RATIOE2E": "%d" % (round(d["e2e"]["value"] / d["cpu_baseline"]["value"], -2)),
    "STEPFRAC": "%.1f" % (100 * r["step"]["frac"]), "RMS": "%.3f" % r["kernel_ms"], "RGBS": "%.0f" % r["achieved"],
    "RFRAC": "%.1f" % (100 * r["frac"]),
    "B1024MS": "%.3f" % o["batch1024"]["ms_per_step"], "B1024F": M(o["batch1024"]["fits_per_s"]),
    "B1024FRAC": "%.1f" % (100 * o["batch1024"]["step_hbm_frac"]),
    "B128MS": "%.3f" % o["C1_batch128"]["ms_per_step"], "B128F": M(o["C1_batch128"]["fits_per_s"]),
    "C3MS": "%.2f" % o["C3_multiview_256_batch512"]["fused_ms"], "C3F": "%.0f" % (o["C3_multiview_256_batch512"]["fused_fits_per_s"] / 1e3),
    "ICPMS": "%.2f" % o["C4_icp_batch1024"]["ms_fwd_bwd"], "ICPFWD": "%.2f" % o["C4_icp_batch1024"]["ms_fwd"],
    "COLLMS": "%.3f" % o["C4_coll_batch1024"]["ms_fwd_bwd"],
    "API128": "%.2f" % o["drop_in_api_literal_batch128"]["ms_per_step"], "API1024": "%.2f" % o["drop_in_api_literal_batch1024"]["ms_per_step"],
    "PCL": "%.3f" % o["img2pcl_batch1024"]["ms"], "I1": "%.1f" % o["I1_intersection_volume_batch256"]["ms"],
    "MFWD": "%.3f" % r["stage_ms"]["mano_forward(3 kernels)"], "MBWD": "%.3f" % r["stage_ms"]["mano_backward(3 kernels)"],
}
def line(path):
    return json.loads([l for l in open(os.path.join(ROOT, path)) if l.startswith("{")][0])
# the multi-GPU lines may come from an earlier build of the round than the N = 1 line: each names its own N = 1 base
BASE = {"n2": 3.567e6, "n8": 3.567e6}
try:
    BASE.update(json.load(open(os.path.join(ROOT, "profiles/raw/r02_scale_bases.json"))))
except FileNotFoundError:
    pass
n2, n8 = line("profiles/raw/r02_bench_n2.json"), line("profiles/raw/r02_bench_n8_strong.json")
n4 = line("profiles/raw/r02_bench_n4_strong.json")
BASE.setdefault("n4", BASE["n8"])
sub.update({"N4MS": "%.3f" % n4["ms_per_step"], "N4V": M(n4["value"]), "N4EFF": "%.2f" % (n4["value"] / (4 * BASE["n4"])),
            "N4E2E": M(n4["e2e"]["value"]), "N2E2E": M(n2["e2e"]["value"])})
sub.update({"N2MS": "%.3f" % n2["ms_per_step"], "N2V": M(n2["value"]), "N2EFF": "%.2f" % (n2["value"] / (2 * BASE["n2"])),
            "N2BASE": M(BASE["n2"]), "N8MS": "%.3f" % n8["ms_per_step"], "N8V": M(n8["value"]),
            "N8EFF": "%.2f" % (n8["value"] / (8 * BASE["n8"])), "N8BASE": M(BASE["n8"]),
            "N8WMS": "%.3f" % n8["weak_scaling"]["ms_per_step"], "N8WV": M(n8["weak_scaling"]["value"]),
            "N8E2E": M(n8["e2e"]["value"])})
rows = [x for x in csv.reader(open(os.path.join(ROOT, "profiles/raw/r02_launches_B4096.csv"))) if len(x) > 5]
h = rows[0]
t = collections.OrderedDict()
for x in rows[1:]:
    try:
        t.setdefault(x[h.index("Kernel Name")], []).append(float(x[-1]))
    except ValueError:
        pass
step = {k: v for k, v in t.items() if any(s in k for s in ("mano_", "tf32x3", "raster_fwd", "sum_totals"))}
tot = sum(2 * sum(v) / len(v) if "sum_totals" not in k else sum(v) / len(v) for k, v in step.items())
lines = ["| kernel | launches in the capture | mean µs per 2048-hand slice | share of the step |", "|---|---|---|---|"]
for k, v in sorted(step.items(), key=lambda kv: -sum(kv[1]) / len(kv[1])):
    m = sum(v) / len(v) / 1e3
    per_step = m * (1 if "sum_totals" in k else 2)
    lines.append("| `%s` | %d | %.1f | %.1f %% |" % (k.split("(")[0].replace("void ", ""), len(v), m, 100 * per_step * 1e3 / tot))
sub["LAUNCHTABLE"] = "\n".join(lines) + ("\n\n(the `at::` kernels, `view_setup_kernel` and `target_from_u16*` in the CSV belong to input set-up and to "
                                          "the end-to-end hand-off comparison, not to the fused step)")
san = os.path.join(ROOT, "profiles/raw/r02_sanitizer.txt")
sub["SANITIZER"] = ("## 6. compute-sanitizer\n\n`tools/sanitize_smoke.py` under memcheck and racecheck with this build: see `raw/r02_sanitizer.txt`."
                    if os.path.exists(san) else "")
s = open(os.path.join(ROOT, "tools/profiles_readme.tmpl.md")).read()
for k, v in sub.items():
    s = s.replace("@%s@" % k, v)
s = s.replace("(round-1 text, unchanged)\n", "")
assert "@" not in s.split("# §R1")[0].replace("@VALUE", ""), [w for w in s.split() if w.startswith("@")][:5]
open(os.path.join(ROOT, "profiles/README.md"), "w").write(s)
print("profiles/README.md written")
