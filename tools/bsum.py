"""Print a compact summary of the last bench.py JSON line in a file."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d.get("roofline") or {}
print("value %.3fM %s  ms/step %.3f" % (d["value"] / 1e6, d["unit"], d["ms_per_step"]))
print("  kernels ms:", {k: round(v, 4) for k, v in r.get("ms_per_launch", {}).items()}, "frac:", {k: round(v, 4) for k, v in r.get("frac_per_kernel", {}).items()})
if d.get("e2e"): print("  e2e %.3fM" % (d["e2e"]["value"] / 1e6 if d["e2e"]["value"] else 0))
if d.get("cpu_baseline"): print("  cpu %.1f (%d cores, %s)" % (d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"]["kind"]))
print("  clocks", d.get("clocks"), "launches", d.get("gpu_launches"))
for k, v in (d.get("per_op") or {}).items():
    print("  %-8s %.3f ms  frac %.3f  [%s]" % (k, v["ms"], v["frac"], v["kernel"]))
for label, res in (d.get("other_configs") or {}).items():
    print("  ==", label)
    if "ct_pairs_per_s" in res:
        print("     %.3f ms/step  %.3gM ct pairs/s  %.0f GB/s per GPU  frac %.3f  launches/step %d" % (res["ms"], res["ct_pairs_per_s"] / 1e6, res["GB/s_per_gpu"], res["frac"], res["kernel_launches_per_step"]))
        for k, v in res.items():
            if isinstance(v, dict) and "ms" in v: print("     %-18s %.3f ms  frac %.3f" % (k, v["ms"], v["frac"]))
        if "cpu_baseline" in res: print("     cpu", res["cpu_baseline"])
        continue
    for k, v in res.items():
        if isinstance(v, dict): print("     %-18s %.3f ms  %.3gM elems/s  frac %.3f  %s" % (k, v["ms"], v["elems_per_s"] / 1e6, v["frac"], v.get("kernel", "")))
