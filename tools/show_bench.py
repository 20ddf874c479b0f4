"""prints the main numbers of a bench.py JSON line (file argument)"""
import json
import sys

s = open(sys.argv[1]).read()
d = json.loads(s[s.index('{"metric'):].splitlines()[0])
print("N", d["n_gpus"], "value %.1f GDoF/s" % (d["value"] / 1e9), "ms %.4f" % d["ms_per_step"],
      "deferred", d["value_deferred_norm"] and round(d["value_deferred_norm"]["value"] / 1e9, 1), "parity", d["parity"].get("match"),
      "same-grid 1 GPU", d["parity"].get("single_gpu_same_grid", {}).get("value", 0) / 1e9, "e2e", d["e2e"] and round(d["e2e"]["value"] / 1e9, 1))
if d.get("dropin"):
    print("dropin", d["dropin"])
a = d.get("amg")
if a:
    for v in ("multicolour_gs_fine_l1_jacobi_coarse", "l1_jacobi_all_levels"):
        r = a[v]
        print(" ", v, "setup %.2f s" % r["setup_s"], [(x["n"], x["sharded"]) for x in r["levels"]][:5])
        for k, x in r["kernels"].items():
            print("    ", k, round(x["ms"], 4), "frac", round(x["frac"], 3), "launches", x["launches"])
        c = r["cycle"]
        print("    cycle ms %.4f" % c["ms_per_cycle"], "launches", c["launches_per_cycle"], "frac %.3f" % c["frac"], "reduction", round(c["reduction_per_cycle"], 4))
    print("  parity", a["parity"])
