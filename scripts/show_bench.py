"""Prints the headline numbers of a bench.py output (the JSON line may be preceded by library banners)."""
import json
import sys

for path in sys.argv[1:]:
    line = [l for l in open(path) if l.startswith("{")][-1]
    d = json.loads(line)

    def show(n, r):
        if "failed" in r:
            print("  ", n, r)
            return
        print("   %-4s %8.3f ms/step %7.0f fps  value %.3g | e2e %.3f ms (h2d %d B) e2e16 %.3f ms | %s frac %.3f (%.4f ms/launch) | launches/step %.0f"
              % (n, r["ms_per_step"], r["frames_per_s"], r["value"], r["e2e"]["ms_per_step"], r["e2e"]["h2d_bytes_per_step"], r["e2e_packed16"]["ms_per_step"],
                 r["roofline"]["kernel"], r["roofline"]["frac"], r["roofline"]["ms_per_launch"], r["gpu_launches"] / r["steps"]))
    print(path, "n_gpus", d["n_gpus"], "clocks", d.get("clocks"))
    show(d["config"]["workload"][:3].strip(":"), d)
    for k, v in d.get("workloads", {}).items():
        show(k, v)
    if d.get("cpu_baseline"):
        c = d["cpu_baseline"]
        print("   cpu_baseline %.3g evals/s on %s cores (%s)" % (c["value"] or 0, c["cores"], c["kind"]))
