"""SASS evidence for the lookup kernel: per instantiation of weight_lists_kernel the static counts of the instructions the
design relies on (bulk copy + mbarrier staging, 256-bit record load, packed fp32 arithmetic) and of local-memory
instructions, with the first occurrences quoted:  python scripts/sass_report.py > profiles/r02_sass_weight_lists.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pcl_tracking_b200", "lib", "libpft.so")
syms = subprocess.run(["cuobjdump", "-elf", LIB], capture_output=True, text=True).stdout
names = sorted(n for n in set(re.findall(r"(_ZN3pft19weight_lists_kernel\w+)", syms)) if "_param_" not in n)
WHAT = [("UBLKCP", "cp.async.bulk global -> shared (staging of the indexed points)"), ("SYNCS", "mbarrier init / expect_tx / try_wait"),
        ("LDG.E.ENL2.256", "one 256-bit load per octant record / pool group"), ("FADD2", "packed fp32 subtract (dx, dy)"), ("FMUL2", "packed fp32 multiply (dx*dx, dy*dy)"),
        ("LDS.128", "candidate point from the staged array"), ("LDL", "local-memory load"), ("STL", "local-memory store"), ("FFMA", "fused multiply-add: only inside the compiler's own sqrt.rn / reciprocal refinement sequences (-fmad=false: no user add/multiply is contracted)")]
print("libpft.so, sm_100a -- weight_lists_kernel<USE_HSV, THREADS, DYN>: static SASS counts (cuobjdump -sass)\n")
for n in names:
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", n, LIB], capture_output=True, text=True).stdout
    ins = [l for l in sass.splitlines() if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", l)]
    demangled = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    print("== %s   (%d instructions)" % (demangled, len(ins)))
    # the list path comes first in the listing; the row-table body (taken only when the lists are off for a crop) is the
    # tail that starts where the kernel copies its parameter block to the stack for the non-inlined call
    first_stl64 = next((i for i, l in enumerate(ins) if "STL.64" in l), len(ins))
    for key, why in WHAT:
        hits = [i for i, l in enumerate(ins) if re.search(r"\b%s\b" % re.escape(key), l)]
        in_list_path = [i for i in hits if i < first_stl64]
        print("   %-16s %5d  (before the first STL.64: %d)   %s" % (key, len(hits), len(in_list_path), why))
        for i in hits[:2]:
            print("        " + ins[i].split("/*")[1].split("*/")[0] + ": " + ins[i].split("*/")[1].split(";")[0].strip())
    print()
print("Reading: the listing of a kernel holds the list path first and then the row-table body (taken only when the index header says the\n"
      "lists are off for a crop); that body is a non-inlined call whose parameter block is copied to the stack with a run of STL.64 --\n"
      "the position of the first STL.64 splits the static counts.  The split is a heuristic (a spill before it moves the line); what\n"
      "actually runs is counted by ncu:\n")
print("Dynamic counts (ncu --set full, profiles/r02_kernel_metrics_c2.json / _c4.json: sass__inst_executed_local_loads / _stores):\n"
      "   weight_lists_kernel<1, 640, 1> on c2: 0 / 0 of 18.7 M warp instructions;  weight_lists_kernel<1, 896, 1> on c4: 104 144 / 4 144 of 1.59 G.")
