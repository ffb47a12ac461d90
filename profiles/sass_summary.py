#!/usr/bin/env python
"""Per-kernel summary of an `ncu --set full --import-source on` report: headline metrics from the
details page and, from the SASS source page, warp instructions per warp by opcode with stall shares.
usage: python profiles/sass_summary.py gpurun_out/prof.ncu-rep"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ("Duration", "DRAM Throughput", "Memory Throughput", "L1/TEX Hit Rate", "L2 Hit Rate", "Registers Per Thread",
        "Theoretical Occupancy", "Achieved Occupancy", "Executed Ipc Active", "Issue Slots Busy", "Eligible Warps Per Scheduler",
        "Warp Cycles Per Issued Instruction", "Dynamic Shared Memory Per Block", "Local Memory Spilling", "Compute (SM) Throughput")


def main(rep):
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    for line in det.splitlines():
        t = line.strip()
        if line.startswith("  ") and not line.startswith("    ") and "(" in t and "Context" in t:
            print("\n== " + t[:110])
        elif any(t.startswith(k) for k in KEYS):
            print("   " + " ".join(t.split()))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    kern, cur = [], None
    for r in csv.reader(io.StringIO(src)):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            kern.append(cur)
        elif r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    for k in kern:
        h = k["hdr"]
        ie, isamp, isrc = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
        nw = max(int(r[ie]) for r in k["rows"][:4])
        tot = sum(int(r[ie]) for r in k["rows"])
        print(f"\n== SASS {k['name'][:90]}: {len(k['rows'])} SASS lines, {tot / nw:.0f} warp instructions per warp")
        byop, sop = collections.Counter(), collections.Counter()
        for r in k["rows"]:
            f = r[isrc].split()
            op = (f[1] if f[0].startswith("@") else f[0]).split(".")[0]
            byop[op] += int(r[ie])
            sop[op] += int(r[isamp])
        ts = max(1, sum(sop.values()))
        print("   " + ", ".join(f"{op} {c / nw:.0f} ({100 * sop[op] / ts:.0f}%)" for op, c in byop.most_common(16)))
        cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
        st = collections.Counter()
        for r in k["rows"]:
            for i in cols:
                st[h[i]] += int(r[i] or 0)
        tt = max(1, sum(st.values()))
        print("   stalls: " + ", ".join(f"{n[6:]} {100 * v / tt:.0f}%" for n, v in st.most_common(7)))
        if TOP:
            # the TOP hottest SASS lines (by samples) with every non-zero memory column: which loads cost the traffic
            mem = [i for i, x in enumerate(h) if any(w in x for w in ("Sector", "Bytes", "Access", "Requests"))]
            print("   memory columns: " + "; ".join(h[i] for i in mem))
            for r in sorted(k["rows"], key=lambda r: -int(r[isamp]))[:TOP]:
                extra = ", ".join(f"{h[i]}={r[i]}" for i in mem if r[i] not in ("", "0"))
                print(f"   {r[0][-6:]} samples {r[isamp]:>6} exec {r[ie]:>9}  {r[isrc][:70]}  {extra}")
            # totals of the memory columns over global loads / stores
            for pref in ("LDG", "LD.", "STG", "ST."):
                tot_m = collections.Counter()
                for r in k["rows"]:
                    f = r[isrc].split()
                    op = f[1] if f[0].startswith("@") else f[0]
                    if op.startswith(pref):
                        for i in mem:
                            try:
                                tot_m[h[i]] += float(r[i] or 0)
                            except ValueError:
                                pass
                if tot_m:
                    print(f"   {pref}* totals: " + ", ".join(f"{n}={v:.3g}" for n, v in tot_m.items() if v))


TOP = 0
if __name__ == "__main__":
    if len(sys.argv) > 2:
        TOP = int(sys.argv[2])
    main(sys.argv[1])
