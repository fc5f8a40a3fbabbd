#!/usr/bin/env python
"""Stall samples of a warp-specialised kernel split by WARP ROLE, read from an .ncu-rep (no GPU needed):

    python tools/ncu_roles.py file.ncu-rep [N_top_instructions]

The SASS page of an `ncu --set full --import-source on` capture carries, per instruction, its execution count and
its stall samples.  In a kernel whose roles run different loops the execution count identifies the role (k_sarl_umma,
T = 256, 1024 blocks: 131072 = 8 producer warps x 16 stages x 1024, 65536 = 8 step warps x 8 stages x 1024, ...),
so grouping by it gives the share of time and the stall mix of every role, and the top instructions show who waits
for whom (the branch after an mbarrier try_wait)."""
import collections
import csv
import subprocess
import sys


def main(path, top=25):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, data = rows[1], rows[2:]
    ia, iS, iE = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    total = sum(int(r[iS]) for r in data)
    print(rows[0][1] if len(rows[0]) > 1 else "")
    print(f"stall samples {total}, warp instructions {sum(int(r[iE]) for r in data)}")
    groups = collections.defaultdict(list)
    for r in data:
        groups[int(r[iE])].append(r)
    print("| executions per instruction | static instructions | samples | share | top stall reasons |")
    print("|---|---|---|---|---|")
    for cnt, rs in sorted(groups.items(), key=lambda kv: -sum(int(r[iS]) for r in kv[1]))[:8]:
        c = collections.Counter()
        for r in rs:
            for s in stalls:
                c[s] += int(r[hdr.index(s)])
        t = max(1, sum(c.values()))
        n = sum(int(r[iS]) for r in rs)
        mix = ", ".join(f"{k[6:]} {100 * v / t:.0f} %" for k, v in c.most_common(5))
        print(f"| {cnt} | {len(rs)} | {n} | {100 * n / total:.1f} % | {mix} |")
    print("\ntop instructions by stall samples (samples, executions, SASS):")
    for r in sorted(data, key=lambda r: -int(r[iS]))[:top]:
        print(f"  {r[iS]:>6s} {r[iE]:>9s}  {r[ia].strip()[:96]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
