"""Aggregate an `ncu --page source --csv` export by opcode: executed instructions and stall samples.

    ncu -i rep --page source --csv --kernel-name regex:NAME --launch-skip S --launch-count 1 > src.csv
    python tools/ncu_source_hist.py src.csv
"""
import csv
import sys
from collections import defaultdict


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    # the SASS table starts at the header row whose first cell is "Address" and has a "Source" column
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {n: i for i, n in enumerate(hdr)}
    stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    by_op = defaultdict(lambda: [0, 0, defaultdict(int)])
    tot_inst = tot_samp = 0
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr) or not r[0].startswith("0x") and not r[0].isdigit():
            if r and r[0] == "Address":
                break
            continue
        src = r[col["Source"]].strip()
        toks = src.split()
        op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
        try:
            inst = int(r[col["Instructions Executed"]].replace(",", "") or 0)
            samp = int(r[col["# Samples"]].replace(",", "") or 0)
        except ValueError:
            continue
        e = by_op[op]
        e[0] += inst
        e[1] += samp
        for s in stall_cols:
            try:
                e[2][s] += int(r[col[s]].replace(",", "") or 0)
            except ValueError:
                pass
        tot_inst += inst
        tot_samp += samp
    print(f"total warp-instructions {tot_inst}, samples {tot_samp}")
    print(f"{'opcode':28s} {'inst':>12s} {'%inst':>6s} {'samples':>9s} {'%samp':>6s}  top stalls")
    for op, (inst, samp, st) in sorted(by_op.items(), key=lambda kv: -kv[1][0])[:28]:
        top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        print(f"{op:28s} {inst:12d} {100 * inst / max(1, tot_inst):6.1f} {samp:9d} {100 * samp / max(1, tot_samp):6.1f}  "
              + ", ".join(f"{k[6:]}={v}" for k, v in top if v))


if __name__ == "__main__":
    main()
