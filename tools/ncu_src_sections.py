"""Per-kernel opcode / stall histogram of an `ncu --page source --csv` export holding several kernels.

    python tools/ncu_src_sections.py file.src.csv[.gz] [section_index ...]
Sections are the "Kernel Name" blocks in file order (ncu prints every kernel twice; duplicates are skipped).
"""
import csv
import gzip
import io
import sys
from collections import defaultdict


def sections(path):
    data = gzip.open(path, "rt").read() if path.endswith(".gz") else open(path).read()
    rows = list(csv.reader(io.StringIO(data)))
    out, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            out.append(cur)
        elif cur is not None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] and len(r) >= len(cur["hdr"]) - 1:
            cur["rows"].append(r)
    seen, uniq = set(), []
    for s in out:
        key = (s["name"], len(s["rows"]), s["rows"][0][0] if s["rows"] else "")
        if key in seen:
            continue
        seen.add(key)
        uniq.append(s)
    return uniq


def hist(sec, top=24):
    col = {n: i for i, n in enumerate(sec["hdr"])}
    stall_cols = [n for n in sec["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
    by_op = defaultdict(lambda: [0, 0, defaultdict(int)])
    tot_i = tot_s = 0
    for r in sec["rows"]:
        toks = r[col["Source"]].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        op = op.rstrip(";")
        try:
            inst = int(r[col["Instructions Executed"]] or 0)
            samp = int(r[col["# Samples"]] or 0)
        except ValueError:
            continue
        e = by_op[op]
        e[0] += inst
        e[1] += samp
        for s in stall_cols:
            try:
                e[2][s] += int(r[col[s]] or 0)
            except ValueError:
                pass
        tot_i += inst
        tot_s += samp
    print(f"## {sec['name'][:100]}")
    print(f"SASS lines {len(sec['rows'])} (~{len(sec['rows']) * 16 // 1024} KB), warp-instructions {tot_i}, samples {tot_s}")
    allst = defaultdict(int)
    for e in by_op.values():
        for k, v in e[2].items():
            allst[k] += v
    print("stall totals: " + ", ".join(f"{k[6:]}={100 * v / max(1, tot_s):.1f}%" for k, v in sorted(allst.items(), key=lambda kv: -kv[1])[:9]))
    print(f"{'opcode':26s} {'inst':>12s} {'%inst':>6s} {'samples':>9s} {'%samp':>6s}  top stalls")
    for op, (inst, samp, st) in sorted(by_op.items(), key=lambda kv: -kv[1][0])[:top]:
        t3 = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        print(f"{op:26s} {inst:12d} {100 * inst / max(1, tot_i):6.1f} {samp:9d} {100 * samp / max(1, tot_s):6.1f}  "
              + ", ".join(f"{k[6:]}={v}" for k, v in t3 if v))
    print()


if __name__ == "__main__":
    secs = sections(sys.argv[1])
    want = [int(x) for x in sys.argv[2:]] or range(len(secs))
    for i in want:
        hist(secs[i])
