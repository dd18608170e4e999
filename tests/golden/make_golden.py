"""Regenerate tests/golden/*.json from the reference tree (run in the build container only; the GPU
box has no /root/reference and never needs to -- the JSON files are committed).

  zolt_log_kats.json : every `[ZOLT] STAGEk_ROUND_j: current_claim/c0/c2/c3/challenge/next_claim` tuple
                       printed by generateBatchedProof (src/zkvm/batched_sumcheck.zig:326-413) in the
                       captured run logs/zolt.log, as canonical little-endian byte lists.
"""
import json
import re
import sys
from pathlib import Path

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
OUT = Path(__file__).resolve().parent

pat = re.compile(r"\[ZOLT\] (STAGE\d+)_ROUND_(\d+): (current_claim|c0|c2|c3|challenge|next_claim) = \{+ ([0-9, ]+?),? \}+")
rounds = {}
for line in (REF / "logs" / "zolt.log").read_text(errors="replace").splitlines():
    m = pat.match(line)
    if not m:
        continue
    stage, rnd, key, body = m.groups()
    vals = [int(x) for x in body.replace(" ", "").split(",") if x]
    if len(vals) != 32:
        continue
    rounds.setdefault((stage, int(rnd)), {})[key] = vals

kats = []
for (stage, rnd), d in sorted(rounds.items()):
    if all(k in d for k in ("current_claim", "c0", "c2", "c3", "challenge", "next_claim")):
        kats.append({"stage": stage, "round": rnd, **d})
(OUT / "zolt_log_kats.json").write_text(json.dumps({"source": "logs/zolt.log", "format": "canonical LE bytes (toBytes)", "kats": kats}))
print(f"{len(kats)} claim-update KATs")
