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

# generateBatchedProof prints `= { 1, 2, ... }`, generateStage3Proof (src/zkvm/spartan/stage3_prover.zig:430-560) `= { { 1, 2, ... } }`
pat = re.compile(r"\[ZOLT\] (STAGE\d+)_ROUND_(\d+): (current_claim|c0|c2|c3|challenge|next_claim|p0|p1|shift_p0|shift_p1) = (?:\{ )+([0-9, ]+?),? (?:\} ?)+$")
rounds = {}
for line in (REF / "logs" / "zolt.log").read_text(errors="replace").splitlines():
    m = pat.match(line)
    if not m:
        continue
    stage, rnd, key, body = m.groups()
    vals = [int(x) for x in body.replace(" ", "").split(",") if x]
    if len(vals) != 32:
        continue
    rounds.setdefault((stage, int(rnd)), {}).setdefault(key, vals)  # the first print of a key in a round

kats = []
for (stage, rnd), d in sorted(rounds.items()):
    if all(k in d for k in ("current_claim", "c0", "c2", "c3", "challenge", "next_claim")):
        kats.append({"stage": stage, "round": rnd, **d})
# the Stage 1 challenges (9 rounds): r_outer of Stage 3 is rounds 8..1 of them (SHIFT_INIT prints r_outer[0] and r_outer[last])
stage1_challenges = [rounds[("STAGE1", i)]["challenge"] for i in range(9) if ("STAGE1", i) in rounds and "challenge" in rounds[("STAGE1", i)]]
# Stage 3 preamble (stage3_prover.zig:150-215): the three input claims and the Shift gamma powers the run printed
pre_pat = re.compile(r"\[ZOLT\] (?:STAGE3_PRE: input_claim\[(\d)\] \(\w+\)|STAGE3_SHIFT: gamma_powers\[(\d)\]) = (?:\{ )+([0-9, ]+?),? (?:\} ?)+$")
stage3_pre = {}
for line in (REF / "logs" / "zolt.log").read_text(errors="replace").splitlines():
    m = pre_pat.match(line)
    if m:
        vals = [int(x) for x in m.group(3).replace(" ", "").split(",") if x]
        if len(vals) == 32:
            stage3_pre.setdefault(f"input_claim_{m.group(1)}" if m.group(1) else f"gamma_power_{m.group(2)}", vals)
# LassoProver's own run in the same log (src/zkvm/lasso/prover.zig:232-456 prints): per round the claim, the coefficients, the
# challenge and the claim after receiveChallenge, as hex of limbs[3..0] -- MONTGOMERY limbs, kept as printed
lasso, cur = [], None
for line in (REF / "logs" / "zolt.log").read_text(errors="replace").splitlines():
    if not line.startswith("[LASSO PROVER]"):
        continue
    m = re.match(r"\[LASSO PROVER\] computeRoundPolynomial round=(\d+) phase=(\w+)", line)
    if m:
        cur = {"round": int(m.group(1)), "phase": m.group(2)}
        lasso.append(cur)
        continue
    if cur is None:
        continue
    m = re.match(r"\[LASSO PROVER\]\s+(current_claim|c\[0\]|c\[1\]|c\[2\]|challenge|current_claim \(after\)) = ([0-9a-f]{64})$", line)
    if m:
        key = {"current_claim": "claim", "c[0]": "c0", "c[1]": "c1", "c[2]": "c2", "challenge": "challenge", "current_claim (after)": "claim_after"}[m.group(1)]
        cur.setdefault(key, m.group(2))
lasso = [r for r in lasso if all(k in r for k in ("claim", "c0", "c1", "c2", "challenge", "claim_after"))]
# Stage4GruenProver's split-eq tables in the same run (src/zkvm/spartan/gruen_eq.zig:19-112 through the STAGE4_GRUEN_INIT prints):
# the first four entries of E_out / E_in (LE bytes); the r_cycle_be lines print the 128-bit challenge encodings, not the field
# elements -- those are the STAGE3_ROUND_k challenges above, reversed
stage4 = {"r_cycle_be": {}, "E_out": {}, "E_in": {}}
for line in (REF / "logs" / "zolt.log").read_text(errors="replace").splitlines():
    m = re.match(r"\[STAGE4_GRUEN_INIT\]\s+r_cycle_be\[(\d+)\] = \{ ([0-9a-f ]+) \}", line)
    if m:
        stage4["r_cycle_be"].setdefault(int(m.group(1)), int(m.group(2).replace(" ", ""), 16))
    m = re.match(r"\[STAGE4_GRUEN_INIT\]\s+(E_out|E_in)\[(\d+)\] = \{ ([0-9, ]+) \}", line)
    if m:
        vals = [int(x) for x in m.group(3).replace(" ", "").split(",") if x]
        if len(vals) == 32:
            stage4[m.group(1)].setdefault(int(m.group(2)), vals)
# ... and its phase-1 rounds: the two weighted sums, the claim, the cubic gruenPolyDeg3 returned and the challenge bound next
# (stage4_gruen_prover.zig:561-757, gruen_eq.zig:120-160, :214-262)
p1_rounds, blk = [], None
for line in (REF / "logs" / "zolt.log").read_text(errors="replace").splitlines():
    m = re.match(r"\[STAGE4\] Round (\d+) -> Phase (\d)", line)
    if m:
        blk = {"round": int(m.group(1))} if m.group(2) == "1" else None
        if blk is not None:
            p1_rounds.append(blk)
        continue
    if blk is None:
        continue
    m = re.match(r"\[ZOLT PHASE1 (q_0|q_X2|previous_claim)\] = \{ ([0-9, ]+) \}", line) or re.match(r"\[ZOLT PHASE1 COEFFS\] (c[0-3])=\{ ([0-9, ]+) \}", line)
    if not m:
        m = re.match(r"\[ZOLT STAGE4\] Round \d+: (challenge) \(LE\) = \{ ([0-9, ]+) \}", line)
    if m:
        vals = [int(x) for x in m.group(2).replace(" ", "").split(",") if x]
        if len(vals) == 32:
            blk.setdefault(m.group(1), vals)
p1_rounds = [b for b in p1_rounds if all(k in b for k in ("q_0", "q_X2", "previous_claim", "c0", "c1", "c2", "c3"))]
stage4 = {"r_cycle_be": [str(stage4["r_cycle_be"][i]) for i in sorted(stage4["r_cycle_be"])], "phase1_rounds": p1_rounds,
          "E_out": [stage4["E_out"][i] for i in sorted(stage4["E_out"])], "E_in": [stage4["E_in"][i] for i in sorted(stage4["E_in"])]}
(OUT / "zolt_log_kats.json").write_text(json.dumps({"source": "logs/zolt.log", "format": "canonical LE bytes (toBytes)", "kats": kats,
                                                   "stage3_pre": stage3_pre, "stage1_challenges": stage1_challenges,
                                                   "stage4_gruen_init": stage4,
                                                   "lasso_rounds": lasso, "lasso_format": "hex of Montgomery limbs[3..0]; log_K = 16, log_T = 8"}))
print(f"{len(lasso)} Lasso round tuples")
print(f"{len(kats)} claim-update KATs")
