"""Print registers / spills of every kernel in a ptxas -v log (zolt_b200/lib/obj/*.ptxas.log)."""
import re
import subprocess
import sys

log = open(sys.argv[1]).read()
pat = sys.argv[2] if len(sys.argv) > 2 else ""
for m in re.finditer(r"Compiling entry function '(\S+)'.*?\n.*?Function properties.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes smem)?", log):
    d = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
    if pat and pat not in d:
        continue
    short = re.sub(r"\(.*", "", d).replace("void (anonymous namespace)::", "")
    print(f"{short:50s} regs {m.group(5):>4s} stack {m.group(2):>5s} spill st/ld {m.group(3)}/{m.group(4)} smem {m.group(6)}")
