"""Size and local-memory traffic (spills) of the event loops of the whole-filter kernels, parsed from the built library.

    python tools/hot_loop_check.py [kernel-name-substring ...]
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("SEM_LIB_PATH") or os.path.join(ROOT, "stochastic-epidemic-modelling_b200", "libsem_b200.so")


def loops(body, min_wide):
    ins = {int(m.group(1), 16): m.group(2) for m in re.finditer(r"/\*([0-9a-f]{4,6})\*/\s+([^;]+);", body)}
    out = []
    for a, t in ins.items():
        m = re.search(r"BRA (?:P\d, )?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            loop = [ins[x] for x in range(int(m.group(1), 16), a + 16, 16) if x in ins]
            if sum("IMAD.WIDE" in i for i in loop) >= min_wide:
                out.append(loop)
    return out


def main():
    pats = sys.argv[1:] or ["pf_persistent"]
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    for chunk in sass.split("Function : ")[1:]:
        name = chunk.split("\n", 1)[0].strip()
        if not any(p in name for p in pats):
            continue
        ls = loops(chunk, 15)
        if not ls:
            continue
        best = min(ls, key=len)
        print(f"{name[:90]:90s} loop {len(best):4d} instr, LDL {sum('LDL' in i for i in best)}, STL {sum('STL' in i for i in best)}, "
              f"total LDL/STL in kernel {chunk.count('LDL')}/{chunk.count('STL')}")


if __name__ == "__main__":
    main()
