"""Instruction mix of the time-step loop of one integrator instantiation, from the built object.

usage: python tools/sass_loop.py fiveeqscm_b200/csrc/build/ufair_inst_f64_g3.o IdLi3ELi0ELb1ELi1ELj0ELb0E
(the second argument is a substring of the mangled kernel name: Real, NGAS, AMODE, EMEM, GPL, FORM, INV)
Prints the largest innermost backward-branch loop's instruction count by opcode class.
"""
import collections
import re
import subprocess
import sys


def main(obj, key):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    name, ins = None, []
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            continue
        if name and key in name:
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
            if m:
                ins.append((int(m.group(1), 16), m.group(2)))
    loops = []
    for a, t in ins:
        m = re.search(r"BRA.*0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            loops.append((int(m.group(1), 16), a))
    # the step loop: the largest loop that contains no other loop larger than 100 instructions
    size = lambda lo, hi: sum(1 for x, _ in ins if lo <= x <= hi)
    inner = [l for l in loops if not any(o != l and l[0] <= o[0] and o[1] <= l[1] and size(*o) > 100 for o in loops)]
    lo, hi = max(inner, key=lambda l: size(*l))
    c = collections.Counter()
    for x, t in ins:
        if lo <= x <= hi:
            op = t.split()[1] if t.startswith("@") else t.split()[0]
            c[op.split(".")[0]] += 1
    fp64 = sum(v for k, v in c.items() if k in ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX"))
    print(f"loop {lo:#x}..{hi:#x}: {sum(c.values())} instructions, {fp64} FP64")
    print(", ".join(f"{k} {v}" for k, v in c.most_common()))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
