"""Instruction mix of the unrolled phase-1 loop of k_classify_pair_mini<K,8,2,2> in an object file
(run here, no GPU needed): finds the largest backward branch of the kernel (the 16-window body)
and counts opcodes per window, split by the pipe they issue to.
usage: sass_loop_mix.py <classify.o> [K]"""
import collections
import re
import subprocess
import sys

ALU = {"SHF", "LOP3", "IADD3", "VIMNMX3", "VIMNMX", "ISETP", "SEL", "LEA", "PRMT", "POPC", "FLO", "BREV", "SGXT",
       "PLOP3", "IADD", "VIADD", "MOV", "LOP", "IMNMX", "CS2R", "R2P", "P2R"}
FMA = {"IMAD", "FFMA", "FMUL", "FADD"}


def main(obj, k="21"):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    blocks = sass.split("Function : ")
    body = next(b for b in blocks if f"pair_miniILi{k}ELi8ELi2ELi2ELb0" in b.split("\n")[0])
    ins = []
    for line in body.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(3), line))
    best = None
    for a, op, line in ins:
        if op.startswith("BRA"):
            t = re.search(r"0x([0-9a-f]+)", line.split("BRA")[1])
            if t and int(t.group(1), 16) < a:
                span = a - int(t.group(1), 16)
                if best is None or span > best[0]:
                    best = (span, int(t.group(1), 16), a)
    _, lo, hi = best
    loop = [op.split(".")[0] for a, op, _ in ins if lo <= a <= hi]
    c = collections.Counter(loop)
    alu = sum(v for o, v in c.items() if o in ALU)
    fma = sum(v for o, v in c.items() if o in FMA)
    regs = re.search(r"REG:(\d+)", subprocess.run(["cuobjdump", "-res-usage", obj], capture_output=True, text=True).stdout)
    print(f"k={k}: {len(ins)} instructions in the kernel; loop body {len(loop)} = {len(loop) / 16:.2f} per window: "
          f"ALU pipe {alu / 16:.2f}, FMA pipe {fma / 16:.2f}, loads {c.get('LDG', 0) / 16:.2f}")
    print("  " + ", ".join(f"{o} {v / 16:.2f}" for o, v in c.most_common()))


if __name__ == "__main__":
    main(*sys.argv[1:])
