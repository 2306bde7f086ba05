"""Dev tool: static SASS of one kernel of an object file, cut into loop nests (a loop = the span of a backward branch), with the
opcode mix and the issue cost (2 cycles per FP64 instruction, 1 otherwise) of every span.  Works without a GPU.

    python tools/sass_loops.py <object.o> <substring of the demangled kernel name> [--dump]
"""
import collections
import re
import subprocess
import sys

FP64 = {"DFMA", "DMUL", "DADD", "DSETP", "DMNMX"}


def kernels(obj):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    out, name, cur = {}, None, []
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                out[name] = cur
            name, cur = m.group(1), []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and name:
            cur.append((int(m.group(1), 16), m.group(2).strip()))
    if name:
        out[name] = cur
    return out


def opcode(ins):
    toks = ins.split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    if op.startswith("IMAD.MOV"):
        return "IMAD.MOV"
    return op.split(".")[0]


def cost(c):
    return sum(v * (2 if k in FP64 else 1) for k, v in c.items())


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    ks = kernels(obj)
    names = {k: subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip() for k in ks}
    sel = [k for k in ks if pat in names[k]]
    for k in sel:
        ins = ks[k]
        print("==", names[k][:160], f"({len(ins)} instructions)")
        addr_ix = {a: i for i, (a, _) in enumerate(ins)}
        loops = []
        for i, (a, s) in enumerate(ins):
            m = re.search(r"\bBRA(?:\.\S+)?\s+(?:!?U?P\d,\s*)?(?:`\(\S+\)|0x([0-9a-f]+))", s)
            if m and m.group(1):
                t = int(m.group(1), 16)
                if t <= a and t in addr_ix:
                    loops.append((addr_ix[t], i))
        loops.sort(key=lambda l: (l[0], -l[1]))
        for lo, hi in loops:
            depth = sum(1 for l2, h2 in loops if l2 <= lo and hi <= h2 and (l2, h2) != (lo, hi))
            inner = [(l2, h2) for l2, h2 in loops if lo <= l2 and h2 <= hi and (l2, h2) != (lo, hi)]
            own = [j for j in range(lo, hi + 1) if not any(l2 <= j <= h2 for l2, h2 in inner)]
            c = collections.Counter(opcode(ins[j][1]) for j in own)
            f64 = sum(v for kk, v in c.items() if kk in FP64)
            print(f"{'  ' * depth}loop [{lo}:{hi}] own instr {len(own)} (fp64 {f64}, issue cost {cost(c)}): " +
                  ", ".join(f"{kk}:{v}" for kk, v in c.most_common(14)))
        c = collections.Counter(opcode(s) for _, s in ins)
        print("whole kernel:", ", ".join(f"{kk}:{v}" for kk, v in c.most_common(20)))
        if "--dump" in sys.argv:
            for i, (a, s) in enumerate(ins):
                print(f"{i:5d} {a:06x} {s}")


if __name__ == "__main__":
    main()
