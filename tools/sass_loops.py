"""static loop structure of one kernel instantiation: python tools/sass_loops.py <object> <name-regex>
prints every backward branch (loop) with the size of its body and the opcode mix of the largest bodies"""
import collections, re, subprocess, sys
obj, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
blocks = out.split("Function : ")[1:]
for blk in blocks:
    name = blk.split("\n", 1)[0]
    if not re.search(pat, name):
        continue
    ins = []
    for l in blk.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    print(name[:120], "instructions", len(ins))
    loops = []
    for a, t in ins:
        m = re.search(r"\bBRA(?:\.\w+)*\s+(?:[!\w]+,\s*)?(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) <= a:
            loops.append((int(m.group(1), 16), a))
    for lo, hi in sorted(loops, key=lambda x: x[0] - x[1])[:8]:
        body = [t for a, t in ins if lo <= a <= hi]
        mix = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for t in body)
        print("  loop %05x..%05x  %5d instr  " % (lo, hi, len(body)), dict(mix.most_common(12)))
