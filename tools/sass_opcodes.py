"""Per-kernel SASS opcode counts of libd3fk.so (cuobjdump -sass; runs without a GPU): the evidence that the tensor-core
kernels are tcgen05 / TMA code (UTCHMMA, LDTM, UTMALDG, UTCBAR), that MMA / TMA issue sits in elect.sync regions (ELECT, no
R2UR.BROADCAST waterfall), which kernels still gather with LDGSTS, and the warp-level path of the head convolution (HMMA, LDSM).
usage: python tools/sass_opcodes.py [libd3fk.so] > profiles/sass_opcodes_r02.txt"""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                          "denoising_diffusion_deep_fake_b200", "libd3fk.so")
OPS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTCBAR", "ELECT", "R2UR.BROADCAST", "STAS", "UCGABAR", "LDGSTS", "SHFL", "HMMA", "LDSM",
       "RED.E.ADD.F32", "RED.E.ADD.F32x4"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
counts, instr, cur = collections.defaultdict(collections.Counter), collections.Counter(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Za-z0-9_.]*)", line)
    if not (cur and m):
        continue
    op = m.group(1)
    instr[cur] += 1
    for o in OPS[:-2]:
        if op == o or op.startswith(o + "."):
            counts[cur][o] += 1
    if op.startswith("REDG.E.ADD.F32"):      # fire-and-forget global reductions: scalar / 16-byte vector form
        counts[cur]["RED.E.ADD.F32x4" if op.startswith("REDG.E.ADD.F32x4") else "RED.E.ADD.F32"] += 1
keep = [k for k in instr if any(counts[k][o] for o in ("UTCHMMA", "UTMALDG", "HMMA", "LDTM"))]
w = 92
print(f"{'kernel (mangled, d3fk:: stripped)':{w}s} " + " ".join(f"{o:>8s}" if len(o) <= 8 else f" {o}" for o in ["instr"] + OPS))
for k in sorted(keep):
    name = k.replace("_ZN4d3fk", "")[:w]
    cells = [instr[k]] + [counts[k][o] for o in OPS]
    print(f"{name:{w}s} " + " ".join(f"{c:>{max(8, len(o) + 1)}d}" for c, o in zip(cells, ["instr"] + OPS)))
