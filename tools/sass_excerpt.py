"""SASS evidence per shipped kernel: for every kernel in libtfhe_b200.so the counts of the mnemonics
that prove the Blackwell path (UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP = 1-D TMA bulk
copy, SYNCS = mbarrier, USETMAXREG = setmaxnreg, STAS = st.async into a cluster peer's shared memory, UCGABAR = cluster
barrier, DFMA/DADD/DMUL) plus the first lines around the first tcgen05 / TMA / DSMEM instruction.  usage: python tools/sass_excerpt.py > profiles/r2_sass_excerpt.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "rs_tfhe_b200", "csrc", "libtfhe_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCIMMA", "UTCHMMA", "LDTM", "STTM", "UBLKCP", "SYNCS", "USETMAXREG", "STAS", "UCGABAR_ARV", "DFMA", "DADD", "DMUL", "IMAD.MOV",
        "LDS", "STS", "SHFL", "BAR", "LDL", "STL"]
funcs = re.split(r"\n\s*Function : ", sass)[1:]
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}   ({len(funcs)} kernels)\n")
for f in funcs:
    name = f.split("\n")[0]
    demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    lines = [l for l in f.split("\n") if re.match(r"\s*/\*[0-9a-f]{4}\*/", l)]
    c = collections.Counter()
    for l in lines:
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if not m:
            continue
        op = m.group(1)
        for k in KEYS:
            if op == k or op.startswith(k + ".") or (k == "IMAD.MOV" and op.startswith("IMAD.MOV")):
                c[k] += 1
    print(f"## {demangled[:150]}\n   instructions {len(lines)}   " + "  ".join(f"{k}={c[k]}" for k in KEYS if c[k]))
    for pat in ("UTCIMMA", "LDTM", "STTM", "UBLKCP", "STAS"):
        hit = [i for i, l in enumerate(lines) if pat in l]
        if hit:
            i = hit[0]
            for l in lines[max(0, i - 1):i + 2]:
                print("      " + l.strip()[:140])
    print()
