"""SASS evidence for the tensor-core kernels: mnemonic histogram + every tcgen05 / TMA / TMEM instruction line, per kernel.
    python tools/sass_summary.py            -> profiles/r2_sass_<kernel>.txt  (full listings gzip'ed beside them)"""
import collections, gzip, os, re, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(REPO, "sfd2_b200", "libsfd2_b200.so")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
txt = subprocess.run(["cuobjdump", "-sass", SO], stdout=subprocess.PIPE, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
KEY = re.compile(r"\b(UTC\w+|UTMA\w+|LDTM\w*|STTM\w*|UTCBAR\w*|SYNCS\w*|REDG\w*|REDUX\w*|CREDUX\w*|ATOMG\w*|UBLKCP\w*|UTMACMDFLUSH|ACQBULK|ERRBAR|CGAERRBAR|UCGABAR\w*)")
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    short = next((k for k in ("tc_conv_kernel", "tc_match_kernel", "tc_desc_sparse_kernel", "conv1a_mma_kernel", "match_prep_kernel", "preprocess_kernel") if k in name), None)
    if not short:
        continue
    variant = "_sub2" if "ILi2E" in name else ("_sub1" if "ILi1E" in name else ("_pair_slim" if "ILb1ELb1E" in name else ("_pair" if "ILb1ELb0E" in name else "")))
    lines = [l for l in f.split("\n") if re.search(r"/\*[0-9a-f]{4,}\*/", l)]
    hist = collections.Counter()
    keep = []
    for l in lines:
        m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", l)
        if m:
            hist[m.group(1).split(".")[0]] += 1
            if KEY.search(m.group(1)):
                keep.append(re.sub(r"\s+", " ", l.split("*/", 1)[1].split("/*")[0]).strip())
    out = os.path.join(REPO, "profiles", f"{tag}_sass_{short}{variant}.txt")
    with open(out, "w") as o:
        o.write(f"# {name}\n# cuobjdump -sass sfd2_b200/libsfd2_b200.so (sm_100a); {len(lines)} instructions\n\n## mnemonic histogram\n")
        for k, v in hist.most_common():
            o.write(f"{v:6d}  {k}\n")
        o.write("\n## tensor-core / TMA / TMEM / async-barrier / reduction instructions (in program order, operands elided where repeated)\n")
        cnt = collections.Counter(k.split(" ")[0] for k in keep)
        for k, v in cnt.most_common():
            o.write(f"{v:6d}  {k}\n")
        o.write("\n## first 60 such lines\n" + "\n".join(keep[:60]) + "\n")
    with gzip.open(out.replace(".txt", "_full.txt.gz"), "wt") as g:
        g.write("Function : " + f)
    print(out, len(lines))
