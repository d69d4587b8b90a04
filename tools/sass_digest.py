"""SASS instruction digest of the in-tree libudh.so -> profiles/<name>.md (tracked evidence that the kernels are Blackwell
native: UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA, LDTM = tcgen05.ld; no HMMA / HGMMA).
usage: python tools/sass_digest.py [profiles/r2_sass_digest.md]"""
import collections, hashlib, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "unsuperviseddeephomographyral2018_b200", "libudh.so")
out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_digest.md")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
fn, per, tot = None, collections.defaultdict(collections.Counter), collections.Counter()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        fn = m.group(1); continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*)", ln)
    if m and fn:
        per[fn][m.group(1)] += 1; tot[m.group(1)] += 1
keys = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "SYNCS", "HMMA", "HGMMA", "FFMA", "REDG", "ELECT", "LDGMC"]
L = ["# SASS instruction digest of the in-tree libudh.so", "",
     "Command: `python tools/sass_digest.py` (= `cuobjdump -sass unsuperviseddeephomographyral2018_b200/libudh.so`, the library",
     "`__graft_entry__.build()` produces with `nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`); md5 of the binary at digest",
     "time: `%s`." % hashlib.md5(open(so, "rb").read()).hexdigest(), "",
     "Blackwell-native evidence (B200_PROFILING.md): `UTCHMMA` = tcgen05.mma, `UTMALDG` / `UTMASTG` = TMA load / store, `LDTM` = tcgen05.ld,",
     "`UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier ops, `LDGMC` = multimem.ld_reduce (NVSwitch multicast load-reduce, csrc/dp_update.cu).  No `HMMA` (legacy mma.sync) and no `HGMMA` (Hopper wgmma) anywhere.  The product",
     "library carries no probe kernels (those are in libudh_probe.so).", "",
     "| mnemonic | whole library |", "|---|---|"]
L += ["| %s | %d |" % (k, tot.get(k, 0)) for k in keys]
L += ["", "Per kernel (kernels that issue tcgen05.mma):", "", "| kernel | UTCHMMA | UTMALDG | UTMASTG | LDTM | UTCBAR |", "|---|---|---|---|---|---|"]
dem = subprocess.run(["c++filt"], input="\n".join(per.keys()), capture_output=True, text=True).stdout.splitlines()
for (f, c), d in sorted(zip(per.items(), dem), key=lambda t: -t[0][1]["UTCHMMA"]):
    if c["UTCHMMA"]:
        name = re.sub(r"\(.*", "", d).replace("void ", "").replace("udh::tc::", "").replace("udh::", "")
        L.append("| `%s` | %d | %d | %d | %d | %d |" % (name[:100], c["UTCHMMA"], c["UTMALDG"], c["UTMASTG"], c["LDTM"], c["UTCBAR"]))
open(out_path, "w").write("\n".join(L) + "\n")
print(out_path, "UTCHMMA", tot["UTCHMMA"])
