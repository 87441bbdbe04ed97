"""profiles/r<N>_sass_opcounts.txt: per-kernel counts of the SASS mnemonics that prove the Blackwell path
(UTCHMMA = tcgen05.mma kind::tf32, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA loads / stores, UTCBAR = tcgen05.commit,
SYNCS = mbarrier) from `cuobjdump -sass` of the in-tree library.   python tools/sass_opcounts.py > profiles/r2_sass_opcounts.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "davo_b200", "csrc", "libdavo_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
OPS = ("UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "UTCATOMSWS", "SYNCS", "HMMA", "IMMA", "FFMA")
arch = re.search(r"arch = (sm_\w+)", txt)
print("# %s  (%s, %d bytes)\n# mnemonic counts per kernel; cm:: = channels-on-M conv, pm:: = pixels-on-M conv" % (os.path.relpath(lib, ROOT), arch.group(1) if arch else "?", os.path.getsize(lib)))
total = collections.Counter()
rows = []
for m in re.finditer(r"Function : (\S+)\n(.*?)(?=\n\s*Function : |\Z)", txt, re.S):
    name, body = demangle(m.group(1)), m.group(2)
    c = collections.Counter()
    for line in body.splitlines():
        mm = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if mm:
            op = mm.group(1).split(".")[0]
            if op in OPS:
                c[op] += 1
            if op == "UTMALDG" or op == "UTMASTG":
                c[mm.group(1)] += 0
    total.update(c)
    short = re.sub(r"\(.*", "", name).replace("davo::", "")
    rows.append((short, c))
print("%-72s %s" % ("kernel", " ".join("%8s" % o for o in OPS)))
for short, c in sorted(r for r in rows if r[0]):
    print("%-72s %s" % (short[:72], " ".join("%8d" % c[o] for o in OPS)))
print("%-72s %s" % ("TOTAL", " ".join("%8d" % total[o] for o in OPS)))
