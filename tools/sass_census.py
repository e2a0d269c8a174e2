"""SASS census of libclstm.so: per kernel, how many tcgen05 / TMEM / TMA instructions the compiled code holds
(B200_PROFILING.md 'What proves a Blackwell-native kernel').  Runs on the build box (cuobjdump, no GPU).
Usage: python tools/sass_census.py [path/to/libclstm.so] > profiles/r2_sass_census.txt"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "satflow_b200", "libclstm.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "SYNCS", "HMMA", "MUFU", "HMNMX2"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, per = None, collections.OrderedDict()
for ln in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        fn = re.sub(r"\(.*", "", fn)
        per[fn] = collections.Counter()
        continue
    if fn is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m:
        op = m.group(1)
        per[fn]["_total"] += 1
        for k in MNEMONICS:
            if op.startswith(k):
                per[fn][k] += 1
                if k == "UTCHMMA" and ".2CTA" in op:
                    per[fn]["UTCHMMA.2CTA"] += 1
links = subprocess.run(["ldd", lib], capture_output=True, text=True).stdout
print(f"# {os.path.relpath(lib, ROOT)}: {os.path.getsize(lib)} bytes; links: " +
      ", ".join(sorted({l.split()[0] for l in links.splitlines() if l.strip()})))
tot = collections.Counter()
cols = MNEMONICS + ["UTCHMMA.2CTA"]
print(f"{'instructions':>12} " + " ".join(f"{c:>8}" for c in cols) + "  kernel")
for fn, c in per.items():
    tot.update(c)
    if any(c[k] for k in cols):
        print(f"{c['_total']:>12} " + " ".join(f"{c[k]:>8}" for k in cols) + f"  {fn[:110]}")
print(f"{tot['_total']:>12} " + " ".join(f"{tot[k]:>8}" for k in cols) + "  TOTAL")
