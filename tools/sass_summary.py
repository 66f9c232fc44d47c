"""Per-kernel counts of the Blackwell tensor-core / TMA / TMEM opcodes in libvidseg_b200.so (cuobjdump -sass):
UTCHMMA (tcgen05.mma kind::f16), UTCQMMA (kind::f8f6f4), UTCIMMA (kind::i8), UTMALDG (TMA tile loads), LDTM / STTM
(tcgen05.ld / st), UTCBAR (tcgen05.commit), SYNCS (mbarrier).  `python tools/sass_summary.py > profiles/r02_sass_summary.txt`"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "vidseg_diffusion_b200", "libvidseg_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
ops = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTMALDG", "LDTM", "STTM", "UTCBAR", "SYNCS", "HMMA", "IMMA"]
cur, counts = None, collections.OrderedDict()
arch = set(re.findall(r"arch = (sm_\w+)", out))
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts.setdefault(cur, collections.Counter())
        continue
    if cur:
        for op in ops:
            if re.search(r"\b" + op + r"\b|\b" + op + r"\.", line):
                counts[cur][op] += 1
                break
        if re.search(r"\b(UTCHMMA|UTCQMMA|UTMALDG)\S*\.2CTA", line):
            counts[cur]["2CTA"] += 1
        counts[cur]["_instr"] += bool(re.search(r"/\*[0-9a-f]{4,6}\*/", line))
print(f"# cuobjdump -sass {os.path.relpath(so, ROOT)}  (architectures in the fatbin: {sorted(arch)})")
print("# 2CTA = UTCHMMA / UTCQMMA / UTMALDG carrying the .2CTA modifier (tcgen05.mma.cta_group::2 and the pair's TMA loads)")
print(f"# {'kernel':70s} " + " ".join(f"{o:>8s}" for o in ops) + "     2CTA   instr")
for k, c in counts.items():
    if any(c[o] for o in ops[:6]):
        print(f"{k[:72]:72s} " + " ".join(f"{c[o]:8d}" for o in ops) + f" {c['2CTA']:8d} {c['_instr']:7d}")
