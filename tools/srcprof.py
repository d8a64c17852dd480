"""Correlate an ncu SASS-level source page with CUDA source lines using nvdisasm -gi line info.
usage: tools_srcprof.py <ncu_source.csv> <nvdisasm.sass> <mangled kernel name> [top N]"""
import csv, re, sys
from collections import Counter, defaultdict
csvf, sassf, kname = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 25
rows = list(csv.reader(open(csvf)))
h = rows[1]
ie, si, ss = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
prof = [(r[si].strip(), int(r[ie]), int(r[ss])) for r in rows[2:] if len(r) > ie]
# parse sass: instructions of the kernel with current line + inline chain
insts = []
infn = False
cur = None
chain = []
for line in open(sassf):
    if line.startswith("//-----") and ".text." in line:
        infn = kname in line
        continue
    if not infn: continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)(.*)', line)
    if m:
        inl = "inlined at" in line
        if not inl:
            cur = (m.group(1).split('/')[-1], int(m.group(2))); chain = [cur]
        else:
            chain.append((m.group(1).split('/')[-1], int(m.group(2))))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
    if m:
        insts.append((m.group(2).strip(), list(chain)))
print(len(prof), "profiled instrs;", len(insts), "disassembled instrs")
n = min(len(prof), len(insts))
tot = sum(p[1] for p in prof)
inner = Counter(); outer = Counter(); samp_outer = Counter()
for i in range(n):
    txt, cnt, smp = prof[i]
    ch = insts[i][1]
    if not ch: continue
    inner[ch[0]] += cnt
    # outermost frame in generate.cu kernel body
    o = ch[-1]
    outer[o] += cnt; samp_outer[o] += smp
print("== by innermost source line ==")
for k, v in inner.most_common(topn): print(f"{k[0]}:{k[1]:5d} {v/tot*100:6.2f}%")
print("== by outermost (kernel body) line ==")
for k, v in outer.most_common(topn): print(f"{k[0]}:{k[1]:5d} {v/tot*100:6.2f}%  samples {samp_outer[k]}")
