"""Map ncu per-SASS-instruction counters (exported `--page source --csv`) onto CUDA source lines with `nvdisasm -g`
of the locally built object (same build).  usage: ncu_lines.py <src.csv> <object.o> <mangled-name-substring> <source.cu>"""
import collections, csv, re, subprocess, sys, os, tempfile
src_csv, obj, sub, cu = sys.argv[1:5]
r = list(csv.reader(open(src_csv)))
start = [i for i, x in enumerate(r) if x and x[0] == 'Address'][-1]
h = r[start]; ie = h.index('Instructions Executed'); sc = h.index('Source'); ws = h.index('Warp Stall Sampling (All Samples)')
counts = []
for x in r[start + 1:]:
    if len(x) <= ie: continue
    try: counts.append((x[sc].strip(), int(x[ie]), int(x[ws] or 0)))
    except ValueError: pass
d = tempfile.mkdtemp()
subprocess.run("cd %s && cuobjdump -xelf all %s > /dev/null && nvdisasm -g -c *.cubin > d.sass" % (d, os.path.abspath(obj)), shell=True, check=True)
lines = open(d + "/d.sass").read().split('\n')
cands = [i for i, l in enumerate(lines) if l.startswith('//--------------------- .text.') and sub in l]
best = None
for s0 in cands:
    cur = None; ins = []
    for l in lines[s0 + 1:]:
        if l.startswith('//--------------------- '): break
        m = re.search(r'//## File "(.*?)", line (\d+)', l)
        if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(.*?);', l)
        if m: ins.append((cur, m.group(1).strip()))
    if len(ins) == len(counts): best = ins; break
if best is None: sys.exit("no function with %d instructions among %d candidates" % (len(counts), len(cands)))
per = collections.Counter(); st = collections.Counter(); ops = collections.Counter()
for (ln, op), (_, c, w) in zip(best, counts):
    per[ln] += c; st[ln] += w
    o = op.split()[0] if not op.startswith('@') else op.split()[1]
    ops[o.split('.')[0]] += c
tot = sum(per.values()); wt = sum(st.values())
srcs = {}
def srcline(key):
    if not key: return ''
    f, ln = key
    if f not in srcs:
        path = os.path.join(os.path.dirname(cu), f)
        srcs[f] = open(path).read().split('\n') if os.path.exists(path) else []
    return srcs[f][ln - 1].strip()[:100] if 0 < ln <= len(srcs[f]) else ''

print("total warp instructions %d, stall samples %d" % (tot, wt))
print("opcodes:", ", ".join("%s %.1f%%" % (o, 100 * c / tot) for o, c in ops.most_common(14)))
for ln, c in per.most_common(int(sys.argv[5]) if len(sys.argv) > 5 else 40):
    print('%5.1f%% inst %5.1f%% stall  %-16s %s' % (100 * c / tot, 100 * st[ln] / max(wt, 1), "%s:%d" % ln if ln else "?", srcline(ln)))
