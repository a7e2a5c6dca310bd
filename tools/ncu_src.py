"""Hot source lines of a kernel from an .ncu-rep, joined with nvdisasm line info of the built library.
   python tools/ncu_src.py <rep> <kernel-substring> [topN]
Ranks source lines by warp-stall samples and prints the dominant stall reasons per line."""
import csv, sys, subprocess, collections, re, os, tempfile, glob
rep, kname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
sort_key = sys.argv[4] if len(sys.argv) > 4 else 'samples'   # 'samples' or 'inst'
lib = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'mmdet-yolov4_b200', 'csrc', 'libyolopp.so')
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = glob.glob(os.path.join(tmp, '*.cubin'))[0]
dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
# parse: function sections start with ".text.<mangled>:" ; line markers '//## File "x", line N'
lines_of = collections.defaultdict(list)
cur_fn, cur_line = None, (None, 0)
for ln in dis.splitlines():
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', ln)
    if m: cur_fn = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur_line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if cur_fn and re.match(r'\s+/\*[0-9a-f]{4,}\*/', ln):
        lines_of[cur_fn].append(cur_line)
fn = [f for f in lines_of if kname in f]
assert fn, f'no function matching {kname}: {list(lines_of)[:20]}'
# (template instances: the benchmark's head is the CSP convention, MODE 0)
fn = ([f for f in fn if 'ILi0E' in f] or fn)[0]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
body = rows[2:]
srcmap = lines_of[fn]
print(f'function {fn}: {len(srcmap)} sass instr (nvdisasm) vs {len(body)} (ncu)')
agg = collections.defaultdict(lambda: collections.Counter())
tot = 0
for k, r in enumerate(body):
    if k >= len(srcmap): break
    s = float(r[ci['# Samples']] or 0)
    tot += s
    a = agg[srcmap[k]]
    a['samples'] += s
    a['inst'] += float(r[ci['Instructions Executed']] or 0)
    for sc in stall_cols:
        a[sc] += float(r[ci[sc]] or 0)
srcfiles = {}
def srcline(f, n):
    if f not in srcfiles:
        p = os.path.join(os.path.dirname(os.path.abspath(lib)), f)
        srcfiles[f] = open(p).read().splitlines() if os.path.isfile(p) else []
    L = srcfiles[f]
    return L[n - 1].strip()[:110] if 0 < n <= len(L) else ''
print('total samples', tot, ' total warp instructions', sum(a['inst'] for a in agg.values()))
for (f, n), a in sorted(agg.items(), key=lambda kv: -kv[1][sort_key])[:top]:
    st = sorted(((a[s], s[6:]) for s in stall_cols if a[s] > 0), reverse=True)[:3]
    print(f"{a['samples']:8.0f} {100*a['samples']/max(tot,1):5.1f}% inst={a['inst']:10.0f} {f}:{n:<4d} {srcline(f,n)}   [{', '.join(f'{n2}:{v:.0f}' for v,n2 in st)}]")
