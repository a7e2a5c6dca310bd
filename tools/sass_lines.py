"""Static SASS size per source line of one kernel (nvdisasm line info): where does the code size go?
   python tools/sass_lines.py <kernel-substring> [topN] [lib]"""
import sys, subprocess, collections, re, os, tempfile, glob
kname = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'mmdet-yolov4_b200', 'csrc', 'libyolopp.so')
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = glob.glob(os.path.join(tmp, '*.cubin'))[0]
dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
cnt = collections.defaultdict(collections.Counter); cur_fn, cur_line = None, (None, 0)
for ln in dis.splitlines():
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', ln)
    if m: cur_fn = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur_line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if cur_fn and re.match(r'\s+/\*[0-9a-f]{4,}\*/', ln): cnt[cur_fn][cur_line] += 1
fn = [f for f in cnt if kname in f][0]
tot = sum(cnt[fn].values()); print(fn, tot, 'instr', tot * 16 / 1024, 'KB')
src = {}
def line(f, n):
    if f not in src:
        p = os.path.join(os.path.dirname(os.path.abspath(lib)), f); src[f] = open(p).read().splitlines() if os.path.isfile(p) else []
    return src[f][n - 1].strip()[:100] if 0 < n <= len(src[f]) else ''
for (f, n), c in cnt[fn].most_common(top): print('%6d %5.1f%%  %s:%d  %s' % (c, 100 * c / tot, f, n, line(f, n)))
