"""SASS evidence of the built library (no GPU needed): per kernel the instruction count, the counts of the mnemonics that
matter (TMA / bulk copies / mbarrier / async copies / shuffles / barriers / MUFU) and the TMA + mbarrier lines themselves.
    python tools/sass_excerpt.py > profiles/r2_sass_excerpt.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'mmdet-yolov4_b200', 'csrc', 'libyolopp.so')
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
fns, cur = collections.OrderedDict(), None
for ln in sass.splitlines():
    m = re.search(r'Function : (\S+)', ln)
    if m:
        cur = m.group(1)
        fns[cur] = []
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if cur and m:
        fns[cur].append((m.group(1), m.group(2).strip()))
KEYS = ('UTMALDG', 'UBLKCP', 'SYNCS', 'LDGSTS', 'FENCE', 'ATOMS', 'ATOMG', 'RED', 'BAR.SYNC', 'SHFL', 'MATCH', 'VOTE', 'REDUX', 'MUFU', 'FFMA', 'LDC', 'LDL', 'STL', 'CALL', 'HMMA', 'UTCMMA')
print('cuobjdump -sass', os.path.relpath(lib, ROOT), '(sm_100a)')
for name in fns:
    if not any(k in name for k in ('select_kernel', 'decode_tma_kernelILi0', 'decode_tma_kernelILi1', 'nms_image_kernel', 'mish_fwd_kernelIf', 'decode_dense_kernelILi1', 'decode_rows_kernelILi0')):
        continue
    ins = fns[name]
    cnt = collections.Counter()
    for _, t in ins:
        op = t.split()[1] if t.startswith('@') else t.split()[0]
        for k in KEYS:
            if op.startswith(k):
                cnt[k] += 1
    print('\n%s\n  %d instructions (%.1f KB); ' % (name, len(ins), len(ins) * 16 / 1024) + ', '.join('%s %d' % (k, cnt[k]) for k in KEYS if cnt[k]))
    shown = 0
    for addr, t in ins:
        if any(k in t for k in ('UTMALDG', 'UBLKCP', 'SYNCS')) and shown < 14:
            print('    /*%s*/  %s' % (addr, t))
            shown += 1
