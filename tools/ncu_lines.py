"""Aggregate an ncu SASS source page (csv) per CUDA source line using nvdisasm -g line info.
usage: ncu_lines.py src.csv kernel.sass kernel_name_substring [template_arg_filter]"""
import csv, re, collections, sys
src_csv, sass, kname = sys.argv[1], sys.argv[2], sys.argv[3]
want = sys.argv[4] if len(sys.argv) > 4 else None
addr2line = {}
cur = None; infn = False
for ln in open(sass):
    if '.section' in ln and '.text.' in ln:
        infn = kname in ln and (want is None or want in ln)
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);', ln)
    if m and infn and cur:
        addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ai, si, ii = hdr.index('Address'), hdr.index('# Samples'), hdr.index('Instructions Executed')
base = None
samp = collections.Counter(); inst = collections.Counter(); ts = ti = 0
for r in rows[2:]:
    try:
        a = int(r[ai], 16); n = int(r[si]); k = int(r[ii])
    except Exception:
        continue
    if base is None: base = a
    key = addr2line.get(a - base, ('?', 0))
    samp[key] += n; inst[key] += k; ts += n; ti += k
srcs = {}
def text(f, l):
    if f not in srcs:
        try: srcs[f] = open('/root/repo/k2transducerasr_b200/csrc/' + f).read().split('\n')
        except Exception: srcs[f] = []
    return srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ''
print(f"total warp-instructions executed {ti}, samples {ts}, mapped {len(addr2line)}")
print("--- by executed instructions")
for (f, l), k in inst.most_common(40):
    print(f"{100 * k / ti:5.1f}% inst {100 * samp[(f, l)] / max(ts, 1):5.1f}% smp  {f}:{l:<4} {text(f, l)}")
print("--- by stall samples")
for (f, l), k in samp.most_common(45):
    print(f"{100 * k / max(ts, 1):5.1f}% smp {100 * inst[(f, l)] / ti:5.1f}% inst  {f}:{l:<4} {text(f, l)}")
