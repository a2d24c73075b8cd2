"""Per-CUDA-source-line stall samples of one kernel: joins `ncu --page source --csv` (SASS view, has the samples) with
`nvdisasm -g` of the cubin (has the line numbers); both list the kernel's instructions in the same order.
    python tools/ncu_lines.py <ncu_sass.csv> <nvdisasm -g output> <mangled kernel name> <source file> [top]"""
import collections, csv, re, sys
sass_csv, dis, name, srcfile = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
lines = open(dis).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.strip().startswith('.section') and ('.text.' + name) in l)
cur, seq = None, []
for l in lines[start + 1:]:
    if l.strip().startswith('.section') and seq:
        break
    m = re.match(r'\s*//## File ".*' + re.escape(srcfile.split('/')[-1]) + r'", line (\d+)', l)
    if m:
        cur = int(m.group(1)); continue
    if re.match(r'\s*//## File', l):
        cur = None; continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        seq.append((m.group(2).strip(), cur))
rows = list(csv.reader(open(sass_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) >= len(hdr) and r[0] != "Address"][:len(seq)]
assert len(data) == len(seq), (len(data), len(seq))
by, bye = collections.Counter(), collections.Counter()
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
bystall = collections.defaultdict(collections.Counter)
for (op, ln), r in zip(seq, data):
    s = int(r[idx['# Samples']] or 0); by[ln] += s; bye[ln] += int(r[idx['Instructions Executed']] or 0)
    for h in stalls:
        bystall[ln][h] += int(r[idx[h]] or 0)
tot = sum(by.values())
src = open(srcfile).read().split('\n')
print("samples", tot)
for ln, s in by.most_common(top):
    st = ", ".join(f"{k[6:]} {v}" for k, v in bystall[ln].most_common(2) if v)
    print(f"{s:5d} {100*s/tot:5.1f}%  inst {bye[ln]:8d}  L{ln}: {src[ln-1].strip()[:90] if ln else '(other file)'}   [{st}]")
