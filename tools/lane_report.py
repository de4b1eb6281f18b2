#!/usr/bin/env python3
"""Lane-occupancy report per call site: joins an ncu SASS source page (CSV) with nvdisasm's inline chains.

  ncu -i rep.ncu-rep --page source --csv --print-source sass > sass.csv
  nvdisasm -g -gi -c variant.cubin > variant.dis
  tools/lane_report.py sass.csv variant.dis <mangled kernel name> [depth]

For every instruction the chain of "inlined at" locations is reduced to its outermost `depth` frames, and
instructions / thread-instructions are summed per chain prefix: where do the issue slots go, and how many lanes
are active there.
"""
import csv, re, sys, collections

def parse_dis(path, kernel):
    chains = {}
    cur = []; block = []; inside = False; in_block = False
    for line in open(path, errors='replace'):
        if line.startswith('//---') and '.text.' in line:
            inside = ('.text.' + kernel + ' ') in line
            continue
        if not inside: continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
        if m:
            if not in_block: block = []; in_block = True
            block.append((m.group(1).split('/')[-1], int(m.group(2))))
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);', line)
        if m:
            if in_block: cur = block; in_block = False
            chains[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return chains

def main():
    sass_csv, dis, kernel = sys.argv[1:4]
    depth = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    chains = parse_dis(dis, kernel)
    rows = list(csv.reader(open(sass_csv)))
    hdr = None; base = None
    agg = collections.defaultdict(lambda: [0, 0, 0])
    mismatch = 0
    for r in rows:
        if r and r[0] == 'Address': hdr = r; continue
        if hdr is None or len(r) < len(hdr) - 2: continue
        addr = int(r[0], 16)
        if base is None: base = addr
        off = addr - base
        ie = int(r[hdr.index('Instructions Executed')]); te = int(r[hdr.index('Thread Instructions Executed')])
        sm = int(r[hdr.index('# Samples')])
        ch, text = chains.get(off, ([], '?'))
        if text.split()[0:1] != r[1].split()[0:1] and not r[1].strip().startswith('@'): mismatch += 1
        outer = tuple(reversed(ch))[:depth]   # outermost first
        a = agg[outer]; a[0] += ie; a[1] += te; a[2] += sm
    ti = sum(a[0] for a in agg.values()); tt = sum(a[1] for a in agg.values()); ts = sum(a[2] for a in agg.values())
    print(f'# instructions {ti}  lanes/inst {tt/ti:.2f}  opcode mismatches {mismatch}')
    print(f'# {"call chain (outermost first)":60s} inst%  lanes  samples%  lost-lane%')
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if a[0] * 1000 < ti: continue
        name = ' > '.join(f'{f}:{l}' for f, l in k)
        print(f'{name:62s} {a[0]/ti*100:5.2f}  {a[1]/max(a[0],1):5.1f}  {a[2]/ts*100:6.2f}  {(a[0]*32-a[1])/(ti*32-tt)*100:6.2f}')

if __name__ == '__main__':
    main()
