"""opcode histogram of the hot loop (largest backward-branch body that holds FFMA2) of mc_per_bin_kernel in a k1_mix executable"""
import re, sys, collections, subprocess
def hist(path, kernel='mc_per_bin_kernel'):
    out = subprocess.run(['cuobjdump', '-sass', path], capture_output=True, text=True).stdout
    i = out.index(kernel)
    ins = []
    for l in out[i:].splitlines():
        if 'Function :' in l and ins: break
        m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
        if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
    best = None
    for a, t in ins:
        m = re.search(r'BRA\s+(0x[0-9a-f]+)', t)
        if m and 'BRA.U' not in t:
            tgt = int(m.group(1), 16)
            if tgt < a:
                body = [x for x in ins if tgt <= x[0] <= a]
                n = sum('FFMA2' in x[1] for x in body)
                if n and (best is None or len(body) > len(best)): best = body
    c = collections.Counter()
    for a, t in best:
        t = re.sub(r'^@!?U?P\d+\s+', '', t)
        op = t.split()[0]
        op = '.'.join(op.split('.')[:2]) if op.startswith('IMAD') else op.split('.')[0]
        c[op] += 1
    return len(best), c
if __name__ == '__main__':
    for p in sys.argv[1:]:
        n, c = hist(p)
        print(p.split('/')[-1], 'loop instrs', n, dict(c.most_common(16)))
