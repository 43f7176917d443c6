"""Compare op_times.py outputs: totals and per-shape GEMM times.  python tools/op_diff.py base.tsv other.tsv [...]"""
import collections, sys
def load(p):
    d = {}
    for ln in open(p):
        f = ln.rstrip('\n').split('\t'); d[int(f[0])] = (f[1], f[2], int(f[3]), int(f[4]), int(f[5]), int(f[6]), float(f[7]))
    return d
files = [a for a in sys.argv[1:] if a != '-v']
D = [load(f) for f in files]
tot = lambda d, k=None, halo=None: sum(v[6] for v in d.values() if (k is None or v[1] == k))
for f, d in zip(files, D):
    voc = sum(v[6] for v in d.values() if v[1] == 'tc' and v[0].startswith('vocoder'))
    print(f"{f:44s} all {tot(d):9.1f}  tc {tot(d, 'tc'):9.1f}  vocoder-tc {voc:8.1f}  other-tc {tot(d, 'tc') - voc:8.1f}")
if '-v' in sys.argv: sys.exit()
grp = collections.OrderedDict()
for i, v in D[0].items():
    if v[1] != 'tc': continue
    k = (v[0].split('.')[0], v[2], v[3], v[4], v[5])
    a = grp.setdefault(k, [0] + [0.0] * len(D)); a[0] += 1
    for j, d in enumerate(D): a[1 + j] += d[i][6]
for k, a in sorted(grp.items(), key=lambda kv: -kv[1][1])[:34]:
    print(f"{str(k):44s} n={a[0]:2d} " + " ".join(f"{x:8.1f}" for x in a[1:]))
