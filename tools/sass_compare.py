"""Per-kernel SASS comparison of two objects built by nvcc (addresses and column padding stripped).
Usage: python tools/sass_compare.py old.o new.o — used to show which kernels of HEAD are byte-identical to the last build that ran on a GPU."""
import subprocess, sys, re, hashlib, collections
def funcs(obj):
    out = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
    d = {}; cur = None
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1); d[cur] = []
        elif cur and re.match(r'\s+/\*[0-9a-f]{4}\*/', line):
            d[cur].append(re.sub(r'\s+', ' ', re.sub(r'/\*[0-9a-f]{4}\*/', '', line, count=1)).strip())
    return {k: hashlib.md5('\n'.join(v).encode()).hexdigest() for k, v in d.items()}
a, b = funcs(sys.argv[1]), funcs(sys.argv[2])
dem = lambda n: subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip().split('(')[0]
same = [k for k in a if k in b and a[k] == b[k]]; diff = [k for k in a if k in b and a[k] != b[k]]
print('identical:', len(same), 'different:', len(diff), 'only old:', len([k for k in a if k not in b]), 'only new:', len([k for k in b if k not in a]))
names = collections.Counter(dem(k).split('<')[0] for k in same)
print(' identical kernels:', dict(names))
print(' different kernels:', dict(collections.Counter(dem(k).split('<')[0] for k in diff)))
print(' only in new:', sorted({dem(k).split('<')[0] for k in b if k not in a}))
