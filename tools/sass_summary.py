"""Per-kernel SASS evidence for the built library: counts of the Blackwell tensor-core / TMA / TMEM
mnemonics (UTCHMMA = tcgen05.mma, UTMALDG = cp.async.bulk.tensor load, LDTM = tcgen05.ld, UTCBAR =
tcgen05.commit, SYNCS = mbarrier ops) and of the FP32 FMA family, from `cuobjdump -sass`.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'dmcnet_b200', '_lib', 'libdmc_b200.so')
KEYS = ['UTCHMMA', 'UTCQMMA', 'UTMALDG', 'UTMASTG', 'LDTM', 'STTM', 'UTCBAR', 'UTCATOMSWS', 'SYNCS', 'FFMA2', 'FFMA',
        'HMMA', 'IMMA']


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], stdout=subprocess.PIPE, stderr=subprocess.STDOUT).stdout.decode()
    archs = collections.Counter(re.findall(r'arch = (sm_\w+)', out))
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r'^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m:
            op = m.group(1).split('.')[0]
            kernels[cur]['TOTAL'] += 1
            for k in KEYS:
                if op == k:
                    kernels[cur][k] += 1
    demangle = subprocess.run(['c++filt'], input='\n'.join(kernels).encode(), stdout=subprocess.PIPE).stdout.decode().splitlines()
    print('library: %s' % os.path.relpath(LIB, ROOT))
    print('cubin architectures: %s' % dict(archs))
    print('%d kernels; columns = instruction counts in the SASS of each kernel\n' % len(kernels))
    hdr = '%-96s %7s ' % ('kernel', 'TOTAL') + ' '.join('%8s' % k for k in KEYS)
    print(hdr)
    tot = collections.Counter()
    for (name, c), dn in zip(kernels.items(), demangle):
        short = re.sub(r'\(.*$', '', dn)
        short = short.replace('dmc::', '')
        if len(short) > 95:
            short = short[:92] + '...'
        print('%-96s %7d ' % (short, c['TOTAL']) + ' '.join('%8d' % c[k] for k in KEYS))
        tot.update(c)
    print('\n%-96s %7d ' % ('ALL KERNELS', tot['TOTAL']) + ' '.join('%8d' % tot[k] for k in KEYS))
    tc = [re.sub(r'\(.*$', '', dn).replace('dmc::', '') for (n, c), dn in zip(kernels.items(), demangle) if c['UTCHMMA']]
    print('\nkernels issuing tcgen05.mma (UTCHMMA): %d' % len(tc))
    for t in tc:
        print('  ' + t)


if __name__ == '__main__':
    sys.exit(main())
