"""SASS opcode summary of the shipped library (CPU only): which kernels contain tcgen05 / TMEM / TMA / multicast opcodes.
  python tools/sass_summary.py > profiles/sass_opcodes_r02.csv
UTCHMMA / UTCQMMA = tcgen05.mma (kind::f16/tf32), LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor loads, UTMASTG = TMA stores,
UTCBAR = tcgen05.commit -> mbarrier, UBLKCP = cp.async.bulk, LDGMC / STGMC(REDG MC) = multimem.ld_reduce / multimem.st,
SYNCS = mbarrier ops, HMMA / FFMA for comparison."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'acme_b200', 'lib', 'libb200rl.so')
OPS = ['UTCHMMA', 'UTCQMMA', 'UTCMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTCBAR', 'UBLKCP', 'LDGMC', 'STGMC', 'SYNCS', 'HMMA', 'FFMA', 'MUFU',
       'ACQBULK', 'PREEXIT']   # griddepcontrol.wait / griddepcontrol.launch_dependents (programmatic dependent launch)


def main():
  sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
  demangle = lambda n: subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip()
  counts, name = collections.OrderedDict(), None
  for line in sass.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
      name = m.group(1)
      counts[name] = collections.Counter()
      continue
    if name is None:
      continue
    m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m:
      op = m.group(1)
      counts[name]['_total'] += 1
      for o in OPS:
        if op.startswith(o):
          counts[name][o] += 1
          break
  w = sys.stdout
  w.write('kernel,instructions,' + ','.join(OPS) + '\n')
  total = collections.Counter()
  for n, c in counts.items():
    if not any(c[o] for o in OPS if o not in ('FFMA', 'MUFU', 'SYNCS')) and '--all' not in sys.argv:
      continue
    short = re.sub(r'\(.*', '', demangle(n))[:100].replace(',', ';')
    w.write(short + ',' + str(c['_total']) + ',' + ','.join(str(c[o]) for o in OPS) + '\n')
    total.update({o: c[o] for o in OPS})
  w.write('TOTAL (listed kernels),,' + ','.join(str(total[o]) for o in OPS) + '\n')


main()
