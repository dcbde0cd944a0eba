"""Where does the reference build contract multiply-adds?  (evidence behind csrc/p3p_*_quad.cuh)

nvcc contracts a*b+c in two places: the front end (fma.rn in the PTX) and ptxas, which folds a `mul.f32` whose result
has ONE use into the add/sub that consumes it; when both operands of the add/sub are such products, ptxas folds the
FIRST operand's product (micro-benchmarked: profiles/r02_p3p_fusion_audit.md).  This script annotates the PTX of a
reference translation unit with the folds that rule predicts, flags the two-product sites, and lists the multi-use
products next to an add/sub (undecided by reading alone; settled by the golden-vector search of
tests/test_cpu_p3p_quad.py).  Check of the rule: the number of predicted folds must equal the drop in FADD count
between `ptxas --fmad false` and the default build of the same PTX.

  nvcc -O3 -arch=sm_100a -ptx /root/reference/gpu-kernels/solve_batch_lambdatwist.cu -o lt.ptx
  awk '/\.entry _Z5solve/,/^}/' lt.ptx > lt_solve.ptx
  python tools/ptx_fusion_sites.py lt_solve.ptx > lt_annotated.txt
"""
import re, sys, collections
src = open(sys.argv[1]).read().splitlines()
# collect instructions
ins = []
bb = 0
for ln, line in enumerate(src):
    s = line.strip()
    if not s or s.startswith('//') or s.startswith('.'): continue
    if re.match(r'^\$L__BB\w+:', s):
        bb += 1
        ins.append(dict(ln=ln, label=s, bb=bb)); continue
    m = re.match(r'^(@!?%p\d+\s+)?([a-z0-9_.]+)\s*(.*);$', s)
    if not m: continue
    pred, op, rest = m.group(1), m.group(2), m.group(3)
    args = [a.strip() for a in re.split(r',\s*(?![^{]*})', rest)] if rest else []
    ins.append(dict(ln=ln, pred=pred, op=op, args=args, bb=bb, text=s))
    if op.startswith('bra') or op == 'ret': bb += 1
uses = collections.Counter()
defs = collections.defaultdict(list)
for i, I in enumerate(ins):
    if 'op' not in I: continue
    a = I['args']
    if not a: continue
    isstore = I['op'].startswith('st.')
    srcs = a if isstore or I['op'].startswith(('bra','setp')) and False else a[1:]
    if I['op'].startswith('setp'): srcs = a[1:]
    if isstore: srcs = a
    for s_ in srcs:
        for r in re.findall(r'%f?d?\w*\d+', s_):
            uses[r] += 1
    if I.get('pred'):
        for r in re.findall(r'%p\d+', I['pred']): uses[r] += 1
    if not isstore and not I['op'].startswith(('bra',)):
        d = a[0]
        for r in re.findall(r'%\w+\d+', d):
            defs[r].append(i)
fused = 0; amb = 0
notes = {}
consumed = set()
for i, I in enumerate(ins):
    if 'op' not in I: continue
    if I['op'] in ('add.f32','sub.f32','add.f64','sub.f64'):
        d, a, b = I['args']
        cands = []
        for k, r in enumerate((a, b)):
            if r in defs and len(defs[r]) == 1:
                J = ins[defs[r][0]]
                if J['op'] == ('mul.f32' if I['op'].endswith('f32') else 'mul.f64') and uses[r] == 1 and J['bb'] == I['bb'] and not J.get('pred'):
                    cands.append((k, defs[r][0]))
        if cands:
            fused += 1
            if len(cands) == 2:
                amb += 1
                notes[i] = 'AMBIGUOUS-FUSE(%s|%s)' % (a, b)
            else:
                k, j = cands[0]
                notes[i] = 'PTXAS-FUSE operand %d (%s)' % (k, (a, b)[k])
                consumed.add(j)
print('predicted ptxas fusions', fused, 'ambiguous', amb, file=sys.stderr)
for i, I in enumerate(ins):
    if 'label' in I:
        print(I['label']); continue
    t = I['text']
    if i in notes: t += '      <== ' + notes[i]
    if i in consumed: t += '      (fused into later add/sub)'
    print('%5d  %s' % (I['ln'], t))
# near misses
print('--- near misses', file=sys.stderr)
for i, I in enumerate(ins):
    if 'op' not in I: continue
    if I['op'] in ('add.f32','sub.f32') and i not in notes:
        d, a, b = I['args']
        for k, r in enumerate((a, b)):
            if r in defs:
                for j in defs[r]:
                    J = ins[j]
                    if J.get('op') == 'mul.f32':
                        print('near', I['ln'], I['text'], '| operand', r, 'uses', uses[r], 'ndefs', len(defs[r]), 'bb', J['bb'], I['bb'], 'pred', J.get('pred'), file=sys.stderr)
