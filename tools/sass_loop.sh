#!/bin/bash
# usage: tools/sass_loop.sh lib.so  -> instruction count of k_refine_g<7> and its row loop (the loop holding the PRMT conversions)
f=$(cuobjdump -sass "$1" | grep -o "Function : .*k_refine_gILi7.*" | sed 's/Function : //')
cuobjdump -sass -fun "$f" "$1" | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" > /tmp/sass_refine.txt
echo "k_refine_g<7>: $(wc -l < /tmp/sass_refine.txt) instructions"
python3 - <<'PY'
import re
L=[l.rstrip() for l in open('/tmp/sass_refine.txt')]
addr=lambda l:int(re.search(r'/\*([0-9a-f]+)\*/',l).group(1),16)
A=[addr(l) for l in L]
# backward branches
best=None
for i,l in enumerate(L):
    m=re.search(r'BRA(?:\.U)?\s+(?:\S+\s+)?(0x[0-9a-f]+)',l)
    if m:
        t=int(m.group(1),16)
        if t<A[i]:
            j=A.index(t) if t in A else None
            if j is not None:
                body=L[j:i+1]
                n=sum('PRMT' in b for b in body)
                if n>=10 and (best is None or len(body)<len(best)): best=body
if best:
    print("row loop: %d instructions"%len(best))
    import collections
    c=collections.Counter(re.sub(r'^@!?U?P\d\s+','',re.search(r'\*/\s+(.*?);',b).group(1)).split()[0].split('.')[0] for b in best)
    print(dict(c))
    if __import__('os').environ.get('SHOW'):
        for b in best: print(re.search(r'\*/\s+(.*?);',b).group(1))
PY
