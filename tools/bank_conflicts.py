#!/usr/bin/env python3
"""
tools/bank_conflicts.py [h] -- shared-memory wavefronts per 128-bit row load of the blocked polynomial (poly_general.cu), from
the real colouring list of h = n - 5 high reads: the threads of a quarter warp take eight consecutive colourings, and a
128-bit load serves the eight 16-byte bank groups once per wavefront.  Compares the plain layout with the 3-bit XOR fold that
k_general uses (table_swizzle in poly_general.cuh) and with a GF(8) column code, for 32-byte rows (16-bit counts) and
128-byte rows (e(S) of the distant model).  CPU only.
"""
import numpy as np, sys
def terms(h):
    out=[]
    full=(1<<h)-1
    for a in range(full+1):
        rest=full^a
        if not rest: continue
        hi=1<<(rest.bit_length()-1); sub=rest^hi
        b=sub
        while True:
            out.append((a,b,rest^b))
            if not b: break
            b=(b-1)&sub
    return np.array(out,dtype=np.int64)
def gf8cols(n, start=1):
    # alpha^k in GF(8) with x^3=x+1 : 1,2,4,3,6,7,5
    seq=[1,2,4,3,6,7,5]
    return [seq[(start+p)%7] for p in range(n)]
def hash_rows(rows, cols):
    h=np.zeros_like(rows)
    for p,c in enumerate(cols):
        h^=((rows>>p)&1)*c
    return h
def wavefronts(rows, slot, T=512):
    # threads tid handle e=tid+k*T ; quarter-warp = 8 consecutive tids at same k -> 8 consecutive e (if within range)
    n=len(rows)//8*8
    r=rows[:n].reshape(-1,8); s=slot[:n].reshape(-1,8)
    tot=0
    # count distinct (row) per slot, max over slots
    mx=np.zeros(len(r),dtype=np.int64)
    for sl in range(8):
        # number of distinct rows with this slot
        cnt=np.zeros(len(r),dtype=np.int64)
        for i in range(8):
            m=(s[:,i]==sl)
            # distinct: not equal to any previous j<i with same slot
            dup=np.zeros(len(r),dtype=bool)
            for j in range(i):
                dup|=(r[:,j]==r[:,i])&(s[:,j]==sl)
            cnt+=(m&~dup)
        mx=np.maximum(mx,cnt)
    return mx.mean()
h=int(sys.argv[1]) if len(sys.argv)>1 else 11
t=terms(h)
for name,rowbytes in (('u16 rows (32 B)',32),('f64 rows (128 B)',128)):
    for role,col in (('Y',1),('Z',2)):
        rows=t[:,col]
        if rowbytes==32:
            base=(rows&3)*2      # unit slot of half q=0
            res={}
            res['none']=wavefronts(rows, base&7)
            f3=np.zeros_like(rows)
            for k in range(0,16,3): f3^=(rows>>(2+k))&7
            res['fold3']=wavefronts(rows,(base^f3)&7)
            g=hash_rows(rows>>2, gf8cols(14,start=3))
            res['gf8']=wavefronts(rows,(base^g)&7)
        else:
            base=np.zeros_like(rows)
            res={}
            res['none']=wavefronts(rows, base)
            f3=np.zeros_like(rows)
            for k in range(0,16,3): f3^=(rows>>k)&7
            res['fold3']=wavefronts(rows,f3&7)
            g=hash_rows(rows, gf8cols(14,start=0))
            res['gf8']=wavefronts(rows,g&7)
        print(name,role,{k:round(v,2) for k,v in res.items()})
