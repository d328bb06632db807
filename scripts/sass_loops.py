"""Opcode histogram and issue-port model of the p = 2 point loops (two / four points per reciprocal) from the SASS of
idw_launch.o -- needs no GPU:   cuobjdump -sass v2r_b200/lib/idw_launch.o > /tmp/all.sass; python scripts/sass_loops.py /tmp/all.sass
An FP64 instruction holds a sub-partition's issue port for two cycles, any other instruction for one."""
import re,collections,sys
def loops(path):
    cur=None; fn={}
    for l in open(path):
        m=re.search(r'Function : (\S+)',l)
        if m: cur=m.group(1); fn[cur]=[]; continue
        m=re.match(r'\s+/\*([0-9a-f]{4})\*/\s+(.*?);',l)
        if m and cur: fn[cur].append((int(m.group(1),16),m.group(2)))
    out=[]
    for name,ins in fn.items():
        addr={a:i for i,(a,_) in enumerate(ins)}
        for i,(a,t) in enumerate(ins):
            m=re.search(r'BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)',t)
            if m:
                tgt=int(m.group(1),16)
                if tgt<a and tgt in addr:
                    body=ins[addr[tgt]:i+1]
                    if sum('MUFU' in x for _,x in body)==16 and len(body)<500 and sum('DMUL' in x for _,x in body)>=40:
                        c=collections.Counter((x.split()[1] if x.startswith('@') else x.split()[0]) for _,x in body)
                        out.append((name,tgt,a,len(body),c))
    return out
for name,t,a,n,c in loops(sys.argv[1]):
    f64=c['DFMA']+c['DADD']+c['DMUL']
    print(name); print(f"  loop {t:#x}-{a:#x}: {n} instructions = {f64} FP64 (DFMA {c['DFMA']}, DADD {c['DADD']}, DMUL {c['DMUL']}) + {n-f64} other "+str({k:v for k,v in c.items() if k not in('DFMA','DADD','DMUL')}))
    print(f"  issue-port model: 2 x {f64} + {n-f64} = {2*f64+n-f64} cycles per warp and iteration")
