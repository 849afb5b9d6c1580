import re,sys,collections
lines=open(sys.argv[1]).read().split('\n')
ins=[]
for l in lines:
    m=re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);',l)
    if m: ins.append((int(m.group(1),16),m.group(2).strip()))
def op(t):
    t=re.sub(r'^@!?U?P\d+\s+','',t)
    return t.split()[0]
def stats(s,name):
    c=collections.Counter(op(t) for a,t in s)
    tot=len(s)
    wide=sum(v for k,v in c.items() if k.startswith('IMAD.WIDE'))
    mov=sum(v for k,v in c.items() if k.startswith('IMAD.MOV'))
    oth=sum(v for k,v in c.items() if k.startswith('IMAD') ) - wide - mov
    call=sum(v for k,v in c.items() if k.startswith('CALL'))
    ldl=sum(v for k,v in c.items() if k.startswith('LDL')); stl=sum(v for k,v in c.items() if k.startswith('STL'))
    print(f"{name:28s} n={tot:5d} WIDE={wide:4d} IMAD.MOV={mov:4d} otherIMAD={oth:4d} MOV={c.get('MOV',0):3d} LOP3={c.get('LOP3.LUT',0):4d} IADD3={c.get('IADD3',0)+c.get('IADD3.X',0):4d} CALL={call:3d} LDL={ldl:3d} STL={stl:3d}")
    return dict(n=tot,wide=wide,mov=mov,oth=oth)
segs=[];cur=[]
for a,t in ins:
    cur.append((a,t))
    if op(t).startswith('RET') or (op(t)=='EXIT' and not t.startswith('@')):
        segs.append(cur);cur=[]
for s in segs: stats(s,f"seg {s[0][0]:#x}-{s[-1][0]:#x}")
# loops
for a,t in ins:
    m=re.search(r'BRA\s+(?:P\d,\s*)?(0x[0-9a-f]+)',t)
    if m and int(m.group(1),16)<a and a-int(m.group(1),16)>0x800:
        lo=int(m.group(1),16)
        stats([(x,y) for x,y in ins if lo<=x<=a],f"loop {lo:#x}-{a:#x}")
