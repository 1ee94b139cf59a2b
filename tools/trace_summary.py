import sys,re,collections
ms=collections.OrderedDict(); n=collections.Counter(); tf={}
seen=False
for line in open(sys.argv[1]):
    if 'traced pass' in line: seen=True; continue
    if not seen: continue
    m=re.match(r'\[afb200\] (.*?)\s+([0-9.]+) ms\s+([0-9.]+) TFLOP/s\s+([0-9.]+) GB/s',line)
    if not m: continue
    k=m.group(1); ms[k]=ms.get(k,0)+float(m.group(2)); n[k]+=1; tf[k]=(float(m.group(3)),float(m.group(4)))
tot=sum(ms.values()); B=int(sys.argv[2]) if len(sys.argv)>2 else 32
print('total %.3f ms for %d clips = %.3f ms/clip'%(tot,B,tot/B))
for k,v in sorted(ms.items(), key=lambda kv:-kv[1])[:int(sys.argv[3]) if len(sys.argv)>3 else 40]: print('%7.3f ms (%4.1f%%) x%3d  %-52s last: %6.1f TF/s %6.1f GB/s'%(v,100*v/tot,n[k],k,tf[k][0],tf[k][1]))
