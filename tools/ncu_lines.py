import csv, collections, sys, subprocess
rep=sys.argv[1]; topn=int(sys.argv[2]) if len(sys.argv)>2 else 40
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=csv.reader(out.splitlines())
cur_file=None; lines={}; kn=None; first=None
for r in rows:
    if not r: continue
    if r[0]=='File Path': cur_file=r[1].split('/')[-1]; continue
    if r[0]=='Function Name':
        kn=r[1]
        if first is None: first=kn
        continue
    if r[0]=='Line No': continue
    if r[0] and r[0].isdigit() and len(r)>10:
        key=(cur_file,int(r[0]))
        try: ie=int(r[7]); smp=int(r[6]); thr=int(r[8])
        except: continue
        if key in lines: continue
        lines[key]=(ie,smp,r[1].strip()[:100],thr)
tot=sum(v[0] for v in lines.values()); ts=sum(v[1] for v in lines.values()); print(first); print('total inst',tot,'samples',ts)
byfile=collections.Counter()
for (f,l),v in lines.items(): byfile[f]+=v[0]
print(byfile.most_common(6))
key=(lambda kv:-kv[1][1]) if len(sys.argv)>3 and sys.argv[3]=='smp' else (lambda kv:-kv[1][0])
for (f,l),v in sorted(lines.items(), key=key)[:topn]:
    print(f'{f}:{l:4d} {v[0]:9d} {100*v[0]/max(tot,1):5.1f}% smp {v[1]:6d} {100*v[1]/max(ts,1):4.1f}% thr/inst {v[3]/max(v[0],1):4.1f}  {v[2]}')
