"""Aggregate `ncu -i rep --page source --csv --print-source cuda,sass` output by CUDA source line:
share of warp-stall samples, instructions executed and the top stall reasons per line.

    ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass --launch-skip N --launch-count 1 > src.csv
    python profiles/tools/ncu_source_by_line.py src.csv 30
"""
import csv, sys, collections
path=sys.argv[1]; topn=int(sys.argv[2]) if len(sys.argv)>2 else 40
rows=list(csv.reader(open(path)))
cur_file=None; hdr=None
agg=[]
for r in rows:
    if not r: continue
    if r[0]=='File Path': cur_file=r[1].split('/')[-1]; continue
    if r[0]=='Function Name': continue
    if r[0]=='Line No': hdr=r; continue
    if r[0] and hdr:
        d=dict(zip(hdr,r))
        try: s=int(d['Warp Stall Sampling (All Samples)'])
        except: continue
        st={k:int(v) for k,v in d.items() if k.startswith('stall_') and 'Not Issued' not in k and v.isdigit() and int(v)>0}
        inst=d.get('Instructions Executed','0')
        agg.append((s,cur_file,r[0],r[1].strip()[:110],inst,st))
tot=sum(a[0] for a in agg)
print('total samples',tot)
for s,f,l,src,inst,st in sorted(agg,reverse=True)[:topn]:
    top=sorted(st.items(),key=lambda kv:-kv[1])[:3]
    print(f'{s/tot*100:5.1f}% {f}:{l} inst={inst} {top}\n        {src}')
