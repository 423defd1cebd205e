import csv,sys,collections
rows=list(csv.reader(open(sys.argv[1])))
hdr=None; cur=None; agg=collections.OrderedDict(); files=None
stall_cols={}
for r in rows:
    if r and r[0]=='File Path': files=r[1]; continue
    if r and r[0]=='Line No':
        hdr=r; si=hdr.index('Warp Stall Sampling (All Samples)')
        stall_cols={h:i for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h}
        continue
    if hdr is None or len(r)<=si: continue
    if r[0].strip().isdigit():
        try: s=int(r[si])
        except: s=0
        key=(files.split('/')[-1], int(r[0]), r[1].strip())
        st={h:int(r[i]) if r[i].isdigit() else 0 for h,i in stall_cols.items()}
        if key in agg:
            agg[key][0]+=s
            for h in st: agg[key][1][h]+=st[h]
        else: agg[key]=[s,st]
tot=sum(v[0] for v in agg.values())
print('total samples',tot)
N=int(sys.argv[2]) if len(sys.argv)>2 else 40
for (f,ln,src),(s,st) in sorted(agg.items(), key=lambda kv:-kv[1][0])[:N]:
    top=sorted(st.items(), key=lambda kv:-kv[1])[:3]
    print(f"{s:6d} {100*s/tot:5.1f}% {f}:{ln:5d} {src[:70]:70s} | "+' '.join(f"{h[6:]}={v}" for h,v in top if v))
# totals by stall
tt=collections.Counter()
for v in agg.values():
    for h,x in v[1].items(): tt[h]+=x
print({h[6:]:round(100*x/sum(tt.values()),1) for h,x in tt.most_common(10)})
