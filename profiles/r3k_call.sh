#!/bin/bash
mkdir -p gpurun_out
NCU_ONE_STEP=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r3k_config5_launches.csv python profiles/profile_config5_host.py > gpurun_out/r3k.log 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/r3k_config5_launches.csv')))
for i,r in enumerate(rows):
    if r and r[0]=='ID': hdr=r; start=i+1; break
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
tot=collections.Counter(); cnt=collections.Counter()
for r in rows[start:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    n=r[ki].split('(')[0].replace('void ','')[:70]; tot[n]+=v/1e3; cnt[n]+=1
T=sum(tot.values()); print('launches',sum(cnt.values()),'total us',round(T,1))
for k,v in tot.most_common(25): print(f'{k:72s} {cnt[k]:5d} {v:9.1f} us  avg {v/cnt[k]:7.1f}  {v/T:.3f}')
PY
