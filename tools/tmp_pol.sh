timeout 300 ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none --cache-control none -s 11 -c 11 --csv --log-file gpurun_out/r1f_launches_throughput.csv python tools/profile_step.py --steps 1 --warmup 1 --policy throughput > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r1f_launches_throughput.csv')) if len(r)>10]
hdr=rows[0]; ik=hdr.index('Kernel Name'); im=hdr.index('Metric Name'); iv=hdr.index('Metric Value'); iid=hdr.index('ID')
d={}
for r in rows[1:]:
    d.setdefault(r[iid],{'k':r[ik][:60]})[r[im]]=float(r[iv].replace(',',''))
tot=0
for k,v in d.items():
    smt=v['gpu__time_duration.sum']/1e3*min(v['launch__grid_size'],148); tot+=smt
    print(v['k'].ljust(62), int(v['launch__grid_size']), round(v['gpu__time_duration.sum']/1e3,2),'us', round(smt),'SM*us')
print('total SM*us', round(tot), '-> /148 =', round(tot/148,1),'us')
PY
