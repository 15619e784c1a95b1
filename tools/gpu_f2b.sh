mkdir -p gpurun_out
python tools/build_world_gpu.py 1024 3
python tools/build_world_gpu.py 2048 2
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/f2_launches.csv python tools/build_world_gpu.py 1024 1 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/f2_launches.csv')) if len(r)>10 and r[0]!='ID']
agg=collections.OrderedDict()
for r in rows:
    k=r[4].split('(')[0][-70:]; a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(r[14].replace(',',''))/1e6
for k,(n,t) in agg.items(): print(f"{t:9.3f} ms {n:4d}x {k}")
print("total %.3f ms"%sum(t for n,t in agg.values()))
PY
