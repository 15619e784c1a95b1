"""Developer tool: per-line instruction counts of a source line range: ncu_line_range.py <rep> <lo> <hi> [launch]"""
import csv, io, subprocess, sys
rep=sys.argv[1]; lo=int(sys.argv[2]); hi=int(sys.argv[3])
txt = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],capture_output=True,text=True).stdout
blocks,cur=[],[]
for row in csv.reader(io.StringIO(txt)):
    if row and row[0]=="File Path":
        if cur: blocks.append(cur)
        cur=[]
    cur.append(row)
if cur: blocks.append(cur)
b=[b for b in blocks if b[0][1].endswith(".cu")][int(sys.argv[4]) if len(sys.argv)>4 else 0]
hdr=next(r for r in b if r and r[0]=="Line No"); ix={h:i for i,h in enumerate(hdr)}
rows=[r for r in b if r and r[0].isdigit()]
ti=sum(int(r[ix["Instructions Executed"]]) for r in rows)
for r in rows:
    n=int(r[0])
    if lo<=n<=hi and int(r[ix["Instructions Executed"]])>0:
        print(f"{n:5d} {100*int(r[ix['Instructions Executed']])/ti:5.2f}%inst {int(r[ix['Instructions Executed']])/12000/1:9.0f}/ray | {r[1].strip()[:120]}")
