"""Summarise one step from an ncu `--metrics gpu__time_duration.sum --csv` launch list."""
import csv, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size'); bi=hdr.index('Block Size')
data=[(r[ki], float(r[vi].replace(',','')), r[gi], r[bi]) for r in rows[1:]]
end=sys.argv[2] if len(sys.argv)>2 else 'adam'
idxs=[i for i,(k,v,g,b) in enumerate(data) if end in k]
a=idxs[-2]+1; b=idxs[-1]+1
tot=sum(v for k,v,g,bb in data[a:b]); small=0
for k,v,g,bb in data[a:b]:
    if v < 4500: small += v; continue
    print(f"{v/1000:9.1f} us {100*v/tot:5.1f}%  {g:>14s} {bb:>12s} {k[:70]}")
print(f"{small/1000:9.1f} us {100*small/tot:5.1f}%  ({sum(1 for k,v,g,bb in data[a:b] if v<4500)} launches < 4.5 us)")
print("step total us", round(tot/1000,1), "launches", b-a)
