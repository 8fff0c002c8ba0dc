#!/usr/bin/env python3
"""Top warp-stall sampling sites (SASS) of the first kernel in an .ncu-rep; with cumulative position."""
import csv, io, subprocess, sys
def main(path, top=35):
    out = subprocess.run(['ncu','-i',path,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = rows[1]
    si, ii, ai = h.index('# Samples'), h.index('Instructions Executed'), h.index('Source')
    data=[]
    for n,r in enumerate(rows[2:]):
        if len(r)<=si or not r[si].isdigit(): 
            if len(r)>1 and r[0]=='Kernel Name': break
            continue
        data.append((n,int(r[si]),int(r[ii] or 0),r[ai].strip()))
    tot=sum(d[1] for d in data)
    print(f'{len(data)} SASS instructions, {tot} samples')
    # coarse histogram by position (10 buckets)
    nb=20; B=[0]*nb
    for d in data: B[min(nb-1,d[0]*nb//len(data))]+=d[1]
    print('samples by code position (5% buckets):',' '.join(f'{100*b/tot:.0f}' for b in B))
    for d in sorted(data,key=lambda x:-x[1])[:top]:
        print(f'{d[0]:5d} {100*d[1]/tot:5.1f}%  exec={d[2]:8d}  {d[3][:110]}')
if __name__=='__main__': main(sys.argv[1], int(sys.argv[2]) if len(sys.argv)>2 else 35)
