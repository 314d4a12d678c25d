import csv,sys,subprocess
rep=sys.argv[1]
raw=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; vals=rows[-1]
m=dict(zip(hdr,vals))
for k in ['gpu__time_duration.sum','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__warps_eligible.avg.per_cycle_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed']:
    print(k, m.get(k))
for k,v in m.items():
    if 'issue_stalled' in k and 'per_issue_active' in k and float(v)>0.1: print(k.replace('smsp__average_warps_issue_stalled_',''),v)
src=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]; data=rows[2:]
iA=hdr.index('Address'); iS=hdr.index('Source'); iE=hdr.index('Instructions Executed'); iT=hdr.index('Thread Instructions Executed')
tot=sum(int(r[iE]) for r in data); tott=sum(int(r[iT]) for r in data)
print("total warp inst %.3fG thread %.1fG avg thr %.2f"%(tot/1e9,tott/1e9,tott/tot))
from collections import Counter
c=Counter()
for r in data:
    t=r[iS].split()
    op=t[1] if t[0].startswith('@') else t[0]
    c[op.split('.')[0]]+=int(r[iE])
print(" ".join("%s %.1f%%"%(k,100*v/tot) for k,v in c.most_common(22)))
base=int(data[0][iA],16)
groups=[]; cur=None
for r in data:
    e=int(r[iE]); off=int(r[iA],16)-base
    if cur and abs(e-cur[2])<=0.2*max(cur[2],1): cur[1]=off; cur[3]+=e; cur[4]+=int(r[iT]); cur[5]+=1
    else:
        cur=[off,off,e,e,int(r[iT]),1]; groups.append(cur)
for g in groups:
    if g[3]>0.01*tot: print("0x%04x-0x%04x n=%3d execs/inst %.1fM share %.1f%% avgthr %.1f"%(g[0],g[1],g[5],g[2]/1e6,100*g[3]/tot,g[4]/max(g[3],1)))
