"""Per-instruction attribution of an `ncu --set full --import-source on` report: for one kernel, every memory
instruction with its executions, L1 tag requests, L2 sectors and stall samples, the 25 instructions with the most
stall samples, the stall-reason totals.   python tools/ncu_src_attrib.py REPORT.ncu-rep KERNEL_REGEX [INDEX]"""
import csv,sys,subprocess,io,collections
rep, kname, idx = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv)>3 else 0
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--kernel-name','regex:'+kname],capture_output=True,text=True).stdout
# multiple kernels -> multiple sections starting with "Kernel Name"
secs=[]; cur=None
for r in csv.reader(io.StringIO(out)):
    if r and r[0]=='Kernel Name': cur=[]; secs.append(cur); continue
    if cur is not None: cur.append(r)
sec=secs[idx]
hdr=sec[0]; ix={h:i for i,h in enumerate(hdr)}
data=[r for r in sec[1:] if r and r[0].startswith('0x')]
tot_inst=sum(int(r[ix['Instructions Executed']]) for r in data)
tot_samp=sum(int(r[ix['# Samples']]) for r in data)
tot_tag=sum(int(r[ix['L1 Tag Requests Global']]) for r in data)
tot_l2=sum(int(r[ix['L2 Theoretical Sectors Global']]) for r in data)
print('kernels',len(secs),'instr',tot_inst,'samples',tot_samp,'L1 tag req',tot_tag,'L2 sectors',tot_l2)
print('--- memory instructions')
for n,r in enumerate(data):
    src=r[ix['Source']].strip()
    if any(k in src for k in ('LDG','STG','LDS','STS','RED','ATOM')):
        print('%4d %-58s exec %9s tag %10s l2sec %10s shwave %8s samp %6s thr/inst %s'%(n,src[:58],r[ix['Instructions Executed']],r[ix['L1 Tag Requests Global']],r[ix['L2 Theoretical Sectors Global']],r[ix['L1 Wavefronts Shared']],r[ix['# Samples']],r[ix['Avg. Threads Executed']]))
print('--- top 25 instructions by samples')
for n,r in sorted(enumerate(data),key=lambda t:-int(t[1][ix['# Samples']]))[:25]:
    st={h[6:]:int(r[ix[h]]) for h in hdr if h.startswith('stall_') and 'Not' not in h and int(r[ix[h]])>0}
    top=sorted(st.items(),key=lambda kv:-kv[1])[:3]
    print('%4d %-50s samp %6s exec %9s %s'%(n,r[ix['Source']].strip()[:50],r[ix['# Samples']],r[ix['Instructions Executed']],top))
# stall totals
tot=collections.Counter()
for r in data:
    for h in hdr:
        if h.startswith('stall_') and 'Not' not in h: tot[h[6:]]+=int(r[ix[h]])
print('stall totals',tot.most_common(10))
# region split: by line number ranges of instruction index; print cumulative instr by index buckets of 50
b=collections.Counter(); bs=collections.Counter()
for n,r in enumerate(data):
    b[n//40]+=int(r[ix['Instructions Executed']]); bs[n//40]+=int(r[ix['# Samples']])
print('instr by 40-instr bucket',[ (k*40, round(v/tot_inst*100,1), round(bs[k]/tot_samp*100,1)) for k,v in sorted(b.items())])
