import json,sys,csv
for f in ('bench_c2','bench_c3'):
    try:
        t=open(f'gpurun_out/{f}.json').read().strip()
        d=json.loads(t)
        print(d['config']['workload'],'value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1),round(d['e2e']['ms_per_step'],4),'launches',d['gpu_launches'], 'build_ms', round(d['config']['index']['build_ms'],2))
        print('  roof',round(d['roofline']['kernel_ms'],4),round(d['roofline']['achieved'],1),round(d['roofline']['frac'],4),{k:round(v,4) for k,v in d['roofline']['other_kernels_ms'].items()})
        if d.get('cpu_baseline'): print('  cpu',round(d['cpu_baseline']['value'],2),d['cpu_baseline']['cores'])
    except Exception as e: print(f,'ERR',e, open(f'gpurun_out/{f}.err').read()[-500:])
rows=list(csv.reader(open('gpurun_out/launches_C2.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=i;break
hdr=rows[h]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
seq=[]
for r in rows[h+1:]:
    if len(r)<=vi: continue
    name=r[ki].split('(')[0].replace('void ','').replace('mp2p::<unnamed>::','').replace('mp2p::rs::','')
    seq.append((name,float(r[vi].replace(',','')),r[gi]))
n=int(sys.argv[1]) if len(sys.argv)>1 else 16
for nme,v,g in seq[-n-8:-8]: print('%-60s %8.2f us  grid %s'%(nme[:60],v/1000,g))
