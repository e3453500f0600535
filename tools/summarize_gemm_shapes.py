import csv,re,sys
lines=[l for l in open(sys.argv[1]) if not l.startswith('==')]
rows=[r for r in csv.DictReader(lines) if r['Metric Name']=='gpu__time_duration.sum']
SH=[(8192,1280,1280,"plain"),(8192,1280,1280,"res"),(8192,1280,5120,"res"),(8192,3840,1280,"plain"),(8192,10240,1280,"geglu"),(32768,640,640,"plain"),(32768,640,640,"res"),(32768,640,2560,"res"),(32768,1920,640,"plain"),(32768,5120,640,"geglu"),(8192,2560,2048,"plain"),(8192,8192,8192,"plain")]
n=len(SH)
for i,s in enumerate(SH):
    ts=[float(rows[i+n*r]['Metric Value'].replace(',',''))/1e3 for r in range(2)]
    k=re.sub(r'\(.*','',rows[i]['Kernel Name'])
    M,N,K,_=s
    print(s, k[-22:], ["%.1f us"%t for t in ts], "%.0f TF/s"%(2*M*N*K/min(ts)/1e6))
