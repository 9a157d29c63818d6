import csv,sys,collections,re
rows=collections.OrderedDict()
lines=[l for l in open(sys.argv[1]) if not l.startswith('==')]
for r in csv.DictReader(lines):
    d=rows.setdefault(int(r['ID']),{'name':r['Kernel Name']}); d[r['Metric Name']]=float(r['Metric Value'].replace(',',''))
agg=collections.OrderedDict()
for k in sorted(rows):
    l=rows[k]; n=re.sub(r'\(.*','',l['name'].replace('ddk::','').replace('void ',''))
    if 'conv' in n or 'acc_tc' in n: print(k, n, round(l['gpu__time_duration.sum']/1e3,1),'us')
