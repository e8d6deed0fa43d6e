"""Source-level hot spots of one kernel from an `ncu --set full --import-source on` report: joins the report's per-SASS-address
sample counts with the file:line markers of `nvdisasm -g` output for the same .so (both compiled with -lineinfo).
usage: python tools/ncu_source_hotspots.py REPORT.ncu-rep KERNEL_REGEX SASS_FILE FUNC_KEY SRC_DIR/ [TOP]"""
import csv,collections,re,sys,subprocess
def analyze(rep, kregex, sassfile, func_key, srcdir, top=26):
    csvf='/tmp/_src.csv'
    subprocess.run("ncu -i %s --page source --csv --kernel-name regex:%s --launch-skip 0 --launch-count 1 2>/dev/null > %s"%(rep,kregex,csvf),shell=True)
    off2line={}; cur=None; infunc=False
    for ln in open(sassfile):
        if ln.startswith('.text.'):
            infunc = func_key in ln; continue
        if not infunc: continue
        m=re.search(r'//## File "(.*?)", line (\d+)',ln)
        if m: cur=(m.group(1).split('/')[-1],int(m.group(2))); continue
        m=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*);',ln)
        if m: off2line[int(m.group(1),16)]=cur
    rows=list(csv.reader(open(csvf)))
    hdr=rows[1]; idx={h:i for i,h in enumerate(hdr)}
    base=None; agg=collections.defaultdict(lambda:[0,0]); tot=0; tots=0
    stall_cols=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    st=collections.defaultdict(collections.Counter); allst=collections.Counter()
    for r in rows[2:]:
        if len(r)<len(hdr) or not r[0].startswith('0x'): continue
        a=int(r[idx['Address']],16)
        if base is None: base=a
        key=off2line.get(a-base,('?',0))
        ne=int(r[idx['Instructions Executed']]); ns=int(r[idx['# Samples']])
        agg[key][0]+=ne; agg[key][1]+=ns; tot+=ne; tots+=ns
        for c in stall_cols:
            try: st[key][c]+=int(r[idx[c]]); allst[c]+=int(r[idx[c]])
            except: pass
    print('== %s: total warp-inst %d, samples %d'%(func_key,tot,tots)); print(allst.most_common(8))
    for k,v in sorted(agg.items(),key=lambda kv:-kv[1][1])[:top]:
        try: src=open(srcdir+k[0]).read().split('\n')[k[1]-1].strip()[:78]
        except Exception: src=''
        tp=', '.join('%s=%d'%(c.replace('stall_',''),n) for c,n in st[k].most_common(2))
        print('%5d smp %4.1f%% inst %7d  %s:%d  %s   [%s]'%(v[1],100*v[1]/max(tots,1),v[0],k[0],k[1],src,tp))
if __name__=='__main__':
    analyze(*sys.argv[1:6], top=int(sys.argv[6]) if len(sys.argv)>6 else 26)
