import re,csv,collections,subprocess,sys,os
rep=sys.argv[1]
os.system(f"ncu -i {rep} --page source --csv 2>/dev/null > /tmp/src.csv; ncu -i {rep} --page raw --csv 2>/dev/null > /tmp/raw.csv")
os.system("cd /tmp && rm -rf xelf && mkdir xelf && cd xelf && cuobjdump -xelf all /root/repo/incubator_pegasus_b200/libpegasus_b200.so >/dev/null 2>&1 && nvdisasm -g -c compact.sm_100a.cubin > /tmp/merge_sass.txt")
rows=list(csv.reader(open('/tmp/raw.csv'))); d=dict(zip(rows[0],rows[2]))
for k in ['gpu__time_duration.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__warps_active.avg.pct_of_peak_sustained_active']:
    print(k, d.get(k))
lines=open('/tmp/merge_sass.txt').read().split('\n')
start=[i for i,l in enumerate(lines) if (l.startswith('_ZN3pgs7k_mergeILj1024ELb0') and l.rstrip().endswith(':'))][0]
end=len(lines)
for i in range(start+1,len(lines)):
    if lines[i].startswith('//--------------------- .text.'): end=i;break
insts=[]; cur=None
for l in lines[start:end]:
    m=re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)',l)
    if m:
        inl=re.findall(r'inlined at "([^"]+)", line (\d+)',m.group(3))
        top=(inl[-1][0],int(inl[-1][1])) if inl else (m.group(1),int(m.group(2)))
        cur=(m.group(1).split('/')[-1],int(m.group(2)),top[0].split('/')[-1],top[1]); continue
    if re.match(r'\s*/\*[0-9a-f]{4,6}\*/\s+\S',l): insts.append(cur)
rows=list(csv.reader(open('/tmp/src.csv'))); hdr=rows[1]; data=rows[2:]
print(len(insts),len(data))
ix=hdr.index('Instructions Executed'); isamp=hdr.index('# Samples'); ib=hdr.index('stall_barrier')
src=open('/root/repo/incubator_pegasus_b200/csrc/compact.cu').read().split('\n')
def find(s): return [i+1 for i,l in enumerate(src) if s in l][0]
marks=[('setup',find('for (;;) {')),('decode1_hdr',find('// ---- decode step 1')),('decode2_keys',find('// ---- decode step 2')),('valid_window',find('// ---- valid range of every run')),('rank',find('// ---- merge rank + shadow')),('filter',find('// ---- compaction filter + tombstone')),('survivors',find('// ---- survivors in merged order')),('lookback',find('// ---- decoupled look-back')),('write',find("// ---- write the tile's blocks")),('end',len(src)+1)]
def phase(l):
    if l<marks[0][1]: return 'helpers'
    for (n,a),(_,b) in zip(marks,marks[1:]):
        if a<=l<b: return n
    return 'other'
# walk instruction stream: attribute non-compact.cu lines to the phase of the most recent compact.cu line in stream order (approx)
ph=collections.Counter(); phs=collections.Counter(); phb=collections.Counter(); last='setup'
agg=collections.Counter(); samp=collections.Counter()
for c,dd in zip(insts,data):
    f,l=c[0],c[1]
    if f=='compact.cu' and l>=marks[0][1]: last=phase(l)
    ph[last]+=int(dd[ix]); phs[last]+=int(dd[isamp]); phb[last]+=int(dd[ib])
    agg[(f,l)]+=int(dd[ix]); samp[(f,l)]+=int(dd[isamp])
tot=sum(ph.values()); ts=sum(phs.values())
print('total inst',tot,'samples',ts)
for p,v in sorted(ph.items(), key=lambda kv:-kv[1]): print(f"{p:16s} inst {100*v/tot:5.1f}%  samples {100*phs[p]/ts:5.1f}%  (barrier {100*phb[p]/ts:5.1f}%)")
print()
for (f,l),v in sorted(samp.items(), key=lambda kv:-kv[1])[:22]:
    s=src[l-1].strip()[:100] if f=='compact.cu' else ''
    print(f"{f}:{l:4d} samp {100*v/ts:5.1f}% inst {100*agg[(f,l)]/tot:5.1f}% | {s}")
