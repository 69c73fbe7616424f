import csv,sys,collections,re
path=sys.argv[1]
rows=[];hdr=None
for r in csv.reader(open(path,errors='ignore')):
    if len(r)>5 and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        if d['Metric Name']=='gpu__time_duration.sum':
            v=float(d['Metric Value'].replace(',','')); u=d['Metric Unit']
            v=v/1e3 if u=='ns' else (v*1e3 if u=='ms' else v)
            rows.append((d['Kernel Name'],v))
idx=[i for i,(k,v) in enumerate(rows) if 'adam_kernel' in k]
iters=[];start=0
for j in range(3,len(idx),4):
    iters.append(rows[start:idx[j]+1]); start=idx[j]+1
last=iters[-2]
def grp(k):
    if 'conv_halo' in k: return 'ours: conv_halo (3x3 fwd/dgrad)'
    if 'tc_gemm' in k: return 'ours: tc_gemm (fc / 1x1 fwd/dgrad)'
    if 'tc_wgrad' in k or 'reduce_slabs' in k: return 'ours: tc_wgrad + slabs'
    if 'adam' in k: return 'ours: adam'
    if 'nms' in k: return 'ours: nms'
    if 'kmeans' in k: return 'ours: kmeans'
    if 'in_' in k and 'unnamed' in k: return 'ours: instance norm'
    if 'upsample2x' in k: return 'ours: upsample'
    if 'roi_pool' in k: return 'ours: roi pool'
    if 'unnamed>::' in k and 'native' not in k: return 'ours: other ('+k.split('::')[1][:20]+')'
    if 'cutlass' in k or 'cudnn' in k or 'convolve' in k or 'nhwcAddPadding' in k or 'engines_precompiled' in k or 'implicit' in k: return 'cuDNN/cuBLAS (GAN convs etc.)'
    if 'RadixSort' in k or 'radixSort' in k or 'sort' in k.lower(): return 'torch: sort'
    if 'reduce_kernel' in k: return 'torch: reduce'
    if 'elementwise' in k or 'fused_dropout' in k: return 'torch: elementwise'
    return 'torch: other'
agg=collections.Counter();cnt=collections.Counter()
for k,v in last:
    agg[grp(k)]+=v; cnt[grp(k)]+=1
tot=sum(agg.values())
print('total %.1f us, %d launches'%(tot,len(last)))
for k,v in agg.most_common(): print('%9.1f %5d %5.1f%% %s'%(v,cnt[k],100*v/tot,k))
if len(sys.argv)>2:
    pat=sys.argv[2]
    a2=collections.Counter();c2=collections.Counter()
    for k,v in last:
        if grp(k).startswith(pat): a2[k[:150]]+=v;c2[k[:150]]+=1
    for k,v in a2.most_common(40): print('   %8.1f %4d %s'%(v,c2[k],k))
