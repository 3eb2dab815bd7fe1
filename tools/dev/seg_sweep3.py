import sys; sys.path.insert(0,'/root/repo')
import torch, statistics
from ssim_b200 import api
lib=api.cuda_lib()
st=torch.cuda.current_stream(); sh=st.cuda_stream
def run(W,H,frames,with_map,seg,iters=10):
    a=torch.empty((frames,H,W),dtype=torch.uint8,device='cuda'); b=torch.empty_like(a)
    m=torch.empty((frames,H,W),dtype=torch.float32,device='cuda') if with_map else None
    sums=torch.empty(frames,dtype=torch.float64,device='cuda')
    for f in range(frames): api.synth_fill(0,sh,a[f].data_ptr(),W,b[f].data_ptr(),W,W,H,0,f)
    lib.ssim_cuda_set_segment_rows(seg)
    ts=[]
    for i in range(iters+3):
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(st)
        api.compute_device(0,sh,W,H,0,H,frames,a.data_ptr(),W,W*H,b.data_ptr(),W,W*H,m.data_ptr() if with_map else None,W,W*H,sums.data_ptr(),None)
        e1.record(st); torch.cuda.synchronize()
        if i>=3: ts.append(e0.elapsed_time(e1)*1e3)
    lib.ssim_cuda_set_segment_rows(0)
    t=statistics.median(ts); return round(t,1), round(W*H*frames/t)
for seg in [0,443,1821,911,683,1366]: print("16k^2 seg",seg,run(16384,16384,1,True,seg))
for seg in [0,1080,540,360,270]: print("512x1080p seg",seg,run(1920,1080,512,True,seg,6))
for seg in [0,1080,540,360,270]: print("64x1080p seg",seg,run(1920,1080,64,True,seg))
for seg in [0,1024,512,2048,683]: print("16384x2058 strip seg",seg,run(16384,2058,1,True,seg))
