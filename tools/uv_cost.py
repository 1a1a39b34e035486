import sys; sys.path.insert(0,".")
import torch, opencv_opencl_b200 as nv
W,H,n=3840,2160,256; pitch=nv.nv12_frame_bytes(W,H)
c=nv.Context(0,W,H,1); st=torch.cuda.current_stream()
a=torch.empty(n*pitch,dtype=torch.uint8,device="cuda"); b=torch.empty_like(a)
c.synth_nv12_device(a,n,pitch,W,H,stream=st)
for mode,name in ((0,"copy"),(2,"skip"),(1,"gray128")):
    for _ in range(3): c.clahe_device(a,b,n,pitch,W,H,2.0,(8,8),uv_mode=mode,stream=st)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(20): c.clahe_device(a,b,n,pitch,W,H,2.0,(8,8),uv_mode=mode,stream=st)
    e1.record(st); torch.cuda.synchronize(); ms=e0.elapsed_time(e1)/20
    print(f"uv {name}: {ms/n*1e3:.2f} us/frame")
