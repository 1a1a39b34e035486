"""Developer probe: what would taking the chroma copy out of the CLAHE kernel buy (device-resident 4K batch)?"""
import sys, os
sys.path.insert(0, "/root/repo")
import torch
from cuda import cudart
import opencv_opencl_b200 as nv
W, H, n = 3840, 2160, 256
pitch = nv.nv12_frame_bytes(W, H)
ctx = nv.Context(0, W, H, 1)
st = torch.cuda.current_stream()
d_in = torch.empty(n * pitch, dtype=torch.uint8, device="cuda"); d_out = torch.empty_like(d_in)
ctx.synth_nv12_device(d_in, n, pitch, W, H, stream=st)
aux = torch.cuda.Stream()
def run(mode, dma):
    for it in range(13):
        if it == 3:
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record(st)
        if dma:
            ev = torch.cuda.Event(); ev.record(st); aux.wait_event(ev)
            with torch.cuda.stream(aux):
                if dma == 1:
                    src = d_in.view(n, pitch)[:, W * H:]; dst = d_out.view(n, pitch)[:, W * H:]
                    dst.copy_(src)
                else:
                    err, = cudart.cudaMemcpy2DAsync(d_out.data_ptr() + W * H, pitch, d_in.data_ptr() + W * H, pitch, W * H // 2, n,
                                                    cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice, aux.cuda_stream)
                    assert err == cudart.cudaError_t.cudaSuccess, err
                ev2 = torch.cuda.Event(); ev2.record(aux)
        ctx.clahe_device(d_in, d_out, n, pitch, W, H, 2.0, (8, 8), uv_mode=mode, stream=st)
        if dma:
            st.wait_event(ev2)
    e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10 / n * 1e3
print("uv copy in kernel : %.2f us/frame" % run(nv.UV_COPY, False))
print("uv skipped        : %.2f us/frame" % run(nv.UV_SKIP, False))
print("uv by torch copy_ : %.2f us/frame" % run(nv.UV_SKIP, 1))
print("uv by cudaMemcpy2D: %.2f us/frame" % run(nv.UV_SKIP, 2))
print("uv copy in kernel : %.2f us/frame" % run(nv.UV_COPY, False))
