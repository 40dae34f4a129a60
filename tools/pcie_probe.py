"""Host <-> device copy times of the buffers pbf_step_host moves at 1 M particles (pinned memory, one stream): what
the e2e leg of bench.py cannot hide. B200 box of this pool: 12.6 MB in 0.23 ms either way (54.5 GB/s).

    python tools/pcie_probe.py
"""
import torch, time
n=1048576
h=torch.empty((n,3),dtype=torch.float32).pin_memory(); d=torch.empty((n,3),dtype=torch.float32,device="cuda")
h4=torch.empty(n,dtype=torch.int32).pin_memory(); d4=torch.empty(n,dtype=torch.int32,device="cuda")
st=torch.cuda.Stream()
with torch.cuda.stream(st):
    for name,fn in (("h2d 12.6MB", lambda: d.copy_(h,non_blocking=True)), ("d2h 12.6MB", lambda: h.copy_(d,non_blocking=True)), ("h2d 4.2MB", lambda: d4.copy_(h4,non_blocking=True)), ("d2h 3.1MB slice", lambda: h[:n//4].copy_(d[:n//4],non_blocking=True))):
        for _ in range(3): fn()
        st.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(20): fn()
        e1.record(st); st.synchronize()
        ms=e0.elapsed_time(e1)/20
        t0=time.perf_counter(); fn(); st.synchronize(); w=(time.perf_counter()-t0)*1e3
        print(name, "%.4f ms back-to-back, %.4f ms wall single" % (ms, w))
