"""Diagnostic: where does the end-to-end Frame query on pinned host columns spend its time?"""
import sys, time
import torch
sys.path.insert(0, ".")
import datatable_b200 as dtb
from datatable_b200 import engine, _lib, f, by
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
g = torch.Generator(device="cuda"); g.manual_seed(42)
k = torch.randint(0, 1_000_000, (n,), generator=g, device="cuda", dtype=torch.int32)
v = torch.rand(n, generator=g, device="cuda", dtype=torch.float64)
kh = torch.empty(n, dtype=torch.int32, pin_memory=True); kh.copy_(k)
vh = torch.empty(n, dtype=torch.float64, pin_memory=True); vh.copy_(v)
del k, v
torch.cuda.synchronize()
DT = dtb.Frame(k=kh, v=vh)
def alloc_stats():
    st = torch.cuda.memory_stats()
    return st.get("num_device_alloc", 0), st.get("num_device_free", 0), st.get("num_alloc_retries", 0)
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 12):
    a0 = alloc_stats()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    R = DT[:, dtb.sum(f.v), by(f.k)]
    torch.cuda.synchronize(); t1 = time.perf_counter()
    a1 = alloc_stats()
    print(f"Frame API: {1e3*(t1-t0):.1f} ms   torch cudaMalloc/cudaFree/retries during the call: {a1[0]-a0[0]}/{a1[1]-a0[1]}/{a1[2]-a0[2]}"
          f"   reserved {torch.cuda.memory_reserved()/2**30:.1f} GiB", flush=True)
# manual pipeline with events
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    cs = torch.cuda.Stream()
    e_start = torch.cuda.Event(enable_timing=True); e_start.record()
    with torch.cuda.stream(cs):
        cs.wait_event(e_start)
        kd = kh.cuda(non_blocking=True); ek = torch.cuda.Event(enable_timing=True); ek.record(cs)
        vd = vh.cuda(non_blocking=True); ev = torch.cuda.Event(enable_timing=True); ev.record(cs)
    ms = torch.cuda.current_stream()
    ms.wait_event(ek)
    gb = engine.Groupby([kd], [0], 1)
    e_sorted = torch.cuda.Event(enable_timing=True); e_sorted.record()
    t_sorted_host = time.perf_counter()
    ms.wait_event(ev)
    s = gb.reduce(_lib.OP_SUM, vd)
    e_red = torch.cuda.Event(enable_timing=True); e_red.record()
    first = gb.first_rows(); keys = engine.gather(kd, first)
    out = (keys.cpu(), s.cpu())
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"manual: total {1e3*(t1-t0):.1f} ms | k uploaded {e_start.elapsed_time(ek):.1f} | v uploaded {e_start.elapsed_time(ev):.1f} | sorted {e_start.elapsed_time(e_sorted):.1f} (host returned at {1e3*(t_sorted_host-t0):.1f}) | reduced {e_start.elapsed_time(e_red):.1f}", flush=True)
    gb.close()
