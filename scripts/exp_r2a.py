"""Round-2 experiment A (diagnostic): rank micro-benchmark, reducer overlap, L2 fetch granularity on the gather reducer."""
import ctypes, glob, os, subprocess, sys
import torch
sys.path.insert(0, ".")
from datatable_b200 import engine, _lib

print(subprocess.run(["scripts/ubench/rank_bench"], capture_output=True, text=True).stdout, flush=True)

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
g = torch.Generator(device="cuda"); g.manual_seed(5)
k = torch.randint(0, 1_000_000, (n,), generator=g, device="cuda", dtype=torch.int32)
v = torch.rand(n, generator=g, device="cuda", dtype=torch.float64)

def timed(fn, reps=4, warm=2):
    for _ in range(warm): fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)

def families(fn):
    engine.set_option("profile", 1); _lib.lib.dtb_profile_reset(); fn()
    fam = {}
    for nm, ms in _lib.profile_records(): fam[nm] = fam.get(nm, 0.0) + ms
    engine.set_option("profile", 0)
    return " ".join(f"{a}={b:.2f}" for a, b in sorted(fam.items(), key=lambda t: -t[1]))

def c2():
    h = engine.Groupby([k], [0], 1, reducers=[(_lib.OP_SUM, v)]); h.close()

for ov in (0, 1):
    engine.set_option("overlap_reducers", ov)
    print(f"C2 overlap_reducers={ov}: {timed(c2):.2f} ms [{families(c2)}]", flush=True)
engine.set_option("overlap_reducers", 0)

# gather reducer through the RowIndex, with the device's L2 fetch granularity at its default and at 32 B
cudart = ctypes.CDLL(glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*"))[0])
lim = ctypes.c_size_t(0)
cudart.cudaDeviceGetLimit(ctypes.byref(lim), 5); print("cudaLimitMaxL2FetchGranularity default:", lim.value)
m = n // 4
order, offsets, ng = engine.group([k[:m]], [0], 1)
def red(): engine.reduce(_lib.OP_SUM, v[:m], order, offsets)
def gat(): engine.gather(v[:m], order)
for gran in (lim.value, 32, 64, 128):
    rc = cudart.cudaDeviceSetLimit(5, ctypes.c_size_t(gran))
    cudart.cudaDeviceGetLimit(ctypes.byref(lim), 5)
    print(f"L2 fetch granularity {gran} (rc {rc}, now {lim.value}): gather-reduce {timed(red):.2f} ms, gather {timed(gat):.2f} ms  (n={m})", flush=True)
