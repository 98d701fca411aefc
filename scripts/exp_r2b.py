"""Round-2 experiment B (diagnostic): C2 step under the digit-plan / rank-mode / key-staging options, f64 sort, L2 fetch granularity."""
import ctypes, glob, os, sys
import torch
sys.path.insert(0, ".")
from datatable_b200 import engine, _lib

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
    for nm, ms in _lib.profile_records(): fam.setdefault(nm, []).append(ms)
    engine.set_option("profile", 0)
    return " ".join(f"{a}=" + "+".join(f"{x:.2f}" for x in b) for a, b in sorted(fam.items(), key=lambda t: -sum(t[1])))

ref = {}
def c2(check=False):
    h = engine.Groupby([k], [0], 1, reducers=[(_lib.OP_SUM, v)])
    if check:
        o, f, s = h.order(), h.offsets(), h.reduced(0)
        if "o" not in ref: ref.update(o=o, f=f, s=s)
        else:
            assert torch.equal(o, ref["o"]), "RowIndex differs between configurations"
            assert torch.equal(f, ref["f"]), "offsets differ between configurations"
            assert torch.allclose(s, ref["s"], rtol=1e-12), "sums differ"
    h.close()

for rb, rm, sk in ((8, 1, 1), (8, 1, 0), (8, 2, 0), (10, 0, 0), (0, 0, 0), (0, 0, 1), (9, 0, 0), (7, 2, 0)):
    engine.set_option("radix_bits", rb); engine.set_option("rank_mode", rm); engine.set_option("stage_keys", sk)
    c2(check=True)
    print(f"C2 radix_bits={rb} rank_mode={rm} stage_keys={sk}: {timed(c2):.2f} ms [{families(c2)}]", flush=True)
engine.set_option("radix_bits", 0); engine.set_option("rank_mode", 0); engine.set_option("stage_keys", 0)
del ref["o"], ref["f"], ref["s"]

# float64 sort (C3 shape at n/4) under the old and the new digit plan
m = n // 4
x = torch.randn(m, generator=g, device="cuda", dtype=torch.float64)
def so(): engine.Groupby([x], [4], 1).close()
for rb, rm in ((8, 1), (0, 0), (10, 0)):
    engine.set_option("radix_bits", rb); engine.set_option("rank_mode", rm)
    print(f"f64 sort n={m} radix_bits={rb} rank_mode={rm}: {timed(so):.2f} ms [{families(so)}]", flush=True)
engine.set_option("radix_bits", 0); engine.set_option("rank_mode", 0)
h = engine.Groupby([x], [4], 1); o = h.order(); h.close()
xs = x[o.long()]
assert bool((xs[1:] >= xs[:-1]).all()), "f64 sort: not sorted"
print("f64 sort check ok", flush=True)

# gather reducer through the RowIndex, with the device's L2 fetch granularity at its default and at 32 B
try:
    cands = glob.glob(os.path.join(os.path.dirname(os.path.dirname(torch.__file__)), "nvidia", "*", "lib", "libcudart.so*")) + glob.glob("/usr/local/cuda/lib64/libcudart.so.12")
    cudart = ctypes.CDLL(cands[0])
    lim = ctypes.c_size_t(0)
    cudart.cudaDeviceGetLimit(ctypes.byref(lim), 5); print("cudaLimitMaxL2FetchGranularity default:", lim.value)
    order, offsets, ng = engine.group([k[:m]], [0], 1)
    def red(): engine.reduce(_lib.OP_SUM, v[:m], order, offsets)
    def gat(): engine.gather(v[:m], order)
    for gran in (lim.value, 32, 64, 128):
        rc = cudart.cudaDeviceSetLimit(5, ctypes.c_size_t(gran))
        cudart.cudaDeviceGetLimit(ctypes.byref(lim), 5)
        print(f"L2 fetch granularity {gran} (rc {rc}, now {lim.value}): gather-reduce {timed(red):.2f} ms, gather {timed(gat):.2f} ms  (n={m})", flush=True)
except Exception as e:
    print("L2 granularity test failed:", e)
