"""TEST HELPER (tests/test_bench_contract.py): execute bench.py's main() on a CPU-only box with torch.cuda and the
B200 package replaced by stand-ins, to check the HOST-SIDE control flow of the measured arm (argument handling, the
collective decision about the pinned staging buffers, the e2e block, JSON assembly) for Python-level errors and for
the keys of the contract line.  Nothing here measures anything: every number it prints is fake."""
import importlib, importlib.util, json, os, sys, types, ctypes
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
real_tensor, real_empty = torch.tensor, torch.empty
torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda d: None
torch.cuda.synchronize = lambda *a, **k: None
class Ev:
    def __init__(self, enable_timing=False): pass
    def record(self): pass
    def elapsed_time(self, other): return 274.6
torch.cuda.Event = Ev
torch.tensor = lambda *a, **k: real_tensor(*a, **{kk: vv for kk, vv in k.items() if kk != "device"})
FAIL_PIN = os.environ.get("MOCK_FAIL_PIN") == "1"
def empty(*a, **k):
    if k.pop("pin_memory", False) and FAIL_PIN:
        raise RuntimeError("mock: cannot pin")
    return real_empty(*[min(x, 1024) if isinstance(x, int) else x for x in a], **k)
torch.empty = empty

class FakeProb:
    def __init__(self, *a, **k): self.n = 0; self.t = 0.0
    def step(self, k): self.n += 93 * k; self.t += 1e-4 * k
    def stats(self): return {"rhs_evals": self.n, "kernel_launches": self.n // 4, "chain_launches": self.n // 4, "chain_stages": self.n, "t": self.t, "max_stages": 92}
    def get_state(self, h): pass
    def set_state(self, h, t): pass
    def run_batches(self, i, o, t, k): self.n += 94 * len(i)
    def close(self): pass
class FakeLib:
    def __init__(self): self.b = 0
    class F:
        def __init__(self, v): self.v = v; self.restype = None
        def __call__(self, *a): return self.v() if callable(self.v) else self.v
    def __getattr__(self, name):
        if name == "b200_algorithmic_bytes":
            st = {"v": 0}
            def f():
                st["v"] += 1533303324672
                return st["v"]
            fn = FakeLib.F(f)
        elif name == "b200_last_chain_kernel": fn = FakeLib.F(b"k_chain_march")
        else: fn = FakeLib.F(0)
        object.__setattr__(self, name, fn)
        return fn
fake = types.ModuleType("ceda-demonstrations_b200")
fake.Diffusion2D = FakeProb
lib = FakeLib()
fake.kernel_lib = lambda: lib
fake.nccl_unique_id = lambda: bytes(128)
sys.modules["ceda-demonstrations_b200"] = fake
spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
mod.run_reference_sample = lambda nsteps, n, ranks, base_n=16384: (1.1e9, 9.9, 93 * nsteps + 1)
sys.argv = ["bench.py"] + sys.argv[1:]
mod.main()
