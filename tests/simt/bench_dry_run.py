"""TEST INFRASTRUCTURE: dry run of bench.py's product arm where there is no GPU.  The interpreter libraries stand in
for the CUDA ones and torch.cuda is faked (streams, events, pinned memory), so every line of the measurement logic
executes against the current C-ABI -- argument tables, launch counting, the per-kernel timers, the e2e loop through
crtx_frames_host, the JSON line.  The numbers it prints mean nothing.

    python tests/simt/bench_dry_run.py --batch 4 --steps 2 --warmup 1 --e2e-batch 8 --no-cpu-baseline [--variant pv1k]
"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "simt"))
os.chdir(ROOT)
import torch
import build as B
B.build()
import pkgload; pkgload.load()
from ntsc_crt_b200 import capi
capi.lib_path = B.lib_path
capi._libs.clear()

class FakeEvent:
    def __init__(self, enable_timing=False): pass
    def record(self, stream=None): pass
    def synchronize(self): pass
    def elapsed_time(self, other): return 1.0
    def query(self): return True
class FakeStream:
    cuda_stream = 0
    def __init__(self, *a, **k): pass
    def wait_event(self, e): pass
    def wait_stream(self, s): pass
    def synchronize(self): pass
    def __enter__(self): return self
    def __exit__(self, *a): return False
real_device = torch.device
def fake_device(*a, **k):
    return real_device("cpu")
torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda *a, **k: None
torch.cuda.current_stream = lambda *a, **k: FakeStream()
torch.cuda.Stream = FakeStream
torch.cuda.Event = FakeEvent
torch.cuda.synchronize = lambda *a, **k: None
torch.cuda.stream = lambda s: s
torch.cuda.get_device_name = lambda *a, **k: "SIMT interpreter"
torch.device = fake_device
torch.Tensor.pin_memory = lambda self, *a, **k: self
torch.Tensor.cuda = lambda self, *a, **k: self
import bench
bench.dropin_isolated = bench.dropin_fps  # (the child process would load the real CUDA library)
sys.argv = ["bench.py"] + sys.argv[1:]
bench.main()
