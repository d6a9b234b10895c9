import sys, os, time, cProfile, pstats, io
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
os.chdir(os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import torch, pkgload
pkgload.load()
from ntsc_crt_b200 import video
dev = torch.device("cuda", 0)
n = int(sys.argv[1]); S = int(sys.argv[2])
frames = torch.randint(0, 256, (n, 480, 640, 4), dtype=torch.uint8, device=dev)
conv = video.VideoConverter("ntsc", 640, 480, noise=0, scanlines=1, segments=S)
conv.convert(frames[:8]); torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
out = conv.convert(frames)
torch.cuda.synchronize()
pr.disable()
print("n", n, "S", S, "wall ms", 1e3 * (time.perf_counter() - t0))
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(25); print(s.getvalue()[:5000])
