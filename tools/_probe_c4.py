import os, sys, time, cProfile, pstats, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
frames = [bench._gen_frame(("c4", i)) for i in range(64)] * 2
import torch
from cm3d_b200.lifter import Lifter
lifter = Lifter("cuda:0")
for _ in lifter.lift_frame_stream(iter(frames), batch_frames=32, pack_workers=6): pass
torch.cuda.synchronize()
print("mallocs before", torch.cuda.memory_stats()["num_device_alloc"], flush=True)
pr = cProfile.Profile(); pr.enable()
t0 = time.perf_counter()
for _ in lifter.lift_frame_stream(iter(frames * 2), batch_frames=32, pack_workers=6): pass
torch.cuda.synchronize()
dt = time.perf_counter() - t0
pr.disable()
print("frames/s", 256 / dt, lifter.stream_stats, "mallocs after", torch.cuda.memory_stats()["num_device_alloc"], "frees", torch.cuda.memory_stats()["num_device_free"])
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14); print(s.getvalue())
