import subprocess, sys, os
CHILD = r"""
import sys, os, statistics, importlib, torch
sys.path.insert(0, os.getcwd())
os.environ["SELD_CUDA_LIB"] = sys.argv[1]
sb = importlib.import_module("sound-event-localization-detection_b200")
B = 256
g = torch.Generator(device="cuda").manual_seed(1)
x = 0.1 * torch.randn((B, 4, 24000 * 60), device="cuda", generator=g)
for nfft in (1024, 960):
    plan = sb.features.get_plan(nfft, 480, 64, 24000)
    T = plan.num_frames(x.shape[2])
    o = torch.empty((B, T, 10, 64), device="cuda")
    for _ in range(4): plan.run(x, mode="mic_gcc", out=o)
    torch.cuda.synchronize(); ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); plan.run(x, mode="mic_gcc", out=o); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"{sys.argv[2]:>8s}: mic n_fft {nfft}: {min(ts):7.3f} / {statistics.median(ts):7.3f} ms  checksum {float(o.abs().sum()):.6e}", flush=True)
"""
for _ in range(2):
    for spec in sys.argv[1:]:
        name, path = spec.split("=", 1)
        subprocess.run([sys.executable, "-c", CHILD, os.path.abspath(path), name])
