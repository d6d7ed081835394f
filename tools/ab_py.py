"""A/B of library builds through the Python binding, ONE process per build (tools/build_variant.sh makes them):
    python tools/ab_py.py name=path/to/libseld_cuda.so ...   [--clips 256] [--seconds 60] [--iters 10] [--passes 2]
Every child loads only the given library, times the FOA feature call (float32 / int16 PCM / normalised bf16 (B,C,T,F))
over the whole batch with CUDA events and prints min / median ms."""
import argparse, os, subprocess, sys

CHILD = r"""
import sys, os, statistics, importlib, torch
sys.path.insert(0, os.getcwd())
sb = importlib.import_module("sound-event-localization-detection_b200")
sb._lib.LIB_PATH = sys.argv[1]
B, sec, iters = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
g = torch.Generator(device="cuda").manual_seed(1)
x = 0.1 * torch.randn((B, 4, 24000 * sec), device="cuda", generator=g)
pcm = (x.clamp(-1, 1) * 32767).to(torch.int16)
mean = torch.zeros((7, 64), device="cuda"); istd = torch.ones((7, 64), device="cuda")
plan = sb.features.get_plan(1024, 480, 64, 24000)
T = plan.num_frames(x.shape[2])
o32 = torch.empty((B, T, 7, 64), device="cuda"); o16 = torch.empty((B, 7, T, 64), device="cuda", dtype=torch.bfloat16)
omic = torch.empty((B, T, 10, 64), device="cuda")
cases = {
  "mic f32": lambda: plan.run(x, mode="mic_gcc", out=omic),
  "foa f32": lambda: plan.run(x, mode="foa_iv", out=o32),
  "foa i16": lambda: plan.run(pcm, mode="foa_iv", out=o32),
  "foa bf16 ctf": lambda: plan.run(x, mode="foa_iv", mean=mean, inv_std=istd, layout="ctf", out_dtype=torch.bfloat16, out=o16),
}
outs = {}
for name, fn in cases.items():
    for _ in range(5): out = fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"{sys.argv[5]:>8s}: {name:14s} {min(ts):7.3f} / {statistics.median(ts):7.3f} ms   checksum {float(out.float().abs().sum()):.6e}", flush=True)
"""

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("libs", nargs="+")
    ap.add_argument("--clips", type=int, default=256)
    ap.add_argument("--seconds", type=int, default=60)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--passes", type=int, default=2)
    a = ap.parse_args()
    for _ in range(a.passes):
        for spec in a.libs:
            name, path = spec.split("=", 1)
            subprocess.run([sys.executable, "-c", CHILD, os.path.abspath(path), str(a.clips), str(a.seconds), str(a.iters), name], check=False)

if __name__ == "__main__":
    main()
