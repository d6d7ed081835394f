"""Differential fuzz of the MIC kernel between two builds of the library (e.g. the tensor-memory build and
-DSELD_MIC_TMEM=0 from tools/build_variant.sh): random batch sizes, ragged lengths, short clips, silent channels, level
differences, frame capacities, both n_fft — outputs must be BIT-identical (the two builds run the same arithmetic).
    python tools/fuzz_mic_builds.py a=path/libseld_cuda.so b=path/libseld_cuda.so [seed] [iters]"""
import hashlib, os, subprocess, sys

CHILD = r"""
import sys, os, hashlib, importlib, numpy as np, torch
sys.path.insert(0, os.getcwd())
os.environ["SELD_CUDA_LIB"] = sys.argv[1]
sb = importlib.import_module("sound-event-localization-detection_b200")
seed, iters = int(sys.argv[2]), int(sys.argv[3])
rng = np.random.default_rng(seed)
for it in range(iters):
    n_fft = int(rng.choice([1024, 960]))
    B = int(rng.integers(1, 7))
    nmax = int(rng.integers(600, 80000))
    ns = [int(rng.integers(n_fft // 2 + 1, nmax + 1)) for _ in range(B)]
    ns[int(rng.integers(0, B))] = nmax
    x = torch.zeros((B, 4, nmax))
    for i, n in enumerate(ns):
        x[i, :, :n] = torch.from_numpy((rng.standard_normal((4, n)) * 10 ** rng.uniform(-4, 0, size=(4, 1))).astype(np.float32))
        if rng.random() < 0.3:
            x[i, int(rng.integers(0, 4))] = 0
        if rng.random() < 0.2:
            s0 = int(rng.integers(0, n)); x[i, :, s0:s0 + 3000] = 0
    lengths = torch.tensor(ns, dtype=torch.int64, device="cuda") if rng.random() < 0.7 else None
    plan = sb.features.get_plan(n_fft, 480, 64, 24000)
    T_full = 1 + nmax // 480
    T_out = int(rng.integers(1, T_full + 1)) if rng.random() < 0.4 else T_full
    out = plan.run(x.cuda(), mode="mic_gcc", lengths=lengths, T_out=T_out, check=False)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    print(it, n_fft, B, nmax, T_out, hashlib.sha1(out.cpu().numpy().tobytes()).hexdigest(), flush=True)
"""

def main():
    specs = [a for a in sys.argv[1:] if "=" in a]
    rest = [a for a in sys.argv[1:] if "=" not in a]
    seed, iters = (rest + ["5", "80"])[0], (rest + ["5", "80"])[1]
    outs = {}
    for spec in specs:
        name, path = spec.split("=", 1)
        r = subprocess.run([sys.executable, "-c", CHILD, os.path.abspath(path), seed, iters], capture_output=True, text=True)
        if r.returncode != 0:
            print(name, "FAILED\n", r.stderr[-2000:]); sys.exit(1)
        outs[name] = r.stdout.strip().splitlines()
    names = list(outs)
    bad = [(a, b) for a, b in zip(outs[names[0]], outs[names[1]]) if a != b]
    print(f"{len(outs[names[0]])} cases, {len(bad)} differ")
    for a, b in bad[:5]:
        print(" ", a, "|", b)
    sys.exit(1 if bad or len(outs[names[0]]) != int(iters) else 0)

if __name__ == "__main__":
    main()
