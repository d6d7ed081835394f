import sys, time, os, tempfile, numpy as np, torch
sys.path.insert(0, "/root/repo")
import seld_b200 as sb
rng = np.random.default_rng(0)
rows = []
for src in range(4):
    az, el, cls = rng.integers(-180, 180), rng.integers(-60, 60), rng.integers(0, 13)
    for f in range(600):
        rows.append((f, cls, src, ((az + f // 10 + 180) % 360) - 180, el))
rows.sort()
path = os.path.join(tempfile.mkdtemp(), "m.csv")
open(path, "w").writelines(",".join(str(int(v)) for v in r) + "\n" for r in rows)
for fn, name in ((sb.metadata_to_labels, "metadata_to_labels"), (sb.augment_with_gaussian_noise, "augment_with_gaussian_noise")):
    np.random.seed(1)
    fn(path, 60.0, sample_rate=24000, I=18, J=36, cell_size_deg=10, num_classes=14)
    t0 = time.perf_counter()
    for _ in range(5):
        lab, I, J = fn(path, 60.0, sample_rate=24000, I=18, J=36, cell_size_deg=10, num_classes=14)
    dt = (time.perf_counter() - t0) / 5
    print(f"{name}: {len(rows)} rows, 60 s clip -> {tuple(lab.shape)} CPU tensor in {dt*1e3:.1f} ms (reference: 22.1 s / 17.3 s, BASELINE.md)")
