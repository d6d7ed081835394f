"""Host<->device copy rates of the box, the ceiling of bench.py's end-to-end leg (no kernels involved):
    python tools/pcie_probe.py [--mb 184] [--chunks 16]
H2D alone, D2H alone, both directions at once (what the 3-stream pipeline of extract_features_host does), each from
ordinary pinned memory (torch pin_memory) and from write-combined pinned memory (cudaHostAlloc(..., WriteCombined): the
device's reads do not snoop the CPU caches).  Prints GB/s per direction.
Measured on a B200 box (PCIe 5 x16): 55.5 GB/s either direction alone; with both at once H2D falls to 49 GB/s (52 from a
write-combined buffer in this two-stream probe — but NOT in the real three-stream pipeline, where a write-combined
input buffer measured 257 k against 260 k clip-s/s, so the library does not offer one)."""
import argparse, ctypes, time
import torch


def host_alloc(nbytes, wc):
    rt = torch.cuda.cudart()
    flags = 4 if wc else 0  # cudaHostAllocWriteCombined = 0x04
    lib = ctypes.CDLL("libcudart.so.12")
    p = ctypes.c_void_p()
    rc = lib.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(flags))
    assert rc == 0, rc
    buf = (ctypes.c_byte * nbytes).from_address(p.value)
    return torch.frombuffer(buf, dtype=torch.uint8), p


def run(src_h, dst_d, src_d, dst_h, chunks, both):
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(chunks):
        if src_h is not None:
            with torch.cuda.stream(s1):
                dst_d[i % 2].copy_(src_h[i], non_blocking=True)
        if both or src_h is None:
            with torch.cuda.stream(s2):
                dst_h[i].copy_(src_d[i % 2], non_blocking=True)
    torch.cuda.synchronize()
    return time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=184)
    ap.add_argument("--chunks", type=int, default=16)
    a = ap.parse_args()
    n = a.mb << 20
    nd = n * 86 // 184  # the D2H : H2D byte ratio of the feature path fed with int16 PCM
    d_in = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(2)]
    d_out = [torch.zeros(nd, dtype=torch.uint8, device="cuda") for _ in range(2)]
    h_out = torch.empty((a.chunks, nd), dtype=torch.uint8, pin_memory=True)
    for wc in (False, True):
        if wc:
            flat, _keep = host_alloc(n * a.chunks, True)
            h_in = flat.view(a.chunks, n)
        else:
            h_in = torch.empty((a.chunks, n), dtype=torch.uint8, pin_memory=True)
        h_in.fill_(1)
        tag = "write-combined pinned" if wc else "pinned"
        for rep in range(2):
            t = run(h_in, d_in, None, None, a.chunks, False)
            t2 = run(None, None, d_out, h_out, a.chunks, False)
            t3 = run(h_in, d_in, d_out, h_out, a.chunks, True)
        print(f"{tag:22s}: H2D alone {n * a.chunks / t / 1e9:5.1f} GB/s | D2H alone {nd * a.chunks / t2 / 1e9:5.1f} GB/s | "
              f"both: H2D {n * a.chunks / t3 / 1e9:5.1f} + D2H {nd * a.chunks / t3 / 1e9:5.1f} GB/s", flush=True)


if __name__ == "__main__":
    main()
