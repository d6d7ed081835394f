import importlib, torch, sys, os
sys.path.insert(0, os.getcwd())
sb = importlib.import_module("sound-event-localization-detection_b200")
B, T, G, M = 16, 250, 648, 14
g = torch.Generator(device="cuda").manual_seed(0)
logits = torch.randn((B, T, G, M), device="cuda", generator=g, requires_grad=True)
mask = torch.zeros((B, T, G), dtype=torch.int16, device="cuda")
mask.view(-1)[::97] = 5
for lt in ("mse", "ce"):
    crit = sb.CompactSMRSELDLoss(loss_type=lt, w_class=1.0, grid_size=(18, 36))
    for _ in range(3):
        loss, _ = crit(logits, mask); loss.backward(); logits.grad = None
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    n = 20; tf = tb = 0.0
    for _ in range(n):
        e[0].record(); loss, _ = crit(logits, mask); e[1].record(); loss.backward(); e[2].record(); torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2]); logits.grad = None
    print(f"compact {lt}: fwd {tf/n*1e3:.0f} us  bwd {tb/n*1e3:.0f} us   (ideal at 6.5 TB/s: fwd {logits.numel()*4/6.5e12*1e6:.0f} us, bwd {2*logits.numel()*4/6.5e12*1e6:.0f} us)")
# dense torch reference of the same loss for scale
y = torch.zeros((B, T, G, M), device="cuda"); y[..., M-1] = 1
for _ in range(2):
    l = ((torch.softmax(logits, -1) - y) ** 2).mean(); l.backward(); logits.grad = None
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    l = ((torch.softmax(logits, -1) - y) ** 2).mean(); l.backward(); logits.grad = None
e1.record(); torch.cuda.synchronize()
print(f"dense torch softmax-MSE fwd+bwd: {e0.elapsed_time(e1)/5*1e3:.0f} us")
