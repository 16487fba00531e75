"""Host-link probe: `python profiles/link_probe.py GPU H2D_BYTES D2H_BYTES` - pinned cudaMemcpyAsync in both directions at once on one GPU;
run one per GPU at the same time to see what the box's host link gives each of them (does the upload's size matter to the download?)."""
import sys
import time
import torch
dev = int(sys.argv[1]); h2d = int(sys.argv[2]); d2h = int(sys.argv[3])
torch.cuda.set_device(dev)
hin = torch.empty(max(h2d, 1), dtype=torch.uint8, pin_memory=True); din = torch.empty(max(h2d, 1), dtype=torch.uint8, device='cuda')
hout = torch.empty(d2h, dtype=torch.uint8, pin_memory=True); dout = torch.empty(d2h, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(iters):
    for _ in range(iters):
        if h2d:
            with torch.cuda.stream(s1):
                din.copy_(hin, non_blocking=True)
        with torch.cuda.stream(s2):
            hout.copy_(dout, non_blocking=True)
    torch.cuda.synchronize()
run(20)
time.sleep(max(0.0, float(sys.argv[4]) - time.time())) if len(sys.argv) > 4 else None     # common start time
t0 = time.perf_counter(); run(400); dt = time.perf_counter() - t0
print(f'gpu {dev}: h2d {h2d} B  d2h {d2h} B: D2H {d2h * 400 / dt / 1e9:.2f} GB/s  H2D {h2d * 400 / dt / 1e9:.2f} GB/s')
