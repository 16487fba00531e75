"""Graph launch floor on this box: 32 trivial torch kernels per graph, time per kernel (CUDA events)."""
import torch
x = torch.zeros(1, device='cuda')
for _ in range(3):
    x.add_(1)
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    with torch.cuda.graph(g, stream=s):
        for _ in range(32):
            x.add_(1)
torch.cuda.current_stream().wait_stream(s)
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50):
    g.replay()
b.record()
torch.cuda.synchronize()
print(f'trivial kernel in a 32-node graph: {a.elapsed_time(b) * 1e3 / (50 * 32):.2f} us per node')
