"""Which part of the end-to-end loop costs time: H2D prefetch, D2H loss read, or the lagged host sync."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from b200gs.renderers import B200VanillaRenderer
from b200gs.scene import SyntheticGaussians, make_ring_cameras, make_scene
dev = torch.device("cuda", 0)
N, W, H = 1_000_000, 1920, 1080
model = SyntheticGaussians(make_scene(N, 0)).to(dev)
cams = [c.to_device(dev) for c in make_ring_cameras(W, H)]
renderer = B200VanillaRenderer().to(dev)
bg = torch.zeros(3, device=dev)
cot_host = (torch.rand(3, H, W) * 2 - 1).pin_memory()
cot = cot_host.to(dev)
copy_stream = torch.cuda.Stream(device=dev)
in_bufs = [torch.empty_like(cot), torch.empty_like(cot)]
in_ready = [torch.cuda.Event(), torch.cuda.Event()]
consumed = [torch.cuda.Event(), torch.cuda.Event()]
loss_hosts = [torch.zeros(1).pin_memory(), torch.zeros(1).pin_memory()]
loss_done = [torch.cuda.Event(), torch.cuda.Event()]

fwd_done = torch.cuda.Event()


def run(k, h2d, d2h, lag_sync, fine_dep, late=False):
    def prefetch(i):
        if fine_dep:
            copy_stream.wait_event(consumed[i & 1])
        else:
            copy_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(copy_stream):
            in_bufs[i & 1].copy_(cot_host, non_blocking=True)
            in_ready[i & 1].record(copy_stream)
    def step(i):
        for p in model.parameters():
            p.grad = None
        if h2d:
            torch.cuda.current_stream().wait_event(in_ready[i & 1]); c = in_bufs[i & 1]
        else:
            c = cot
        out = renderer(cams[i % len(cams)], model, bg)
        if h2d and late and i + 1 < k:      # upload the next image while THIS step's backward runs (compute-bound kernels)
            fwd_done.record()
            copy_stream.wait_event(fwd_done)
            with torch.cuda.stream(copy_stream):
                in_bufs[(i + 1) & 1].copy_(cot_host, non_blocking=True)
                in_ready[(i + 1) & 1].record(copy_stream)
        loss = (out["render"] * c).sum()
        loss.backward()
        consumed[i & 1].record()
        if d2h:
            loss_hosts[i & 1].copy_(loss.detach().reshape(1), non_blocking=True)
        loss_done[i & 1].record()
    torch.cuda.synchronize()
    for b in range(2): consumed[b].record()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if h2d: prefetch(0)
    for i in range(k):
        if h2d and not late and i + 1 < k: prefetch(i + 1)
        step(i)
        if lag_sync and i > 0: loss_done[(i - 1) & 1].synchronize()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k

for i in range(3): run(16, False, False, False, False)
for name, cfg in [("resident", (False, False, False, False)), ("resident+lagsync", (False, False, True, False)), ("d2h+lagsync", (False, True, True, False)),
                  ("h2d", (True, False, False, False)), ("h2d fine dep", (True, False, False, True)), ("full e2e", (True, True, True, False)),
                  ("full e2e fine dep", (True, True, True, True)), ("full e2e late upload", (True, True, True, False, True)),
                  ("h2d late upload", (True, False, False, False, True)), ("resident again", (False, False, False, False))]:
    print(f"{name:22s} {run(64, *cfg):.4f} ms/step")
