"""Host-side cost of one training step: (a) pure enqueue time with the GPU far behind, (b) cProfile of the step loop."""
import cProfile, pstats, sys, time, os
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
cot = (torch.rand(3, H, W) * 2 - 1).to(dev)

def step(i):
    for p in model.parameters():
        p.grad = None
    out = renderer(cams[i % len(cams)], model, bg)
    loss = (out["render"] * cot).sum()
    loss.backward()

for i in range(10):
    step(i)
torch.cuda.synchronize()
# (a) enqueue-only time: 40 steps back to back, host clock, then sync
t0 = time.perf_counter()
for i in range(40):
    step(i)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host loop {1e3*(t1-t0)/40:.3f} ms/step   incl. drain {1e3*(t2-t0)/40:.3f} ms/step")
# (b) where the host time goes
pr = cProfile.Profile()
pr.enable()
for i in range(60):
    step(i)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
