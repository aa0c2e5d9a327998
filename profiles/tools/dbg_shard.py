"""2-GPU diagnostic (torchrun --nproc-per-node 2): where does the sharded image differ from the single-GPU image?
Captures the row buffer the sharded renderer rasterizes (monkeypatched ops.bin_and_blend_rows) and compares it, bit for bit, with the
visible rows of the single-view K1 on the full model; then compares images / final_T / n_contrib."""
import ctypes
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import b200gs  # noqa: F401
    from b200gs import ops
    from b200gs._lib import MODE_GSPLAT, check, lib, ptr
    from b200gs.distributed import B200DistributedRenderer, shard_range
    from b200gs.renderers import B200GSplatRenderer
    from b200gs.scene import SyntheticGaussians, make_ring_cameras, make_scene
    n, W, H = 20000, 400, 304
    raw = make_scene(n, 21, mean_scale=0.03)
    cams = make_ring_cameras(W, H)
    bg = torch.tensor([0.2, 0.1, 0.4], device=dev)
    lo, hi = shard_range(n, world, rank)
    full = SyntheticGaussians(raw).to(dev)
    shard = SyntheticGaussians({k: v[lo:hi] for k, v in raw.items()}).to(dev)
    single = B200GSplatRenderer(fused_activations=True).to(dev)

    captured = {}
    orig = ops.bin_and_blend_rows

    def spy(mode, width, height, rows, bg_, *a, **k):
        out = orig(mode, width, height, rows, bg_, *a, **k)
        captured["rows"] = rows.detach().clone()
        captured["args"] = (a, {kk: (vv.detach().clone() if torch.is_tensor(vv) else vv) for kk, vv in k.items()})
        captured["out"] = tuple(t.detach().clone() for t in out[1])
        return out

    ops.bin_and_blend_rows = spy
    rend = B200DistributedRenderer(fused=True).to(dev)
    for step in range(3):
        cam = cams[3 * rank + step].to_device(dev)
        with torch.no_grad():
            want = single(cam, full, bg)
        s_rows, s_out = captured["rows"], captured["out"]
        got = rend(cam, shard, bg)
        d_rows, d_out, d_args = captured["rows"], captured["out"], captured["args"]
        diff = (got["render"].detach() - want["render"]).abs()
        nbad = int((diff > 0).sum())
        print(f"[rank {rank} step {step}] image: {nbad} differing values, max {float(diff.max()):.3e}; rows single {tuple(s_rows.shape)} sharded {tuple(d_rows.shape)} "
              f"extra args {[(tuple(x.shape) if torch.is_tensor(x) else x) for x in d_args[0]]}", flush=True)
        vis = s_rows[:, 11].view(torch.int32) > 0
        sv = s_rows[vis]
        # valid rows of the sharded buffer, in order
        if len(d_args[0]) >= 4 and torch.is_tensor(d_args[0][2]):
            counts, cap = d_args[0][2].tolist(), int(d_args[0][3])
            dv = torch.cat([d_rows[b * cap:b * cap + c] for b, c in enumerate(counts)])
        else:
            dv = d_rows[d_rows[:, 11].view(torch.int32) > 0]
        print(f"[rank {rank} step {step}] visible rows: single {sv.shape[0]} sharded {dv.shape[0]}", flush=True)
        if sv.shape == dv.shape:
            for c in range(12):
                if c == 6:
                    continue
                ne = sv[:, c].view(torch.int32) != dv[:, c].view(torch.int32)
                if bool(ne.any()):
                    idx = ne.nonzero()[:4, 0].tolist()
                    print(f"[rank {rank} step {step}]   column {c}: {int(ne.sum())} rows differ, e.g. rows {idx}: single {sv[idx, c].tolist()} sharded {dv[idx, c].tolist()}", flush=True)
        for name, a, b in (("final_T", s_out[1], d_out[1]), ("n_contrib", s_out[2], d_out[2]), ("image", s_out[0], d_out[0])):
            ne = (a != b)
            print(f"[rank {rank} step {step}]   {name}: {int(ne.sum())} differ", flush=True)
        if nbad:
            ys, xs = (diff.sum(-1) > 0).nonzero()[:6].T.tolist() if diff.dim() == 3 else ([], [])
            print(f"[rank {rank} step {step}]   first differing pixels (y,x): {list(zip(ys, xs))}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
