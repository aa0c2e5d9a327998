"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: index sharding and the differentiable all-to-all of
projected-splat rows that the Gaussian-sharded renderer is built on (b200gs/distributed.py; reference:
internal/renderers/gsplat_distributed_renderer.py:76-89,127-217).  The kernels themselves need a GPU
(tests/test_gpu_distributed.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from b200gs.distributed import ROW_FLOATS, exchange_rows, pack_rows, shard_range, unpack_rows
        g = torch.Generator().manual_seed(100 + rank)
        n = 50 + 7 * rank
        # rows destined to each peer: different counts per (src, dst)
        per_dest = []
        leaves = []
        for dst in range(world):
            v = 5 + 3 * rank + 2 * dst
            xys = torch.randn(v, 2, generator=g, requires_grad=True)
            radii = torch.randint(1, 50, (v,), generator=g, dtype=torch.int32)
            rows = pack_rows(xys, torch.rand(v, generator=g), torch.rand(v, 3, generator=g), torch.rand(v, generator=g),
                             torch.rand(v, 1, generator=g), torch.rand(v, 3, generator=g), radii, torch.ones(v, dtype=torch.bool))
            assert rows.shape == (v, ROW_FLOATS)
            per_dest.append(rows)
            leaves.append((xys, radii))
        got, counts = exchange_rows(per_dest)
        # what I must have received from src: 5 + 3*src + 2*rank rows
        assert counts == [5 + 3 * src + 2 * rank for src in range(world)]
        xys_r, depths_r, conics_r, comp_r, opac_r, rgbs_r, radii_r = unpack_rows(got)
        assert radii_r.dtype == torch.int32 and int(radii_r.min()) >= 1 and int(radii_r.max()) < 50
        # gradient: d/d(xys sent to dst) of sum(weights_at_dst * xys) — weight = (dst + 1)
        (xys_r * float(rank + 1)).sum().backward()
        for dst, (xys, _) in enumerate(leaves):
            assert torch.allclose(xys.grad, torch.full_like(xys, float(dst + 1)))
        # payload integrity: gather everything on rank 0 and compare with what was sent
        sent = [torch.cat([r.detach() for r in per_dest], 0)]
        allsent = [None] * world
        dist.all_gather_object(allsent, (sent[0], [int(r.shape[0]) for r in per_dest]))
        off = 0
        for src in range(world):
            rows_src, cnts = allsent[src]
            start = sum(cnts[:rank])
            exp = rows_src[start:start + cnts[rank]]
            assert torch.equal(got.detach()[off:off + cnts[rank]].view(torch.int32), exp.view(torch.int32))  # bit-exact incl. radius bits
            off += cnts[rank]
        # host-side camera exchange (gloo): every rank ends up with every rank's camera, in rank order, without device tensors
        from b200gs.distributed import VIEW_FLOATS, GatheredView, gather_views_host
        from b200gs.scene import make_ring_cameras
        cams = make_ring_cameras(64, 48)
        flat = gather_views_host(cams[3 + rank])
        assert flat.shape == (world, VIEW_FLOATS) and flat.device.type == "cpu"
        for src in range(world):
            gv = GatheredView(flat[src], "cpu")
            ref = cams[3 + src]
            assert (gv.width, gv.height) == (64, 48) and abs(gv.fx - float(ref.fx)) < 1e-4
            assert torch.equal(gv.world_to_camera, ref.world_to_camera.float()) and torch.equal(gv.camera_center_host, ref.camera_center.float())
        # sharding
        tot = 0
        for r in range(world):
            lo, hi = shard_range(1001, world, r)
            assert lo == tot
            tot = hi
        assert tot == 1001
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_exchange_rows_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}: {msg}"


class _FakeModel:
    """The slice of the reference's GaussianModel interface the renderer's sharding code touches (gaussian.py: properties dict,
    get_property_names / get_property, n_gaussians, get_xyz)."""

    def __init__(self, props):
        self.properties = props

    @property
    def n_gaussians(self):
        return self.properties["means"].shape[0]

    @property
    def get_xyz(self):
        return self.properties["means"]

    def get_property_names(self):
        return list(self.properties.keys())

    def get_property(self, name):
        return self.properties[name]


class _FakeModule:
    def __init__(self, model, optimizers, world, rank):
        import types
        self.gaussian_model, self.gaussian_optimizers = model, optimizers
        self.trainer = types.SimpleNamespace(world_size=world, global_rank=rank)
        self.density_changes = 0

    def density_updated_by_renderer(self):
        self.density_changes += 1


def _shard_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from b200gs.distributed import B200DistributedRenderer, B200DistributedRendererConfig, shard_range
        n = 101
        ids = torch.arange(n, dtype=torch.float32)
        props = {"means": torch.nn.Parameter(ids[:, None].repeat(1, 3).clone()), "opacities": torch.nn.Parameter(ids[:, None].clone()),
                 "frozen": torch.nn.Parameter(ids[:, None].clone(), requires_grad=False)}
        opts = [torch.optim.Adam([{"params": [props["means"]], "name": "means", "lr": 0.0}]),
                torch.optim.Adam([{"params": [props["opacities"]], "name": "opacities", "lr": 0.0}])]
        model = _FakeModel(dict(props))
        module = _FakeModule(model, opts, world, rank)
        r = B200DistributedRendererConfig(redistribute_interval=10, redistribute_until=100, redistribute_threshold=1.0).instantiate()
        # training_setup: contiguous index shards (gsplat_distributed_renderer.py:63-118), optimizers follow, module notified
        assert r.training_setup(module) == (None, None)
        lo, hi = shard_range(n, world, rank)
        assert model.n_gaussians == hi - lo and module.density_changes == 1
        assert torch.equal(model.properties["means"][:, 0].detach(), ids[lo:hi]) and torch.equal(model.properties["frozen"][:, 0], ids[lo:hi])
        assert opts[0].param_groups[0]["params"][0] is model.properties["means"] and model.properties["means"].requires_grad
        assert not model.properties["frozen"].requires_grad
        # give Adam a state whose rows are recognisable: exp_avg = 10 * id, exp_avg_sq = 100 * id
        for name, opt in zip(("means", "opacities"), opts):
            p_ = model.properties[name]
            p_.grad = torch.zeros_like(p_)
            opt.step()
            st = opt.state[p_]
            st["exp_avg"] = p_.detach() * 10
            st["exp_avg_sq"] = p_.detach() * 100
        torch.manual_seed(1234 + rank)
        r.after_training_step(5, module)                # not a multiple of the interval: nothing happens
        assert model.n_gaussians == hi - lo
        r.after_training_step(10, module)               # threshold 1.0: shards of 50 / 51 are unbalanced enough
        assert module.density_changes == 2
        got = model.properties["means"].detach()
        for name, opt in zip(("means", "opacities"), opts):
            p_ = opt.param_groups[0]["params"][0]
            assert p_ is model.properties[name] and p_.requires_grad
            st = opt.state[p_]
            assert torch.equal(st["exp_avg"], p_.detach() * 10) and torch.equal(st["exp_avg_sq"], p_.detach() * 100)   # rows moved together
        assert torch.equal(model.properties["opacities"].detach()[:, 0], got[:, 0]) and torch.equal(model.properties["frozen"][:, 0], got[:, 0])
        everyone = [None] * world
        dist.all_gather_object(everyone, got[:, 0].tolist())
        assert sorted(x for part in everyone for x in part) == ids.tolist()          # nothing lost, nothing duplicated
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_training_setup_sharding_and_redistribute_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shard_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}: {msg}"


def test_shard_range_matches_reference_rule():
    from b200gs.distributed import shard_range
    # gsplat_distributed_renderer.py:76-83: per = round(n / W); l = per * rank; r = l + per, last rank takes the remainder
    for n, w in ((10, 3), (1000000, 8), (3000000, 4), (7, 2)):
        per = round(n / w)
        for r in range(w):
            lo, hi = shard_range(n, w, r)
            assert lo == per * r and hi == (n if r == w - 1 else per * (r + 1))


def _board_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import time
        from b200gs import distributed as D
        board = D._board(None)
        assert board is not None, "ranks of one host with /dev/shm must get the shared-memory board"
        for step in range(300):
            if step % 37 == rank:            # skew: one rank arrives late
                time.sleep(0.002)
            mine = torch.arange(D.VIEW_FLOATS, dtype=torch.float32) + 1000.0 * step + 100000.0 * rank
            got = board.all_gather(mine)
            assert got.shape == (world, D.VIEW_FLOATS)
            for j in range(world):
                assert torch.equal(got[j], torch.arange(D.VIEW_FLOATS, dtype=torch.float32) + 1000.0 * step + 100000.0 * j), (step, j)
        # the public entry point takes the board, and gives what the gloo path gives
        from b200gs.scene import make_ring_cameras
        cams = make_ring_cameras(64, 48)
        a = D.gather_views_host(cams[2 + rank], None, cache=False)
        with D._STATE_LOCK:
            D._BOARDS[None] = None           # force the gloo path
        b = D.gather_views_host(cams[2 + rank], None, cache=False)
        assert torch.equal(a, b)
        q.put((rank, "ok"))
    except Exception:
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_shared_memory_board_world_3():
    """The host-side camera exchange of one-host groups (b200gs.distributed._ShmBoard): 300 steps with skewed arrival, 3 ranks."""
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_board_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}: {msg}"
