"""CPU, world_size 2 over gloo: frame sharding + the single flat all-reduce of the splat gradients.  The rasterizer in
these tests is the CPU oracle wrapped as an autograd.Function (tests may use the oracle; the product may not), so the
property checked is the host logic: N-rank sharded gradients == 1-rank accumulated gradients."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _scene():
    from tests import helpers as h

    return h.avatar_scene(P=600, W=64, H=48, seed=7, n_lat=6, n_lon=8, scale_gain=6.0)


def _frame_grads(sc, cams, flat_sink=None):
    """Accumulated raw-parameter gradients over `cams` through eager getters -> oracle rasterizer."""
    from gaussianavatars_b200 import GaussianRasterizationSettings
    from oracle import binding as ob
    from oracle import rasterizer as orc

    Fn = orc.make_autograd_function()
    p = sc["params"]
    names = ("_xyz", "_rotation", "_scaling", "_opacity", "_features_dc", "_features_rest")
    leaves = {k: p[k].clone().requires_grad_(True) for k in names}
    b = p["binding"].long()
    fr = ob.update_mesh_properties(sc["verts"], sc["faces"])
    for cam in cams:
        rs = GaussianRasterizationSettings(sc["H"], sc["W"], cam.tanfovx, cam.tanfovy, sc["bg"], 1.0,
                                           cam.world_view_transform, cam.full_proj_transform, 3, cam.camera_center,
                                           False, False)
        img, _ = Fn.apply(ob.get_xyz(leaves["_xyz"], b, fr["face_center"], fr["face_orien_mat"], fr["face_scaling"]),
                          torch.zeros(b.shape[0], 3), ob.get_features(leaves["_features_dc"], leaves["_features_rest"]).contiguous(),
                          None, ob.get_opacity(leaves["_opacity"]), ob.get_scaling(leaves["_scaling"], b, fr["face_scaling"]),
                          ob.get_rotation(leaves["_rotation"], b, fr["face_orien_quat"]), None, rs)
        img.square().sum().backward()
    return leaves


class _PC:
    def __init__(self, leaves):
        self.leaves = list(leaves.values())

    def parameters(self):
        return self.leaves


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gaussianavatars_b200 import dist as gdist
    from gaussianavatars_b200 import synthetic as syn

    torch.set_num_threads(1)
    sc = _scene()
    cams = [syn.orbit_camera(sc["W"], sc["H"], azimuth_deg=a) for a in (-30.0, -10.0, 10.0, 30.0)]
    mine = [cams[i] for i in gdist.shard_frames(len(cams), rank, world)]
    leaves = _frame_grads(sc, mine)
    pc = _PC(leaves)
    # (a) fallback path: separate .grad tensors -> one coalesced all-reduce
    n1 = gdist.allreduce_splat_grads(pc)
    # (b) flat path: make the grads alias one flat buffer as the fused backward does, all-reduce again (x world)
    flat = torch.cat([p.grad.reshape(-1) for p in pc.parameters()])
    off = 0
    for p in pc.parameters():
        n = p.grad.numel()
        p.grad = flat[off:off + n].view_as(p.grad)
        off += n
    pc.flat_grad = flat
    snapshot = flat.clone()
    n2 = gdist.allreduce_splat_grads(pc)
    assert n1 == 1 and n2 == 1
    assert torch.allclose(flat, snapshot * world)
    stats = [torch.full((4,), float(rank + 1)), torch.full((4,), 1.0), torch.full((4,), float(rank))]
    gdist.allreduce_densification_stats(*stats)
    assert stats[0][0] == sum(range(1, world + 1)) and stats[1][0] == world and stats[2][0] == world - 1
    if rank == 0:
        np.savez(os.path.join(out_dir, "sharded.npz"), **{f"g{i}": (snapshot[o:o + p.numel()]).numpy() for i, (p, o) in
                 enumerate(zip(pc.parameters(), np.cumsum([0] + [q.numel() for q in pc.parameters()][:-1])))})
    dist.destroy_process_group()


def test_frame_sharding_is_a_partition():
    from gaussianavatars_b200 import dist as gdist

    for world in (1, 2, 3, 8):
        seen = sorted(sum((gdist.shard_frames(17, r, world) for r in range(world)), []))
        assert seen == list(range(17))
    assert gdist.shard_frames(3, 5, 8) == []


def test_cost_sorted_dealing_is_a_partition_with_neighbouring_costs_per_step():
    from gaussianavatars_b200 import dist as gdist

    costs = [530, 910, 100, 740, 745, 300, 120, 905, 610, 615, 290, 101, 999, 500, 480, 770]
    for world in (1, 2, 4, 8):
        shards = [gdist.shard_frames_by_cost(costs, r, world) for r in range(world)]
        assert sorted(sum(shards, [])) == list(range(len(costs)))            # every camera exactly once
        assert len({len(s) for s in shards}) == 1
        order = sorted(range(len(costs)), key=lambda i: (costs[i], i))
        for j in range(len(costs) // world):                                  # step j = a run of the sorted order
            assert sorted(s[j] for s in shards) == sorted(order[j * world:(j + 1) * world])
    # ties are broken by index: deterministic on every rank
    assert gdist.shard_frames_by_cost([5, 5, 5, 5], 1, 2) == [1, 3]


def test_two_rank_sharded_grads_equal_single_rank_accumulation(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    sys.path.insert(0, ROOT)
    from gaussianavatars_b200 import synthetic as syn

    sc = _scene()
    cams = [syn.orbit_camera(sc["W"], sc["H"], azimuth_deg=a) for a in (-30.0, -10.0, 10.0, 30.0)]
    ref = _frame_grads(sc, cams)
    got = np.load(os.path.join(str(tmp_path), "sharded.npz"))
    for i, p in enumerate(ref.values()):
        a, b = got[f"g{i}"].reshape(-1), p.grad.numpy().reshape(-1)
        assert np.allclose(a, b, rtol=1e-4, atol=1e-6 * np.abs(b).max()), f"param {i}"


class _Model:
    """Six parameters laid out like the fused backward's flat buffer (widths 3 + 3 + 45 + 1 + 3 + 4 = 59)."""

    def __init__(self, P):
        self.ps = [torch.nn.Parameter(torch.zeros(P, w)) for w in (3, 3, 45, 1, 3, 4)]

    def parameters(self):
        return self.ps


def _fill(buf, pc, value):
    """What the fused backward does with a caller-owned buffer: store the step's local gradients, flag the model."""
    buf.flat.fill_(value)
    pc._gab200_mc_used = True


def _deferred_worker(rank, world, port):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gaussianavatars_b200 import dist as gdist

    pc = _Model(7)
    bufs = [gdist.SymmetricGradBuffer(pc, mode="plain") for _ in range(2)]
    assert all(b.enabled and b.numel == 7 * 59 for b in bufs)
    total = sum(range(1, world + 1))
    # synchronous protocol: begin / backward / end
    bufs[0].begin()
    _fill(bufs[0], pc, float(rank + 1))
    assert bufs[0].end() is True
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(pc.parameters(), bufs[0].all_views[0]))
    assert torch.all(pc.parameters()[2].grad == total)
    # deferred: step j fills buffer j % 2 and adopts the LOCAL gradients; the other buffer is reduced "meanwhile"
    seen = []
    for j in range(5):
        k = j % 2
        bufs[k].begin()
        _fill(bufs[k], pc, float((rank + 1) * 10 ** j))
        assert bufs[k].end(reduce=False) is True
        assert torch.all(pc.parameters()[0].grad == (rank + 1) * 10 ** j)      # local, unreduced
        if j > 0:
            bufs[1 - k].reduce()                                              # the previous step's gradients
            seen.append(float(bufs[1 - k].all_views[0][5][0, 0]))
    bufs[4 % 2].reduce()
    seen.append(float(bufs[0].all_views[0][5][0, 0]))
    assert seen == [float(total * 10 ** j) for j in range(5)], seen
    # a model whose parameters were replaced (densification): the buffer refuses, the fallback all-reduce takes over
    pc.ps[0] = torch.nn.Parameter(torch.zeros(7, 3))
    for p in pc.parameters():
        p.grad = torch.full_like(p, float(rank + 1))
    bufs[0].begin()
    pc._gab200_mc_used = True
    assert bufs[0].end() is False and not bufs[0].matches()
    assert torch.all(pc.parameters()[0].grad == total)
    dist.destroy_process_group()


def test_caller_owned_gradient_buffer_sync_and_deferred_reduction():
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_deferred_worker, args=(2, port), nprocs=2, join=True)
