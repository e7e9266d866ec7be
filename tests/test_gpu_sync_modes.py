"""-m gpu: how the forward learns its instance count N (include/gab200_rasterizer.h gab200_sync_mode) must never change a
result -- a wait in the middle of the frame (EXACT), a check at the end with transparent re-run on overflow (LATE), or
no wait at all inside a captured CUDA graph (NONE, graph.py) with detection + re-capture."""
import numpy as np
import pytest
import torch

from tests import helpers as h

pytestmark = pytest.mark.gpu


class Pipe:
    debug = False
    compute_cov3D_python = False
    convert_SHs_python = False


def _model(sc, dev, requires_grad=True):
    from gaussianavatars_b200 import synthetic as syn
    from gaussianavatars_b200.model import MeshBoundGaussians

    return MeshBoundGaussians(sc["params"], 3, sc["verts"], sc["faces"], pose_fn=syn.pose_mesh, device=dev,
                              requires_grad=requires_grad)


def _frame(pc, sc, dev, gout, verts=None):
    from gaussianavatars_b200.renderer import render

    for p in pc.parameters():
        p.grad = None
    v = (sc["verts"] if verts is None else verts).to(dev).clone().requires_grad_(True)
    pc.update_mesh_properties(v)
    out = render(sc["cam"].to(dev), pc, Pipe, sc["bg"].to(dev))
    out["render"].backward(gout)
    torch.cuda.synchronize()
    grads = [p.grad.clone() for p in pc.parameters()] + [v.grad.clone(), out["viewspace_points"].grad.clone()]
    return out["render"].detach().clone(), out["radii"].clone(), grads


def _grads_close(a, b):
    for x, y in zip(a, b):
        scale = float(y.abs().max()) + 1e-30
        assert float((x - y).abs().max()) <= 2e-5 * scale, "gradients differ beyond atomic-order noise"


def test_late_check_equals_mid_frame_sync_and_reruns_on_overflow():
    from gaussianavatars_b200 import rasterizer as R

    dev = torch.device("cuda:0")
    sc = h.avatar_scene(P=20_000, W=480, H=352, seed=3)
    gout = torch.randn(3, sc["H"], sc["W"], generator=torch.Generator().manual_seed(2)).to(dev)
    R.set_exact_binning(False)
    try:
        R.set_sync_policy("exact")
        pc = _model(sc, dev)
        img0, radii0, g0 = _frame(pc, sc, dev, gout)
        assert R.last_frame_info()["sync_mode"] == 0
        img0b, _, _ = _frame(pc, sc, dev, gout)   # second frame: bucket depth sort, still mid-frame sync
        assert R.last_frame_info()["sync_mode"] == 0 and torch.equal(img0, img0b)
        n = R.last_frame_info()["num_rendered"]

        R.set_sync_policy("late")
        pc = _model(sc, dev)
        _frame(pc, sc, dev, gout)                 # first frame of a model has no capacity hint -> EXACT
        assert R.last_frame_info()["sync_mode"] == 0
        img1, radii1, g1 = _frame(pc, sc, dev, gout)
        info = R.last_frame_info()
        assert info["sync_mode"] == 1 and info["attempts"] == 1 and info["num_rendered"] == n and info["capacity"] > n
        assert torch.equal(img0, img1) and torch.equal(radii0, radii1)
        _grads_close(g1, g0)

        # capacity far too small: the speculative pass is truncated, the end-of-call check re-enqueues binning + blend
        key = (dev, sc["W"], sc["H"], 20_000)
        hints = R.hints_of(pc)
        hints.set_capacity(key, 1000)
        img2, radii2, g2 = _frame(pc, sc, dev, gout)
        info = R.last_frame_info()
        assert info["sync_mode"] == 1 and info["attempts"] == 2 and info["num_rendered"] == n
        assert torch.equal(img0, img2) and torch.equal(radii0, radii2)
        _grads_close(g2, g0)

        # depth hint that fits nothing (one end bucket holds every splat) AND a tiny capacity
        hints.set_capacity(key, 777)
        hints.set_depth(key, (1, 2))
        img3, _, g3 = _frame(pc, sc, dev, gout)
        info = R.last_frame_info()
        assert info["depth_sort_path"] == 2 and info["attempts"] >= 2
        assert torch.equal(img0, img3)
        _grads_close(g3, g0)
        # and the frame after it is back on the fast path with fresh hints
        img4, _, _ = _frame(pc, sc, dev, gout)
        info = R.last_frame_info()
        assert info["attempts"] == 1 and info["depth_sort_path"] == 1 and torch.equal(img0, img4)
    finally:
        R.set_sync_policy("late")


@pytest.mark.parametrize("loss,host_inputs", [("dL_dimage", False), ("l1_u8", True), ("photometric", False)])
def test_graphed_frame_equals_eager_frame(loss, host_inputs):
    import gaussianavatars_b200 as g
    from gaussianavatars_b200 import synthetic as syn
    from gaussianavatars_b200.graph import GraphedFrame, camera_block
    from gaussianavatars_b200.renderer import render

    dev = torch.device("cuda:0")
    sc = h.avatar_scene(P=15_000, W=400, H=304, seed=5)
    cam = sc["cam"]
    gen = torch.Generator().manual_seed(7)
    gout = torch.randn(3, sc["H"], sc["W"], generator=gen).to(dev)
    gt = torch.randint(0, 256, (3, sc["H"], sc["W"]), generator=gen, dtype=torch.uint8)
    verts2 = syn.pose_mesh(sc["verts"], 9)

    def eager(pc, verts):
        for p in pc.parameters():
            p.grad = None
        v = verts.to(dev).clone().requires_grad_(True)
        pc.update_mesh_properties(v)
        out = render(cam.to(dev), pc, Pipe, sc["bg"].to(dev))
        if loss == "dL_dimage":
            out["render"].backward(gout)
            lv = None
        elif loss == "l1_u8":
            lv = g.l1_loss_u8(out["render"], gt.to(dev))
            lv.backward()
        else:
            lv = g.photometric_loss(out["render"], gt.to(dev), 0.2)
            lv.backward()
        torch.cuda.synchronize()
        return (out["render"].detach().clone(), [p.grad.clone() for p in pc.parameters()] + [v.grad.clone()],
                None if lv is None else float(lv))

    pc_e = _model(sc, dev)
    ref1 = eager(pc_e, sc["verts"])
    ref2 = eager(pc_e, verts2)

    pc = _model(sc, dev)
    fr = GraphedFrame(pc, sc["W"], sc["H"], cam.FoVx, cam.FoVy, sc["bg"], loss=loss, host_inputs=host_inputs)
    fr.set_inputs(camera=camera_block(cam), verts=sc["verts"].to(dev), gt_u8=None if loss == "dL_dimage" else gt,
                  dL_dimage=gout if loss == "dL_dimage" else None)
    fr.capture()
    for verts, ref in ((sc["verts"], ref1), (verts2, ref2), (sc["verts"], ref1)):
        fr.set_inputs(verts=verts.to(dev))
        fr.run(check=True)
        torch.cuda.synchronize()
        assert fr.captures == 1, "the frame overflowed a capacity sized from its own warm-up"
        assert torch.equal(fr.image, ref[0]), "graph replay image differs from the eager frame"
        _grads_close([p.grad for p in pc.parameters()] + [fr.verts.grad], ref[1])
        if ref[2] is not None:
            assert abs(float(fr.loss_host) - ref[2]) <= 1e-6 * max(1.0, abs(ref[2]))
    assert fr.replays == 3


def test_graphed_frame_detects_overflow_and_regrows():
    from gaussianavatars_b200.graph import GraphedFrame, camera_block

    dev = torch.device("cuda:0")
    sc = h.avatar_scene(P=15_000, W=400, H=304, seed=6)
    cam = sc["cam"]
    gout = torch.randn(3, sc["H"], sc["W"], generator=torch.Generator().manual_seed(3)).to(dev)
    pc_e = _model(sc, dev)
    img_ref, _, g_ref = _frame(pc_e, sc, dev, gout)

    pc = _model(sc, dev)
    fr = GraphedFrame(pc, sc["W"], sc["H"], cam.FoVx, cam.FoVy, sc["bg"], loss="dL_dimage")
    fr.set_inputs(camera=camera_block(cam), verts=sc["verts"].to(dev), dL_dimage=gout)
    fr.capture(capacity=4096)            # far below the ~1e5 instances this frame needs
    fr.run(check=False)
    assert fr.overflowed(wait=True), "an overflowing replay was not flagged"
    c = fr.counters()
    assert c["num_rendered"] > c["capacity"] == 4096
    assert not torch.equal(fr.image, img_ref)      # truncated list: wrong picture, but no fault
    fr.run(check=True)                             # waits, sees the flag, re-captures with room, replays
    assert fr.captures == 2 and not fr.overflowed(wait=True)
    assert torch.equal(fr.image, img_ref)
    _grads_close([p.grad for p in pc.parameters()] + [fr.verts.grad, fr.viewspace_points.grad], g_ref)


def test_two_frames_prefetching_each_others_host_inputs_inside_their_graphs():
    """The e2e loader pattern of bench.py: frame A's graph uploads frame B's staged camera + ground truth on a forked
    branch while it computes, and vice versa.  Every step must equal the eager step on the same inputs."""
    import gaussianavatars_b200 as g
    from gaussianavatars_b200 import synthetic as syn
    from gaussianavatars_b200.graph import GraphedFrame, camera_block
    from gaussianavatars_b200.renderer import render

    dev = torch.device("cuda:0")
    sc = h.avatar_scene(P=10_000, W=320, H=240, seed=9)
    cams = [syn.orbit_camera(sc["W"], sc["H"], r=1.0, fovy_deg=20.0, azimuth_deg=a) for a in (-20.0, 5.0, 30.0)]
    gen = torch.Generator().manual_seed(11)
    gts = [torch.randint(0, 256, (3, sc["H"], sc["W"]), generator=gen, dtype=torch.uint8) for _ in range(3)]
    pc = _model(sc, dev)
    frames = []
    for k in range(2):
        f = GraphedFrame(pc, sc["W"], sc["H"], cams[0].FoVx, cams[0].FoVy, sc["bg"], loss="l1_u8", host_inputs=True,
                         warm_cameras=[camera_block(c) for c in cams])
        f.cam_stage.copy_(camera_block(cams[0]))
        f.gt_stage.copy_(gts[0])
        f.set_inputs(verts=sc["verts"].to(dev))
        f.upload_staged()
        frames.append(f)
    frames[0].prefetch_for(frames[1])
    frames[1].prefetch_for(frames[0])
    for f in frames:
        f.capture()
    pc_e = _model(sc, dev)
    frames[0].cam_stage.copy_(camera_block(cams[0]))
    frames[0].gt_stage.copy_(gts[0])
    frames[0].upload_staged()
    for i in range(5):
        cur, nxt = frames[i % 2], frames[(i + 1) % 2]
        torch.cuda.synchronize()                       # the loader may only refill a staging buffer nobody is reading
        nxt.cam_stage.copy_(camera_block(cams[(i + 1) % 3]))
        nxt.gt_stage.copy_(gts[(i + 1) % 3])
        cur.run(check=True)
        torch.cuda.synchronize()
        for p in pc_e.parameters():
            p.grad = None
        pc_e.update_mesh_properties(sc["verts"].to(dev))
        out = render(cams[i % 3].to(dev), pc_e, Pipe, sc["bg"].to(dev))
        loss = g.l1_loss_u8(out["render"], gts[i % 3].to(dev))
        loss.backward()
        assert torch.equal(cur.image, out["render"].detach()), f"step {i}: image differs"
        assert abs(float(cur.loss_host) - float(loss)) <= 1e-6
        _grads_close([p.grad for p in pc.parameters()], [p.grad for p in pc_e.parameters()])
