"""CPU: structural invariants of the oracle (SURVEY.md 8c "independent correctness anchors") and the edge cases the
path has (empty input, everything culled, ragged image sizes, ties)."""
import numpy as np
import torch

from oracle import rasterizer as orc
from tests import helpers as h


def test_tile_bits_table():
    for n, bits in ((256, 9), (1785, 11), (3600, 12), (8160, 13), (16384, 15), (32400, 15), (1, 1), (2, 2)):
        assert orc.tile_bits(n) == bits


def test_keys_sorted_ranges_partition_and_stable_ties():
    scene = h.random_scene(3000, 200, 136, sh_degree=0, seed=1, scale_shift=0.8)
    # force exact depth ties: duplicate the first 500 splats' positions
    scene["means3D"][500:1000] = scene["means3D"][:500]
    st = h.oracle_forward(scene)
    k = st.keys_sorted
    assert (k[1:] >= k[:-1]).all()
    lens = (st.ranges[:, 1] - st.ranges[:, 0]).astype(np.int64)
    assert lens.sum() == st.N
    tiles = (k >> np.uint64(32)).astype(np.int64)
    for tile in np.unique(tiles)[:50]:
        a, b = st.ranges[tile]
        assert (tiles[a:b] == tile).all()
    # stable: among equal keys the splat ids ascend (emission order)
    same = k[1:] == k[:-1]
    assert same.any()
    assert (st.vals_sorted[1:][same] > st.vals_sorted[:-1][same]).all()
    # radii == 0 <=> no instance emitted
    emitted = np.zeros(3000, bool)
    emitted[st.vals_sorted] = True
    assert np.array_equal(emitted, st.radii > 0)
    assert st.tiles_touched.sum() == st.N and st.offsets[-1] == st.N


def test_weights_plus_final_transmittance_is_one_and_background_where_empty():
    scene = h.random_scene(800, 96, 80, sh_degree=0, seed=2, scale_shift=0.5)
    white = np.ones((800, 3), np.float32)
    a = h.oracle_forward(dict(scene, bg=torch.zeros(3)), shs=None, colors_precomp=white)
    b = h.oracle_forward(dict(scene, bg=torch.ones(3)), shs=None, colors_precomp=white)
    assert np.abs(b.out_color - 1.0).max() < 2e-6  # sum_i w_i + T_final = 1
    assert np.allclose(a.out_color[0], 1.0 - a.final_T, atol=2e-6)
    st = h.oracle_forward(scene)
    empty = st.n_contrib == 0
    for ch in range(3):
        assert (st.out_color[ch][empty & (st.final_T == 1.0)] == scene["bg"][ch].item()).all()


def test_permutation_invariance_with_distinct_depths():
    scene = h.random_scene(1500, 120, 88, sh_degree=1, seed=3, scale_shift=0.7)
    st = h.oracle_forward(scene)
    perm = torch.randperm(1500, generator=torch.Generator().manual_seed(0))
    sp = dict(scene)
    for k in ("means3D", "scales", "rotations", "opacities", "shs"):
        sp[k] = scene[k][perm].contiguous()
    st2 = h.oracle_forward(sp)
    assert np.array_equal(st2.radii, st.radii[perm.numpy()])
    assert len(np.unique(st.depths[st.radii > 0])) == (st.radii > 0).sum(), "fixture should have distinct depths"
    assert np.array_equal(st2.out_color, st.out_color)


def test_edge_cases_empty_culled_and_ragged():
    cam_scene = h.random_scene(10, 33, 17, sh_degree=0, seed=4)  # W,H not multiples of 16
    st = h.oracle_forward(cam_scene)
    assert st.out_color.shape == (3, 17, 33) and st.ranges.shape == (3 * 2, 2)
    # everything behind the near plane
    behind = dict(cam_scene)
    behind["means3D"] = cam_scene["means3D"].clone()
    behind["means3D"][:, 2] = 0.1
    sb = h.oracle_forward(behind)
    assert sb.N == 0 and (sb.radii == 0).all()
    for ch in range(3):
        assert (sb.out_color[ch] == cam_scene["bg"][ch].item()).all()
    g = h.oracle_backward(behind, sb, np.ones((3, 17, 33), np.float32))
    assert all(np.abs(v).sum() == 0 for k, v in g.items() if v is not None)
    # P = 0
    cam = cam_scene["cam"]
    s0 = orc.forward(np.zeros((0, 3), np.float32), np.zeros((0, 1), np.float32), cam.world_view_transform.numpy(),
                     cam.full_proj_transform.numpy(), cam.camera_center.numpy(), 33, 17, cam.tanfovx, cam.tanfovy,
                     np.zeros(3, np.float32), shs=np.zeros((0, 1, 3), np.float32), sh_degree=0,
                     scales=np.zeros((0, 3), np.float32), rotations=np.zeros((0, 4), np.float32))
    assert s0.N == 0 and (s0.out_color == 0).all()
    # an opacity below 1/255 can never contribute but keeps radii > 0 (drop-in: visibility_filter unchanged)
    faint = dict(cam_scene)
    faint["opacities"] = torch.full_like(cam_scene["opacities"], 0.003)
    sf = h.oracle_forward(faint)
    assert (sf.radii > 0).any() and (sf.n_contrib == 0).all()
