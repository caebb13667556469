"""Host-side logic that needs no GPU: camera conventions, scene generator, operator argument validation,
window sharding helpers, oracle edge cases."""
import hashlib
import math

import numpy as np
import pytest
import torch

import gs_livm_b200  # noqa: F401
from gs_livm_b200 import rasterizer, scene, window

import helpers


def test_camera_conventions_match_gs_livm():
    """viewmatrix = Tcw^T, projmatrix = (P Tcw)^T, camera_center = Twc translation (src/gs/camera.cu:36-48)."""
    cam = scene.track_camera(2)
    V = cam.viewmatrix.numpy()
    Tcw = V.T
    assert np.allclose(Tcw[3], [0, 0, 0, 1])
    R = Tcw[:3, :3]
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-6)
    c = cam.camera_center.numpy()
    assert np.allclose(Tcw @ np.append(c, 1.0), [0, 0, 0, 1], atol=1e-5)
    # a point 10 m in front of the camera projects to the image centre with w = depth
    fwd_world = R.T @ np.array([0, 0, 10.0]) + c
    hom = np.append(fwd_world, 1.0) @ cam.projmatrix.numpy()
    assert abs(hom[3] - 10.0) < 1e-4 and abs(hom[0]) < 1e-4 and abs(hom[1]) < 1e-4
    assert math.isclose(cam.tanfovx, 752 / (2 * 425.0259), rel_tol=1e-6)
    assert math.isclose(cam.tanfovy, 480 / (2 * 426.7976), rel_tol=1e-6)


def test_projection_matrix_is_fov_only():
    P = scene.projection_matrix(0.01, 100.0, 1.0, 0.7)
    assert P[0, 2] == 0 and P[1, 2] == 0 and P[3, 2] == 1
    assert math.isclose(P[0, 0], 1 / math.tan(0.5), rel_tol=1e-6)
    assert math.isclose(P[2, 2], 100.0 / (100.0 - 0.01), rel_tol=1e-6)


def test_scene_generator_is_deterministic_and_shaped():
    a = scene.make_scene(5000, 123, 3)
    b = scene.make_scene(5000, 123, 3)
    c = scene.make_scene(5000, 124, 3)
    for k in ("means3D", "scales", "rotations", "opacities", "shs"):
        assert torch.equal(a[k], b[k])
    assert not torch.equal(a["means3D"], c["means3D"])
    assert a["means3D"].shape == (5000, 3) and a["shs"].shape == (5000, 16, 3) and a["opacities"].shape == (5000, 1)
    assert torch.allclose(a["rotations"].norm(dim=1), torch.ones(5000), atol=1e-5)
    assert (a["scales"] > 0.3).any(), "some Gaussians must trip the reference's scale cull"
    assert (a["means3D"][:, 0] < 0).any(), "some Gaussians must sit behind the first camera"
    assert a["means3D"][:, 0].max() < 75.0, "the map must stay within LiDAR range of the track"
    h = hashlib.sha256(a["means3D"].numpy().tobytes()).hexdigest()
    assert h == hashlib.sha256(b["means3D"].numpy().tobytes()).hexdigest()


def test_rasterizer_argument_validation_mirrors_reference():
    """std::invalid_argument in the reference (src/gs/rasterizer.cu:161-169) -> ValueError here, before any GPU use."""
    cam = scene.track_camera(0)
    rs = scene.settings_for(cam, torch.ones(3), 3)
    r = rasterizer.GaussianRasterizer(rs)
    m = torch.zeros(4, 3)
    with pytest.raises(ValueError, match="exactly one of either SHs or precomputed colors"):
        r(m, m, torch.zeros(4, 1), shs=None, colors_precomp=None, scales=m, rotations=torch.zeros(4, 4))
    with pytest.raises(ValueError, match="exactly one of either SHs or precomputed colors"):
        r(m, m, torch.zeros(4, 1), shs=torch.zeros(4, 16, 3), colors_precomp=m, scales=m, rotations=torch.zeros(4, 4))
    with pytest.raises(ValueError, match="scale/rotation pair or precomputed 3D covariance"):
        r(m, m, torch.zeros(4, 1), shs=torch.zeros(4, 16, 3))
    with pytest.raises(ValueError, match="scale/rotation pair or precomputed 3D covariance"):
        r(m, m, torch.zeros(4, 1), shs=torch.zeros(4, 16, 3), scales=m, rotations=torch.zeros(4, 4),
          cov3D_precomp=torch.zeros(4, 6))


def test_cpu_tensors_are_rejected_not_silently_computed():
    with pytest.raises(RuntimeError, match="no CPU path"):
        rasterizer.rasterize_gaussians_cuda(torch.ones(3), torch.zeros(4, 3), torch.empty(0), torch.zeros(4, 1),
                                            torch.zeros(4, 3), torch.zeros(4, 4), 1.0, torch.empty(0), torch.eye(4),
                                            torch.eye(4), 1.0, 1.0, 16, 16, torch.zeros(4, 16, 3), 3, torch.zeros(3),
                                            False, False)
    with pytest.raises(RuntimeError, match="num_points, 3"):
        rasterizer.rasterize_gaussians_cuda(torch.ones(3), torch.zeros(4, 2), torch.empty(0), torch.zeros(4, 1),
                                            torch.zeros(4, 3), torch.zeros(4, 4), 1.0, torch.empty(0), torch.eye(4),
                                            torch.eye(4), 1.0, 1.0, 16, 16, torch.zeros(4, 16, 3), 3, torch.zeros(3),
                                            False, False)


def test_views_for_rank_partitions_the_window():
    for world in (1, 2, 3, 4, 8):
        seen = sorted(v for r in range(world) for v in window.views_for_rank(8, r, world))
        assert seen == list(range(8))
    assert window.views_for_rank(8, 1, 8) == [1]
    assert window.views_for_rank(3, 5, 8) == []
    with pytest.raises(ValueError):
        window.views_for_rank(8, 8, 8)


def test_oracle_edge_cases_empty_and_culled():
    """No visible Gaussian (all behind the camera / all oversized): R = 0, image = background, zero gradients."""
    case = helpers.scene_case(300, 77, 1, intr=dict(width=40, height=24, fx=30.0, fy=30.0))
    case.means3D[:, 0] = -5.0 - case.means3D[:, 0].abs()  # behind the camera (camera looks along +x)
    out = helpers.run_oracle_cpu(case)
    assert out["num_rendered"] == 0 and (out["radii"] == 0).all()
    assert np.allclose(out["out_color"], 1.0) and (out["out_acc"] == 0).all() and (out["n_contrib"] == 0).all()
    for n in helpers.GRAD_NAMES:
        assert not out[n].any()
    case2 = helpers.scene_case(300, 78, 1, intr=dict(width=40, height=24, fx=30.0, fy=30.0))
    case2.scales[:] = 0.31  # scale_abnormal culls everything (forward.cu:19-25)
    out2 = helpers.run_oracle_cpu(case2)
    assert out2["num_rendered"] == 0


def test_oracle_key_layout_and_sortedness():
    case = helpers.scene_case(4000, 5, 0, intr=dict(width=100, height=50, fx=56.5, fy=56.7), heavy=True)
    out = helpers.run_oracle_cpu(case, backward=False)
    keys = out["point_list_keys"]
    assert (np.diff(keys.astype(np.uint64).astype(object)) >= 0).all() if len(keys) < 5000 else (keys[1:] >= keys[:-1]).all()
    tiles = (keys >> np.uint64(32)).astype(np.int64)
    gx, gy = (100 + 15) // 16, (50 + 15) // 16
    assert tiles.max() < gx * gy
    depth_bits = (keys & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    np.testing.assert_array_equal(depth_bits, out["depths"][out["point_list"]].view(np.uint32))
    r = out["ranges"].astype(np.int64)
    lens = r[:, 1] - r[:, 0]
    assert lens.sum() == out["num_rendered"]
    for t in np.nonzero(lens)[0][:50]:
        assert (tiles[r[t, 0]:r[t, 1]] == t).all()
    # stable: equal keys keep ascending Gaussian index
    same = keys[1:] == keys[:-1]
    assert (out["point_list"][1:][same] > out["point_list"][:-1][same]).all()


def test_gradient_layout_is_16_byte_aligned_and_ordered():
    """engine.grad_layout: every array starts on a 16-byte boundary (float4 accesses in geom_bwd and the exchange
    kernels) whatever P, the SH rows come first, the 11 other parameter-gradient floats behind them and the four arrays
    the blend backward accumulates into are adjacent (one memset) when P is a multiple of 4."""
    import gs_livm_b200  # noqa: F401
    from gs_livm_b200 import engine
    for P in (1, 3, 130, 49_999, 1_000_000):
        at, total, param_floats, sh_floats = engine.grad_layout(P, 16)
        assert list(at)[0] == "dL_dsh" and at["dL_dsh"][0] == 0 and sh_floats == at["dL_dmean3D"][0] >= 48 * P
        assert all(off % 4 == 0 for off, _ in at.values()) and total % 4 == 0
        order = [n for n, _ in engine.GRAD_LAYOUT]
        for a, b in zip(order, order[1:]):
            assert at[b][0] >= at[a][0] + at[a][1] * P, "arrays must not overlap"
        assert param_floats == at["dL_dmean2D"][0]
        assert param_floats - sh_floats >= 11 * P
    at = engine.grad_layout(1000, 16)[0]
    acc = ["dL_dopacity", "dL_dmean2D", "dL_dconic", "dL_dcolor"]
    for a, b in zip(acc, acc[1:]):
        assert at[b][0] == at[a][0] + at[a][1] * 1000


def test_bench_arms_share_one_config_object():
    """bench.py: both arms build their `config` with the same function from the same constants (the driver compares the
    two dicts), and everything arm-specific lives elsewhere."""
    import importlib.util
    import inspect
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    a = bench.bench_config(752, 480, [3, 1, 2, 4, 5, 6, 7, 8])
    b = bench.bench_config(752, 480, (3, 1, 2, 4, 5, 6, 7, 8))
    assert a == b and a["P"] == 1_000_000 and a["window_poses"] == 8 and len(a["R_per_pose"]) == 8
    assert set(a) == {"workload", "P", "sh_degree", "width", "height", "window_poses", "R_per_pose", "seed", "l2"}
    for fn in (bench.run_ours, bench.run_reference):
        assert '"config": bench_config(W, H, R_per_pose)' in inspect.getsource(fn)
    names = [e[0] for e in bench.SWEEP]
    assert len(names) == len(set(names)) == 3
