"""CPU: the oracles against the golden fixtures (made by oracle/make_golden.py with the reference's
own utils/pcd.py and kitti_utils.Calibration, imported from /root/reference at generation time)."""
import numpy as np
import pytest

from conftest import GOLDEN, load_golden


@pytest.mark.parametrize("name", GOLDEN)
def test_c_oracle_matches_golden(name):
    from oracle import c_oracle as CO
    frame, d = load_golden(name)
    r = CO.lift_frame_c(frame)
    assert r["n_points"] == int(d["n_points"])
    assert np.array_equal(r["aggr"].view(np.uint32), d["aggr"].view(np.uint32))
    idx = np.concatenate(r["idx"]) if r["idx"] else np.zeros(0, np.int64)
    assert np.array_equal(idx, d["seg_point_idx"])
    assert np.array_equal(np.cumsum([0] + [len(x) for x in r["idx"]]), d["seg_offsets"])
    assert np.array_equal(r["medoid_local"], d["medoid_local"])
    assert np.array_equal(r["medoid_point_idx"], d["medoid_point_idx"])
    has = d["medoid_local"] >= 0
    assert np.array_equal(r["centroids"][has].view(np.uint32), d["centroids"][has].view(np.uint32))
    for c in d["pix_cams"]:
        sel, fx, fy = r["pix"][int(c)]
        assert np.array_equal(sel, d[f"pix_{c}_idx"])
        assert np.array_equal(fx, d[f"pix_{c}_fx"])
        assert np.array_equal(fy, d[f"pix_{c}_fy"])


@pytest.mark.parametrize("name", ["nusc_small", "nusc_edge", "kitti_small", "waymo_small"])
def test_torch_restatement_matches_golden(name):
    """oracle/ref_lift.py with its own restated helpers (no reference import) reproduces the
    fixtures that were generated with the reference's helpers injected."""
    import torch
    from oracle import ref_lift as RL
    torch.set_num_threads(1)
    frame, d = load_golden(name)
    r = RL.lift_frame(frame)
    assert np.array_equal(np.ascontiguousarray(r["aggr"]).view(np.uint32), d["aggr"].view(np.uint32))
    idx = np.concatenate(r["idx"]) if r["idx"] else np.zeros(0, np.int64)
    assert np.array_equal(idx, d["seg_point_idx"])
    assert np.array_equal(r["medoid_local"], d["medoid_local"])


def test_edge_fixture_covers_the_quirks():
    """nusc_edge holds: an empty mask, a full-image mask (fx==0 / fy==0 points dropped), a 1-px
    line (erodes to nothing), a 3x3 block (erodes to one pixel), masks touching the border."""
    frame, d = load_golden("nusc_edge")
    cnt = np.diff(d["seg_offsets"])
    assert cnt[0] == 0 and cnt[2] == 0            # empty mask, eroded-away line
    assert cnt[1] > 0                             # full image
    full = d["seg_point_idx"][d["seg_offsets"][1]:d["seg_offsets"][2]]
    c = int(frame.cam_nums[1])
    sel = d[f"pix_{c}_idx"]
    fx, fy = d[f"pix_{c}_fx"], d[f"pix_{c}_fy"]
    inside = sel[(fx != 0) & (fy != 0) & (fx < 511) & (fy < 287)]
    # every in-image point of that camera is a member unless its floor is 0 or on the eroded border
    assert set(full.tolist()) <= set(sel.tolist())
    assert not (set(sel[(fx == 0) | (fy == 0)].tolist()) & set(full.tolist()))
    assert (d["medoid_local"][cnt == 0] == -1).all()


def test_medoid_oracle_vs_torch_many_sizes():
    """C medoid (ATen cascade order, IEEE sqrt) against torch.cdist(...).sum(0).argmin() for every
    M in 1..100 and a few larger sizes, in nuScenes-like global coordinates where cdist's matmul
    formula is noisy.  Sums may differ by the last bit (torch's vectorised sqrt), the argmin not."""
    import torch
    from oracle import c_oracle as CO
    torch.set_num_threads(1)
    rng = np.random.default_rng(5)
    worst = 0.0
    for m in list(range(1, 101)) + [127, 128, 129, 255, 256, 257, 511, 513, 1000, 2049]:
        p = (rng.normal(0, 2.0, (3, m)) + np.array([[1234.5], [987.25], [1.5]])).astype(np.float32)
        t = torch.from_numpy(p)
        ref = torch.cdist(t.T, t.T, p=2).sum(axis=0)
        j, sums = CO.medoid(p, want_sums=True)
        rs = ref.numpy()
        worst = max(worst, float(np.max(np.abs(sums - rs) / np.maximum(np.abs(rs), 1e-6))))
        assert j == int(torch.argmin(ref)), m
    assert worst < 1e-6


def test_obb_oracle_yaw_matches_live_scipy_on_proper_rotations():
    """oracle/obb_oracle.py restates scipy 1.11.4's from_matrix/as_euler; where the installed scipy
    accepts the input (right-handed R') both must agree."""
    from scipy.spatial.transform import Rotation
    from oracle import obb_oracle as O
    rng = np.random.default_rng(3)
    proper = improper = 0
    for _ in range(200):
        m = int(rng.integers(4, 300))
        p = rng.uniform(-0.5, 0.5, (m, 3)) * rng.uniform(0.3, 5, 3)
        p = (p @ Rotation.from_euler("zyx", rng.uniform(-3, 3, 3)).as_matrix().T + rng.uniform(-20, 20, 3)).astype(np.float32)
        center, wlh, R = O.get_depth_bbox(p)
        assert abs(abs(np.linalg.det(R)) - 1) < 1e-9
        if np.linalg.det(R) > 0:
            ref = Rotation.from_matrix(R).as_euler("zyx")[0]
            assert abs(((O.yaw_of(R) - ref + np.pi) % (2 * np.pi)) - np.pi) < 1e-9
            proper += 1
        else:
            assert np.isfinite(O.yaw_of(R))
            improper += 1
    assert proper > 20 and improper > 20


def _screen_groups(p):
    """Python restatement of k_medoid_classify / k_medoid_permute (cm3d_b200/csrc/medoid.cu): None when the
    instance is not eligible for a symmetric screen, else the group (0 lower binade, 1 its top sliver, 2 upper
    binade) of every point; one binade -> all zeros."""
    x, y, z = p.astype(np.float32)
    n = (x * x + y * y) + z * z
    bits = n.view(np.uint32)
    e, mant = bits >> 23, bits & 0x7FFFFF
    ext = np.array([np.ptp(x), np.ptp(y), np.ptp(z)], np.float32)
    diag2 = np.float32(ext @ ext) * np.float32(1.001)
    nmin, nmax = n.min(), n.max()
    top_margin = (np.float32(nmax).view(np.uint32) & 0x7FFFFF) <= 0x7FFFE0
    if not (nmin > 0 and top_margin and diag2 <= np.float32(0.25) * nmin):
        return None
    if e.min() == e.max():
        return np.zeros(n.size, int)
    if e.max() == e.min() + 1:
        return np.where(e == e.min(), np.where(mant <= 0x7FFFF8, 0, 1), 2)
    return None


def test_squared_distance_matrix_is_symmetric_where_the_screen_assumes_it():
    """The proof behind the CUDA medoid's symmetric screen, checked on the C oracle's fp32 arithmetic: inside a
    group that k_medoid_classify declares symmetric, r(i, j) == r(j, i) bit for bit; across the groups it is
    not in general (so the test can see an asymmetry when there is one)."""
    from oracle import c_oracle as CO
    rng = np.random.default_rng(3)
    cases = []
    for c, sg in (((1200.0, 950.0, 1.0), 3.0), ((420.0, 1900.0, -0.5), 6.0), ((1023.0, 1025.0, 0.0), 1.5), ((1024.0, 1024.0, 0.0), 1.5),
                  ((35.0, 2.0, 15.0), 0.5), ((6.0, -1.0, 5.5), 0.15), ((0.9, 0.1, 0.2), 0.02), ((300.5, 300.5, 2.0), 8.0),
                  ((724.0, 724.2, 0.3), 4.0), ((1448.0, 10.0, 0.0), 2.0), ((2000.0, 2000.0, 1.0), 10.0)):
        for m in (300, 700):
            cases.append((rng.normal(0, sg, (3, m)) + np.array(c)[:, None]).astype(np.float32))
    n_sym = n_grouped = cross_asym = 0
    for p in cases:
        g = _screen_groups(p)
        if g is None:
            continue
        r = CO.sqdist(p).view(np.uint32)
        same = (g[:, None] == g[None, :]) & (g[:, None] != 1)
        assert np.array_equal(r[same], r.T[same]), p.mean(1)
        if g.max() == 0:
            n_sym += 1
        else:
            n_grouped += 1
            cross_asym += int((r != r.T)[~same].sum())
    assert n_sym >= 8 and n_grouped >= 3 and cross_asym > 0
    # and an instance that fails the conditions really is asymmetric (sensor frame, several binades)
    p = (rng.normal(0, 2, (3, 400)) + np.float32(10.0)).astype(np.float32)
    assert _screen_groups(p) is None
    r = CO.sqdist(p).view(np.uint32)
    assert (r != r.T).any()


def test_screen_error_bound_of_the_blocked_summation():
    """The bound the CUDA screen's candidate threshold rests on (csrc/medoid.cu: screen_threshold): the screen's
    summation order (16 rows -> a0, a 1024-row tile -> a1, tiles -> a2) over square roots that are one ulp off
    is within (168 + M/1024 + M/4096) u of the reference's cascade sum, u = 2^-24.  Restated in numpy on the C
    oracle's distances; what is observed stays far inside the bound."""
    from oracle import c_oracle as CO
    rng = np.random.default_rng(5)
    worst = 0.0
    for m, c, sg in ((700, (1200.0, 950.0, 1.0), 3.0), (2100, (420.0, 1900.0, -0.5), 5.0), (1500, (10.0, 1.0, 20.0), 2.0)):
        p = (rng.normal(0, sg, (3, m)) + np.array(c)[:, None]).astype(np.float32)
        _, ref = CO.medoid(p, want_sums=True)
        d = np.sqrt(np.maximum(CO.sqdist(p), np.float32(0.0)))                 # rows i, columns j, correctly rounded
        d1 = np.nextafter(d, np.float32(np.inf)) * (d > 0)                     # a root that is one ulp off everywhere
        blk = np.zeros(m, np.float32)
        for t0 in range(0, m, 1024):
            a1 = np.zeros(m, np.float32)
            for b in range(t0, min(m, t0 + 1024), 16):
                a0 = np.zeros(m, np.float32)
                for i in range(b, min(m, b + 16)):
                    a0 += d1[i]
                a1 += a0
            blk += a1
        bound = (168 + m // 1024 + m // 4096) * 2.0 ** -24
        rel = float(np.max(np.abs(blk.astype(np.float64) - ref) / ref))
        assert rel <= bound, (m, rel, bound)
        worst = max(worst, rel / bound)
    assert worst < 0.25
