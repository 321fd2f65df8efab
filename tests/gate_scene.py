"""Seeded stereo scene for the geometric match gate tests: true correspondences (consistent geometry, similar
descriptors) and distractors (similar descriptors, inconsistent geometry) -- what the production matching algorithm
(VioKeyframeWindowMatchingAlgorithm) sees."""
import numpy as np

from okvis_b200 import abi, synthetic


def undistort_radtan(cam, px):
    """backProject of a radtan pinhole: fixed-point iteration of the distortion model (input of the gate, not under test)."""
    k1, k2, p1, p2 = cam["dist"][:4]
    d = np.stack([(px[:, 0] - cam["cu"]) / cam["fu"], (px[:, 1] - cam["cv"]) / cam["fv"]], 1)
    u = d.copy()
    for _ in range(30):
        r2 = (u ** 2).sum(1)
        rad = 1 + k1 * r2 + k2 * r2 * r2
        dx = 2 * p1 * u[:, 0] * u[:, 1] + p2 * (r2 + 2 * u[:, 0] ** 2)
        dy = 2 * p2 * u[:, 0] * u[:, 1] + p1 * (r2 + 2 * u[:, 1] ** 2)
        u = np.stack([(d[:, 0] - dx) / rad, (d[:, 1] - dy) / rad], 1)
    return np.concatenate([u, np.ones((len(u), 1))], 1)


def make_scene(seed=0, n_true=300, n_distract=150, nbytes=48):
    rng = np.random.Generator(np.random.PCG64(0x0B200 + 4000 + seed))
    cam = abi.make_camera(abi.DIST_RADTAN, 752, 480, 458.654, 457.296, 367.215, 248.375, [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05])
    T_AB = np.array([0.11, 0.002, -0.003, 0, 0, 0, 1.0])
    rot = synthetic.delta_q(np.array([0.004, -0.01, 0.006]))
    T_AB[3:] = rot
    R_AB = synthetic.R_from_quat(T_AB[3:])
    pts, kpa, kpb = [], [], []
    while len(pts) < n_true:
        p = np.array([rng.uniform(-4, 4), rng.uniform(-2.5, 2.5), rng.uniform(1.5, 12.0)])
        pa, oka = synthetic.project_points(cam, p[None])
        pb, okb = synthetic.project_points(cam, (R_AB.T @ (p - T_AB[:3]))[None])
        if oka[0] and okb[0]:
            pts.append(p); kpa.append(pa[0]); kpb.append(pb[0])
    kpa = np.array(kpa) + rng.normal(0, 0.4, (n_true, 2))
    kpb = np.array(kpb) + rng.normal(0, 0.4, (n_true, 2))
    descA = rng.integers(0, 256, (n_true, nbytes), dtype=np.uint8)

    def flipped(rows, max_flips):
        out = rows.copy()
        for r in out:
            for i in rng.integers(0, nbytes * 8, rng.integers(0, max_flips)):
                r[i >> 3] ^= np.uint8(1 << (i & 7))
        return out

    descB = flipped(descA, 25)
    # distractors in B: look like some A descriptor but sit somewhere else in the image
    src = rng.integers(0, n_true, n_distract)
    dB = flipped(descA[src], 20)
    kB = np.stack([rng.uniform(20, 730, n_distract), rng.uniform(20, 460, n_distract)], 1)
    perm = rng.permutation(n_true + n_distract)
    B = np.concatenate([descB, dB])[perm]
    kp_b = np.concatenate([kpb, kB])[perm]
    truth_b = np.concatenate([np.arange(n_true), -np.ones(n_distract, int)])[perm]      # index of the matching A (or -1)
    size_a = rng.uniform(6, 12, n_true)
    size_b = rng.uniform(6, 12, len(B))
    bearing_a = undistort_radtan(cam, kpa) * rng.uniform(0.5, 2.0, (n_true, 1))       # any length
    bearing_b = undistort_radtan(cam, kp_b)
    rs = lambda size, f: np.sqrt(np.sqrt(2)) * (0.8 * size / 12.0) / f
    # 3D-2D inputs: projection of landmark a into B with a pose-induced uncertainty
    proj = np.zeros((n_true, 2))
    for b, a in enumerate(truth_b):
        if a >= 0:
            proj[a] = kp_b[b]
    proj += rng.normal(0, 0.7, proj.shape)
    unc = np.zeros((n_true, 2, 2))
    for i in range(n_true):
        M = rng.normal(0, 0.6, (2, 2))
        unc[i] = M @ M.T + 0.05 * np.eye(2)
    return dict(cam=cam, T_AB=T_AB, A=descA, B=B, kp_a=kpa, kp_b=kp_b, size_a=size_a, size_b=size_b, bearing_a=bearing_a, bearing_b=bearing_b,
                ray_sigma_a=rs(size_a, cam["fu"]), ray_sigma_b=rs(size_b, cam["fu"]), proj=proj, unc=unc.reshape(n_true, 4), truth_b=truth_b)


def gates(sc):
    g3 = abi.make_match_gate(abi.GATE_3D2D, sc["kp_b"], sc["size_b"], proj_into_b=sc["proj"], proj_uncertainty=sc["unc"])
    g2 = abi.make_match_gate(abi.GATE_2D2D, sc["kp_b"], sc["size_b"], kp_a=sc["kp_a"], kp_size_a=sc["size_a"], bearing_a=sc["bearing_a"],
                             bearing_b=sc["bearing_b"], ray_sigma_a=sc["ray_sigma_a"], ray_sigma_b=sc["ray_sigma_b"], cam_a=sc["cam"],
                             cam_b=sc["cam"], T_AB=sc["T_AB"])
    return g3, g2
