"""CPU checks of the frontend oracle (this project's BRISK-style specification; parity with
brisk 2.0.5 is unpinned -- the reference's TestFrame.cpp:47-84 asserts nothing about keypoints)."""
import numpy as np

from okvis_b200 import abi, images

CAM = abi.make_camera(abi.DIST_RADTAN, 752, 480, 458.654880721, 457.296696463, 367.215803962, 248.37534061,
                      (-0.28340811217, 0.0739590738929, 0.000193595028569, 1.76187114545e-05))


def test_detect_properties(oracle):
    img = images.textured_image()
    kps, desc = oracle.detect_describe(img, CAM, np.eye(3), uniformity_radius=15, max_keypoints=1000)
    assert 500 <= len(kps) <= 1000
    xy = np.stack([kps["x"], kps["y"]], 1).astype(np.float64)
    assert xy[:, 0].min() >= 16 and xy[:, 0].max() < 752 - 16 and xy[:, 1].min() >= 16 and xy[:, 1].max() < 480 - 16
    # descending score, uniformity radius respected
    assert np.all(np.diff(kps["response"]) <= 0)
    d2 = ((xy[:, None, :] - xy[None, :, :]) ** 2).sum(2)
    np.fill_diagonal(d2, 1e9)
    assert d2.min() >= 15 * 15
    assert np.all(kps["response"] >= 800)
    assert desc.shape == (len(kps), 48) and desc.dtype == np.uint8
    # max_keypoints truncates to the strongest
    k2, _ = oracle.detect_describe(img, CAM, np.eye(3), uniformity_radius=15, max_keypoints=100)
    assert len(k2) == 100 and np.array_equal(k2["x"], kps["x"][:100])


def test_descriptor_is_distinctive_and_matches_second_view(oracle):
    left, right = images.stereo_pair()
    ka, da = oracle.detect_describe(left, CAM, np.eye(3), uniformity_radius=15, max_keypoints=1000)
    kb, db = oracle.detect_describe(right, CAM, np.eye(3), uniformity_radius=15, max_keypoints=1000)
    res = oracle.match_hamming(da, db, threshold=60.0)
    m = res["matches"]
    assert len(m) > 0.3 * min(len(ka), len(kb))
    dx = ka["x"][m[:, 0]] - kb["x"][m[:, 1]]
    dy = ka["y"][m[:, 0]] - kb["y"][m[:, 1]]
    good = (np.abs(dx - 12) <= 1.5) & (np.abs(dy) <= 1.5)
    assert good.mean() > 0.9          # matches follow the 12 px disparity
    # self distances are zero, random pairs are far
    D = np.unpackbits(da[:50, None, :] ^ da[None, :50, :], axis=2).sum(2)
    assert np.all(np.diag(D) == 0) and np.median(D[np.triu_indices(50, 1)]) > 100


def test_gravity_angle_follows_rotation(oracle):
    img = images.textured_image()
    k0, _ = oracle.detect_describe(img, CAM, np.eye(3), uniformity_radius=40, max_keypoints=50)
    # camera looking horizontally: gravity (0,0,-1)_W maps to +y in the image => angle ~ +90 deg
    R_CW = np.array([[1, 0, 0], [0, 0, -1], [0, 1, 0]], dtype=np.float64)
    k1, _ = oracle.detect_describe(img, CAM, R_CW, uniformity_radius=40, max_keypoints=50)
    assert np.all(np.abs(k1["angle"] - 90.0) < 25.0)
    k2, d2 = oracle.detect_describe(img, CAM, R_CW, uniformity_radius=40, max_keypoints=50, rotation_invariance=False)
    assert np.all(k2["angle"] == 0)


def test_64_byte_descriptors(oracle):
    img = images.textured_image()
    k, d = oracle.detect_describe(img, CAM, np.eye(3), uniformity_radius=40, max_keypoints=100, desc_bytes=64)
    k48, d48 = oracle.detect_describe(img, CAM, np.eye(3), uniformity_radius=40, max_keypoints=100, desc_bytes=48)
    assert d.shape[1] == 64 and np.array_equal(d[:, :48], d48)   # the 384 shortest pairs come first
