"""Input leg: ImageDataset.__getitem__ (extract_localization.py:158-190).  The oracle restates cv2's float
INTER_CUBIC resize; it is pinned against cv2 itself (opencv-python in this image) and against a committed fixture made by
cv2 (oracle/make_golden_preprocess.py).  The GPU tests compare the CUDA kernel / the device loader with both."""
import os

import numpy as np
import pytest
import torch

from oracle import sfd2_oracle as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden", "preprocess_cases.npz")
# cv2's own SIMD path differs from its scalar path by up to ~0.03 grey levels on down-scales (different float
# rounding of the source coordinate); the restatement follows the scalar code and equals it to fp32 rounding
TOL_SCALAR, TOL_SIMD = 2e-4 / 255, 0.05 / 255


def _img(seed, h, w):
    from sfd2_b200.synth import synth_image_u8
    return np.ascontiguousarray(synth_image_u8(seed, h, w, sigma=1.5)[:, :, ::-1])     # "BGR as imread returns it"


def test_resize_target_matches_reference_rule():
    assert orc.resize_target(3024, 4032, 1600) == (1200, 1600)
    assert orc.resize_target(1063, 1600, 1600) == (1063, 1600)          # max side == resize_max: untouched
    assert orc.resize_target(600, 800, 1600) == (600, 800)
    assert orc.resize_target(600, 800, 1600, resize_force=True) == (1200, 1600)
    assert orc.resize_target(1000, 3001, 1024) == (341, 1024)
    from sfd2_b200.preprocess import resize_target
    for args in [(3024, 4032, 1600), (777, 1333, 1024), (600, 800, 1600, True), (50, 40, None)]:
        assert resize_target(*args) == orc.resize_target(*args)


def test_oracle_resize_equals_cv2_scalar_path():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(0)
    opt = cv2.useOptimized()
    try:
        for (h, w, hn, wn) in [(300, 400, 1200, 1600), (480, 640, 300, 400), (531, 800, 398, 600), (37, 53, 111, 160), (64, 48, 64, 100)]:
            im = (rng.rand(h, w, 3) * 255).astype(np.uint8).astype(np.float32)
            mine = orc.cv2_resize_cubic_f32(im, wn, hn)
            cv2.setUseOptimized(False)
            assert np.abs(cv2.resize(im, (wn, hn), interpolation=cv2.INTER_CUBIC) - mine).max() <= TOL_SCALAR * 255
            cv2.setUseOptimized(True)
            assert np.abs(cv2.resize(im, (wn, hn), interpolation=cv2.INTER_CUBIC) - mine).max() <= TOL_SIMD * 255
    finally:
        cv2.setUseOptimized(opt)


def test_oracle_item_equals_committed_cv2_fixture():
    g = np.load(GOLD)
    for tag in ("up", "down", "same"):
        item = orc.image_dataset_item(g[f"{tag}_bgr"], resize_max=int(g[f"{tag}_resize_max"]), resize_force=bool(g[f"{tag}_force"]))
        assert item["image"].shape == g[f"{tag}_image"].shape and item["image"].dtype == np.float32
        assert np.abs(item["image"] - g[f"{tag}_image"]).max() <= TOL_SIMD
        assert np.array_equal(item["original_size"], g[f"{tag}_original_size"])


@pytest.mark.gpu
def test_preprocess_kernel_equals_oracle_and_cv2():
    from gpu_util import model
    from sfd2_b200.preprocess import preprocess_dev
    ctx = model("exact").ctx
    g = np.load(GOLD)
    for tag in ("up", "down", "same"):
        bgr = g[f"{tag}_bgr"]
        out = preprocess_dev(torch.from_numpy(bgr).cuda(), ctx, resize_max=int(g[f"{tag}_resize_max"]), resize_force=bool(g[f"{tag}_force"]))
        got = out[0].cpu().numpy()
        ref = orc.image_dataset_item(bgr, resize_max=int(g[f"{tag}_resize_max"]), resize_force=bool(g[f"{tag}_force"]))["image"]
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() <= TOL_SCALAR, tag           # the kernel follows the scalar code operation by operation
        assert np.abs(got - g[f"{tag}_image"]).max() <= TOL_SIMD, tag
    # full size: a 3024 x 4032 photo to the r1600 preset, and the RGB (no swap) variant
    big = _img(3, 1512, 2016)
    out = preprocess_dev(torch.from_numpy(big).cuda(), ctx, resize_max=1600)[0].cpu().numpy()
    ref = orc.image_dataset_item(big, resize_max=1600)["image"]
    assert out.shape == (3, 1200, 1600) and np.abs(out - ref).max() <= TOL_SCALAR
    rgb = preprocess_dev(torch.from_numpy(np.ascontiguousarray(big[:, :, ::-1])).cuda(), ctx, resize_max=1600, bgr=False)[0].cpu().numpy()
    assert np.array_equal(rgb, out)


@pytest.mark.gpu
def test_device_loader_feeds_the_reference_loop(tmp_path):
    """Files on disk -> DeviceImageLoader -> the reference's loop body (extract_to_store) must give the same features as
    the reference's own ImageDataset item (restated by the oracle, resize by cv2 itself) through the same extractor."""
    cv2 = pytest.importorskip("cv2")
    from gpu_util import model
    from sfd2_b200 import extract_resnet_return
    from sfd2_b200.preprocess import DeviceImageLoader
    sizes = [(240, 320), (300, 200), (151, 333), (480, 640), (240, 320)]
    names = []
    for i, (h, w) in enumerate(sizes):
        names.append(f"db/img_{i}.png")
        os.makedirs(tmp_path / "db", exist_ok=True)
        cv2.imwrite(str(tmp_path / names[-1]), _img(10 + i, h, w))
    lst = tmp_path / "list.txt"
    lst.write_text("\n".join(names) + "\n")
    m = model("exact")
    loader = DeviceImageLoader(tmp_path, {"resize_max": 256, "grayscale": False}, m, image_list=str(lst), workers=2, depth=2)
    assert len(loader) == len(sizes)
    seen = []
    for data in loader:
        name = data["name"][0]
        seen.append(name)
        bgr = cv2.imread(str(tmp_path / name), cv2.IMREAD_COLOR)
        item = orc.image_dataset_item(bgr, resize_max=256, use_cv2=True)
        assert data["image"].is_cuda and tuple(data["image"].shape[1:]) == item["image"].shape
        assert np.array_equal(data["original_size"][0].numpy(), item["original_size"])
        assert np.abs(data["image"][0].cpu().numpy() - item["image"]).max() <= TOL_SIMD
        a = extract_resnet_return(m, data["image"], topK=300, conf_th=0.001, scales=[1.0])
        b = extract_resnet_return(m, torch.from_numpy(item["image"][None]), topK=300, conf_th=0.001, scales=[1.0])
        ka, kb = set(map(tuple, a["keypoints"].astype(int))), set(map(tuple, b["keypoints"].astype(int)))
        assert len(ka ^ kb) <= max(2, len(kb) // 50), (name, len(ka ^ kb), len(kb))
    assert seen == names
