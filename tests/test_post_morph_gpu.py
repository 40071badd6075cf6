"""GPU parity for the post-filters: speckle, median, validateDisparity, erode/dilate, open+close."""
import numpy as np
import pytest

from conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu


def test_speckle_matches_golden(gpu):
    g = load_golden("post_speckle_320x240")
    for key in g.files:
        if key.startswith("sp_"):
            _, ms, md = key.split("_")
            assert np.array_equal(gpu.filter_speckles(g["raw"], -16, int(ms), int(md)), g[key]), key


def test_speckle_random_and_adversarial(gpu, orc):
    rng = np.random.default_rng(3)
    for (W, H) in [(1, 1), (1, 7), (9, 1), (33, 17), (257, 129), (640, 480)]:
        a = (rng.integers(0, 6, (H, W)) * 16).astype(np.int16)
        a[rng.random((H, W)) < 0.2] = -16
        for (ms, md) in [(0, 0), (5, 16), (50, 0), (100000, 32)]:
            assert np.array_equal(gpu.filter_speckles(a, -16, ms, md), orc.filter_speckles(a, -16, ms, md)), (W, H, ms, md)
    # serpentine component: one long thin snake (worst case for union-find depth)
    a = np.full((64, 64), -16, np.int16)
    for r in range(0, 64, 2):
        a[r, :] = 100
        if r + 1 < 64:
            a[r + 1, 63 if (r // 2) % 2 == 0 else 0] = 100
    for ms in (10, 5000):
        assert np.array_equal(gpu.filter_speckles(a, -16, ms, 1), orc.filter_speckles(a, -16, ms, 1))
    # one big constant plane (single huge component) + isolated pixels
    a = np.full((200, 300), 160, np.int16); a[::7, ::5] = 800
    assert np.array_equal(gpu.filter_speckles(a, -16, 100, 32), orc.filter_speckles(a, -16, 100, 32))


def test_speckle_vector_and_scalar_kernels_agree(gpu, orc, monkeypatch):
    """W % 8 == 0 takes the 8-pixels-per-thread kernels; RTDM_SPECKLE_SCALAR forces the per-pixel ones."""
    rng = np.random.default_rng(31)
    for (W, H) in [(8, 1), (8, 9), (40, 50), (1280, 40), (2048, 3)]:
        a = (rng.integers(0, 4, (H, W)) * 24).astype(np.int16)
        a[rng.random((H, W)) < 0.3] = -16
        a[:, : W // 3] = 48                                   # long runs crossing many threads
        ref = orc.filter_speckles(a, -16, 60, 24)
        monkeypatch.delenv("RTDM_SPECKLE_SCALAR", raising=False)
        assert np.array_equal(gpu.filter_speckles(a, -16, 60, 24), ref), (W, H, "vector")
        monkeypatch.setenv("RTDM_SPECKLE_SCALAR", "1")
        assert np.array_equal(gpu.filter_speckles(a, -16, 60, 24), ref), (W, H, "scalar")
    monkeypatch.delenv("RTDM_SPECKLE_SCALAR", raising=False)


def test_median_matches_golden_and_oracle(gpu, orc):
    g = load_golden("post_median_131x97")
    assert np.array_equal(gpu.median3_s16(g["src"]), g["median"])
    rng = np.random.default_rng(4)
    for (W, H) in [(1, 1), (2, 2), (1, 9), (9, 1), (640, 480)]:
        a = rng.integers(-16, 4096, (H, W)).astype(np.int16)
        assert np.array_equal(gpu.median3_s16(a), orc.median3_s16(a)), (W, H)


def test_validate_matches_golden(gpu):
    g = load_golden("post_validate_320x240")
    for d12 in (0, 1, 3):
        assert np.array_equal(gpu.validate_disparity(g["raw"], g["cost"], 0, 64, d12), g[f"d12_{d12}"]), d12


@pytest.mark.parametrize("name", golden_names("morph_"))
def test_morph_matches_golden(gpu, name):
    g = load_golden(name)
    src = g["src"]
    H, W = src.shape
    assert np.array_equal(gpu.morph_op(src, 0), g["erode"])
    assert np.array_equal(gpu.morph_op(src, 1), g["dilate"])
    f = gpu.CUDAMorphologicalFilter(W, H, 8)
    assert (f.getWidth(), f.getHeight(), f.getBpp(), f.getFrameSize()) == (W, H, 8, W * H)
    # Estimator writes into the device-owned input buffer and reads the output buffer
    f.getVideoInBuffer()[:] = src
    assert f.run() == 0
    assert np.array_equal(f.getVideoOutBuffer(), g["openclose"])
    out = np.empty_like(src)
    f.run(src, out)
    assert np.array_equal(out, g["openclose"])
    assert f.last_launches() >= 4


def test_morph_720p_and_other_kernels(gpu, orc):
    from rtdm_b200 import synth
    m = synth.binary_mask(1280, 720, 9)
    f = gpu.CUDAMorphologicalFilter(1280, 720, 8)
    out = np.empty_like(m)
    f.run(m, out)
    assert np.array_equal(out, orc.morph_open_close(m))
    # idempotence of open-close on its own output's opening (size-independent property)
    out2 = np.empty_like(m)
    f.run(out, out2)
    assert np.array_equal(out2, orc.morph_open_close(out))
    g = synth.gray_image(333, 211, 10)
    for (kw, kh) in [(3, 3), (5, 9), (10, 10), (21, 7), (31, 31), (1, 1)]:
        for op in (0, 1):
            assert np.array_equal(gpu.morph_op(g, op, kw, kh), orc.morph(g, op, kw, kh)), (kw, kh, op)


def test_filter_binary_fast_path_and_gray_fallback_in_one_batch(gpu, orc):
    """Binary {0,255} masks take the bit-packed kernel; any other byte value makes that frame fall back to the
    generic kernels.  Both must agree with the oracle, also when mixed in one batch and at awkward sizes."""
    from rtdm_b200 import synth
    for (W, H) in [(1280, 720), (934, 404), (225, 33), (31, 7)]:
        frames = [synth.binary_mask(W, H, 40), synth.gray_image(W, H, 41), synth.binary_mask(W, H, 42)]
        almost = synth.binary_mask(W, H, 43).copy(); almost[H // 2, W // 3] = 254      # one stray value
        frames.append(almost)
        full = np.full((H, W), 255, np.uint8); empty = np.zeros((H, W), np.uint8)
        frames += [full, empty]
        batch = np.stack(frames)
        f = gpu.CUDAMorphologicalFilter(W, H, 8, max_batch=len(frames))
        out = f.run_batch(batch)
        for i, fr in enumerate(frames):
            assert np.array_equal(out[i], orc.morph_open_close(fr)), (W, H, i)
        # in-place device-style call through the single-frame API
        single = np.empty((H, W), np.uint8)
        f.run(frames[0], single)
        assert np.array_equal(single, orc.morph_open_close(frames[0]))


def test_filter_device_call_in_place_and_out_of_place(gpu, orc):
    """rtdm_morph_run_device with device pointers: separate output (fast path writes it directly) and
    in-place (input == output: the fast path goes through a scratch plane)."""
    import torch
    from rtdm_b200 import synth
    W, H = 640, 360
    frames = np.stack([synth.binary_mask(W, H, 50), synth.gray_image(W, H, 51), synth.binary_mask(W, H, 52)])
    ref = np.stack([orc.morph_open_close(f) for f in frames])
    f = gpu.CUDAMorphologicalFilter(W, H, 8, max_batch=3)
    src = torch.from_numpy(frames).cuda()
    dst = torch.empty_like(src)
    f.run_device(3, src.data_ptr(), dst.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(dst.cpu().numpy(), ref)
    f.run_device(3, src.data_ptr(), src.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(src.cpu().numpy(), ref)


def test_filter_async_batch(gpu, orc):
    import torch
    from rtdm_b200 import synth
    W, H = 320, 200
    frames = np.stack([synth.binary_mask(W, H, 60 + i) for i in range(3)])
    pin = torch.from_numpy(frames).pin_memory()
    out = torch.empty_like(pin).pin_memory()
    f = gpu.CUDAMorphologicalFilter(W, H, 8, max_batch=3)
    f.run_batch_async(pin.numpy(), out.numpy())
    f.sync()
    for i in range(3):
        assert np.array_equal(out.numpy()[i], orc.morph_open_close(frames[i]))
