"""Crop + resize in front of the scoring path (app.py:1964-1978): the oracle is pinned bit-exactly on
Pillow (the reference's own dependency), the CUDA kernels bit-exactly on the oracle and on Pillow."""
import numpy as np
import pytest
import torch
from PIL import Image


def _cases(seed, n, hmax=500, wmax=640):
    rng = np.random.default_rng(seed)
    out = []
    for it in range(n):
        H, W = int(rng.integers(24, hmax)), int(rng.integers(24, wmax))
        if it % 3 == 0:            # smooth content exercises the rounding of the negative lobes, noise the clipping
            fr = (np.add.outer(np.arange(H) * 3, np.arange(W) * 2)[..., None] * np.array([1, 2, 3]) % 256).astype(np.uint8)
        else:
            fr = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        x1 = int(rng.integers(0, W - 1)); x2 = int(rng.integers(x1 + 1, W + 1))
        y1 = int(rng.integers(0, H - 1)); y2 = int(rng.integers(y1 + 1, H + 1))
        out.append((fr, (x1, y1, x2, y2)))
    fr = out[0][0]; H, W = fr.shape[:2]
    out += [(fr, (0, 0, 1, 1)), (fr, (0, 0, W, H)), (fr, (W - 2, H - 3, W, H)), (fr, (0, 0, min(W, 224), min(H, 224)))]   # edge boxes
    return out


def _pil(fr, box, size=224):
    return np.array(Image.fromarray(fr).convert("RGB").crop(box).resize((size, size)))       # the reference's own call


def test_resize_oracle_is_bit_exact_with_pillow():
    from oracle.pil_resize_oracle import crop_resize
    for fr, box in _cases(0, 40):
        assert np.array_equal(crop_resize(fr, box), _pil(fr, box)), (fr.shape, box)
    fr, box = _cases(1, 1)[0]
    assert np.array_equal(crop_resize(fr, box, 96), _pil(fr, box, 96))                          # another face_size


def test_clamp_boxes_follows_the_reference():
    from deepfake_video_detection_b200.crop_resize import clamp_boxes
    kept, idx = clamp_boxes([(0, -5.7, 3.2, 50.9, 40.1), (1, 10, 10, 10, 30), (0, 90, 90, 400, 400), (0, 120, 5, 130, 8)], 100, 100)
    assert kept == [(0, 0, 3, 50, 40), (0, 90, 90, 100, 100)] and idx == [0, 2]                 # app.py:1964-1975


@pytest.mark.gpu
def test_crop_resize_cuda_bit_exact():
    from deepfake_video_detection_b200.crop_resize import crop_resize
    from oracle.pil_resize_oracle import crop_resize as oracle_resize
    rng = np.random.default_rng(3)
    frames = rng.integers(0, 256, (3, 360, 480, 3), dtype=np.uint8)
    frames[1] = (np.add.outer(np.arange(360) * 2, np.arange(480))[..., None] * np.array([1, 3, 5]) % 256).astype(np.uint8)
    boxes = [(0, 10, 20, 234, 244), (1, 0, 0, 480, 360), (2, 100, 50, 131, 99), (1, 300, 200, 480, 360), (0, 5, 5, 6, 6),
             (2, 17, 33, 440, 350), (0, 200, 100, 424, 324), (1, -20.5, -3, 50.2, 1000), (2, 50, 50, 50, 80)]
    faces, idx = crop_resize(torch.from_numpy(frames).cuda(), boxes)
    assert idx == [0, 1, 2, 3, 4, 5, 6, 7] and faces.shape == (8, 224, 224, 3)
    faces = faces.cpu().numpy()
    for j, i in enumerate(idx):
        f, x1, y1, x2, y2 = boxes[i]
        box = (max(0, int(x1)), max(0, int(y1)), min(480, int(x2)), min(360, int(y2)))
        assert np.array_equal(faces[j], oracle_resize(frames[f], box)), boxes[i]
        assert np.array_equal(faces[j], _pil(frames[f], box)), boxes[i]
    small, _ = crop_resize(torch.from_numpy(frames).cuda(), boxes[:3], size=64)
    assert np.array_equal(small[2].cpu().numpy(), _pil(frames[2], (100, 50, 131, 99), 64))
    with pytest.raises(ValueError):
        crop_resize(torch.from_numpy(frames), boxes)                                             # CPU tensor: no fallback


@pytest.mark.gpu
def test_crop_resize_feeds_the_scorer(synth_sd):
    """frames + boxes -> crops on the GPU -> logits == logits of the crops Pillow makes (bit-identical inputs)."""
    from deepfake_video_detection_b200 import FrameScorer, make_offsets
    from deepfake_video_detection_b200.crop_resize import crop_resize
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 256, (4, 300, 400, 3), dtype=np.uint8)
    boxes = [(i, 40 + 10 * i, 30, 300 + 5 * i, 280) for i in range(4)]
    faces, _ = crop_resize(torch.from_numpy(frames).cuda(), boxes)
    ref = np.stack([_pil(frames[f], (x1, y1, x2, y2)) for f, x1, y1, x2, y2 in boxes])
    sc = FrameScorer(synth_sd, device="cuda")
    off = make_offsets([4], "cuda")
    a, _ = sc.score(faces, off)
    b, _ = sc.score(torch.from_numpy(ref).cuda(), off)
    assert torch.equal(a, b)
