"""Image resize ahead of the hot path: torchvision `Resize((224, 224))` on PIL images
(deephumor_demo.ipynb cell 11, data/datasets.py:48-53,94-98).  CPU: oracle/resize.py (numpy restatement of Pillow's
ImagingResample) against PIL.Image.resize and the torchvision transform themselves, bit for bit.  GPU: dh_resize_bilinear_u8
against the oracle and against PIL on the same images, bit for bit, then through the fused uint8 stem."""
import numpy as np
import pytest
import torch
from PIL import Image

from oracle.resize import resize_bilinear

SIZES = [(224, 224), (300, 500), (500, 300), (100, 80), (1000, 1333), (224, 500), (777, 224), (50, 900), (3, 3), (1, 1),
         (1, 300), (300, 1), (2000, 17), (1701, 17), (1700, 17), (501, 5), (500, 5), (201, 2), (225, 2), (224, 2), (1200, 1600)]


def image(h, w, seed, kind='noise'):
    rng = np.random.default_rng(seed)
    if kind == 'noise':
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    if kind == 'extremes':                                     # saturating blocks: exercises the clip to [0, 255]
        return (rng.integers(0, 2, (h, w, 3)) * 255).astype(np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    return np.stack([(xx * 255 // max(w - 1, 1)), (yy * 255 // max(h - 1, 1)), ((xx + yy) % 256)], -1).astype(np.uint8)


def pil_resize(a, out=224):
    return np.asarray(Image.fromarray(a).resize((out, out), Image.BILINEAR))


@pytest.mark.parametrize('kind', ['noise', 'extremes', 'gradient'])
@pytest.mark.parametrize('h,w', SIZES)
def test_oracle_resize_equals_pillow(h, w, kind):
    a = image(h, w, h * 7919 + w, kind)
    assert np.array_equal(resize_bilinear(a), pil_resize(a))


def test_oracle_resize_equals_the_torchvision_transform():
    """The reference's transform chain up to ToTensor (nb cell 11): Resize((224,224)) + ToTensor on a PIL image."""
    import torchvision.transforms as T
    a = image(480, 640, 3)
    tv = T.Compose([T.Resize((224, 224)), T.ToTensor()])(Image.fromarray(a))
    mine = torch.from_numpy(resize_bilinear(a)).permute(2, 0, 1).float() / 255.0
    assert torch.equal(tv, mine)


@pytest.mark.parametrize('out', [50, 100, 256])
def test_oracle_resize_other_output_sizes(out):
    for h, w in [(300, 500), (40, 30), (10001 // 10, 9), (out, out)]:
        a = image(h, w, out + h)
        assert np.array_equal(resize_bilinear(a, out, out), pil_resize(a, out))


@pytest.mark.gpu
def test_cuda_resize_equals_pillow_and_the_oracle():
    from deephumor_b200.runtime import ops
    imgs = [image(h, w, 11 * i, ('noise', 'extremes', 'gradient')[i % 3]) for i, (h, w) in enumerate(SIZES)]
    out = ops.resize_images(imgs).cpu().numpy()
    assert out.shape == (len(SIZES), 3, 224, 224)
    for i, a in enumerate(imgs):
        got = out[i].transpose(1, 2, 0)
        assert np.array_equal(got, resize_bilinear(a)), f'oracle mismatch on size {SIZES[i]}'
        assert np.array_equal(got, pil_resize(a)), f'Pillow mismatch on size {SIZES[i]}'


@pytest.mark.gpu
@pytest.mark.parametrize('out', [50, 256])
def test_cuda_resize_other_output_sizes_and_device_inputs(out):
    from deephumor_b200.runtime import ops
    imgs = [image(h, w, h + out) for h, w in [(300, 500), (40, 30), (1000, 9), (out, out), (2000, 3000)]]
    got = ops.resize_images([torch.from_numpy(a).cuda() for a in imgs], out).cpu().numpy()
    for i, a in enumerate(imgs):
        assert np.array_equal(got[i].transpose(1, 2, 0), pil_resize(a, out))
    assert ops.resize_images([]).shape == (0, 3, 224, 224)


@pytest.mark.gpu
def test_cuda_resize_rejects_bad_inputs():
    from deephumor_b200._lib import DeepHumorLibError
    from deephumor_b200.runtime import ops
    with pytest.raises(ValueError):
        ops.resize_images([np.zeros((10, 10), dtype=np.uint8)])
    with pytest.raises(DeepHumorLibError):
        ops.resize_images([np.zeros((20000, 2, 3), dtype=np.uint8)])          # 20000 / 224 > 79: beyond the tap table


@pytest.mark.gpu
def test_raw_photos_generate_the_same_captions_as_the_reference_preprocessing():
    """End of the f2 row: raw uint8 photos of arbitrary size -> resize kernel -> fused uint8 stem -> captions, against the
    reference's host chain (PIL Resize + ToTensor + Normalize, nb cell 11) feeding the same model: identical ids."""
    import torchvision.transforms as T
    from deephumor_b200 import models
    from deephumor_b200.utils import synth_weights
    hp = synth_weights.default_hp('lstm', 1000, small=True)
    sd = synth_weights.make_state_dict('lstm', hp, seed=1)
    m = models.CaptioningLSTM(**hp)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval().set_precision('bf16')
    photos = [image(h, w, 5 + i, 'gradient' if i % 2 else 'noise') for i, (h, w) in enumerate([(480, 640), (300, 200), (224, 224), (1000, 50)])]
    tf = T.Compose([T.Resize((224, 224)), T.ToTensor(), T.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    ref_in = torch.stack([tf(Image.fromarray(a)) for a in photos]).cuda()
    kw = dict(max_len=10, beam_size=3, top_k=10, noise='injected', seed=3)
    with torch.no_grad():
        ids_ref, len_ref = m.generate(ref_in, **kw)
        ids, lens = m.generate(m.preprocess(photos), **kw)
    assert torch.equal(ids, ids_ref) and torch.equal(lens, len_ref)


@pytest.mark.gpu
def test_preprocess_accepts_pil_images_of_any_mode():
    """model.preprocess (the reference's Image.open + Resize((224,224)) on the device): PIL images in RGB, greyscale and
    palette + alpha modes go through .convert('RGB') like a user of the reference would do before the transform; the result
    equals torchvision's Resize on the converted image, bit for bit."""
    import torchvision.transforms as T
    from deephumor_b200.models import CaptioningLSTM
    rgb = Image.fromarray(image(300, 420, 1))
    grey = Image.fromarray(image(200, 150, 2)[..., 0], mode='L')
    rgba = Image.fromarray(np.concatenate([image(333, 222, 3), np.full((333, 222, 1), 200, np.uint8)], -1), mode='RGBA')
    out = CaptioningLSTM.preprocess([rgb, grey, rgba]).cpu()
    for i, im in enumerate((rgb, grey, rgba)):
        ref = T.PILToTensor()(T.Resize((224, 224))(im.convert('RGB')))
        assert torch.equal(out[i], ref)
