"""Shared plumbing of the API-mirror modules: precision modes, runtime caches, noise options."""
import torch
from torch import nn

from ..runtime import ops

PRECISIONS = {'bf16': torch.bfloat16, 'fp32': torch.float32}
DEFAULT_PRECISION = 'bf16'


def resolve_noise(noise, seed):
    """noise in {None, 'injected', 'deterministic'}; None = the reference's behaviour (random sampling): an
    'injected' stream whose seed is drawn from torch's global CPU generator, so torch.manual_seed() makes
    generation reproducible exactly as it does for the reference's torch.multinomial."""
    if noise is None:
        noise = 'injected'
    if noise not in ops.NOISE:
        raise ValueError(f"noise must be one of {sorted(ops.NOISE)}, got {noise!r}")
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,))) if noise == 'injected' else 0
    return ops.NOISE[noise], int(seed)


class RTModule(nn.Module):
    """nn.Module whose compute runs through packed device runtimes built lazily from its own state_dict."""

    precision = DEFAULT_PRECISION

    def _rt_cache(self):
        if '_rt_store' not in self.__dict__:
            self.__dict__['_rt_store'] = {}
        return self.__dict__['_rt_store']

    def invalidate(self):
        """Drop packed weights (call after mutating parameters in place)."""
        self._rt_cache().clear()
        for m in self.children():
            if isinstance(m, RTModule):
                m.invalidate()

    def load_state_dict(self, *a, **kw):
        r = super().load_state_dict(*a, **kw)
        self.invalidate()
        return r

    def _apply(self, fn, *a, **kw):
        r = super()._apply(fn, *a, **kw)
        self.invalidate()
        return r

    def _device(self):
        if self.training:
            raise RuntimeError('deephumor_b200 modules implement the eval-mode path only (folded BatchNorm statistics, no dropout, '
                               'no autograd): call .eval() first; training through them is not supported')
        dev = next(self.parameters()).device
        if dev.type != 'cuda':
            raise RuntimeError('deephumor_b200 models run on a CUDA (sm_100a) device only: call .cuda() first '
                               '(there is no CPU fallback)')
        return dev

    def _get_rt(self, key, build):
        cache = self._rt_cache()
        k = (key, self.precision, str(self._device()))
        if k not in cache:
            cache[k] = build(PRECISIONS[self.precision], self._device())
        return cache[k]

    @staticmethod
    def preprocess(images, size=224):
        """The reference's image transform up to ToTensor (deephumor_demo.ipynb cell 11: `Resize((224, 224))` on PIL
        images; data/datasets.py:48-53,94-98) on the device, bit-exact with Pillow: a list of PIL images or uint8 [H,W,3]
        RGB arrays of any sizes -> uint8 [n,3,size,size] (NCHW) on the current CUDA device.  Pass the result to
        `generate` / `forward` / `perplexity`: ToTensor + Normalize are fused into the stem kernel."""
        import numpy as np
        from ..runtime import ops
        arrs = [np.asarray(im.convert('RGB')) if hasattr(im, 'convert') else im for im in images]
        return ops.resize_images(arrs, size)

    def set_precision(self, precision):
        assert precision in PRECISIONS
        self.precision = precision
        for m in self.modules():
            if isinstance(m, RTModule):
                m.precision = precision
        return self


def prefixed(module, p):
    return {f'{p}.{k}': v.detach() for k, v in module.state_dict().items()}


def as_caption(caption, device):
    if caption is None:
        return None
    c = caption.to(device=device, dtype=torch.int32)
    if c.dim() == 1:
        c = c.unsqueeze(0)
    return c.contiguous() if c.shape[1] > 0 else None


def finish(ids, lens, status, n):
    """Turn the device outputs into the reference's return convention and raise on device-side errors."""
    st = int(status.item())
    if st & 1:
        raise RuntimeError('invalid multinomial distribution (sum of probabilities <= 0)')   # torch.multinomial text
    if st & 2:
        raise RuntimeError('more than 4096 logits tie at the top-k threshold')
    if n == 1:
        return ids[0, :int(lens[0])]
    return ids, lens
