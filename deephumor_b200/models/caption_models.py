"""The four captioners with the reference's constructor / forward / generate / save / from_pretrained
signatures, ``_hp`` dicts and ``state_dict`` layout (deephumor/models/caption_models.py:9-461).

Extensions (keyword-only, reference-preserving defaults): ``image`` may carry N > 1 images, in which case
``generate`` returns ``(ids [N,max_len] padded with pad_index, lengths [N])`` -- the reference is batch-1 only
 and for N == 1 the return value has exactly its shape; ``noise``/``seed`` choose the shared
noise model (``None`` = random sampling like the reference); ``image_base`` is the global index of image 0
(data-parallel sharding); ``set_precision('bf16'|'fp32')`` picks tensor-core or fp32 check mode.
"""
import torch

from ._base import RTModule
from .encoders import ImageEncoder, ImageLabelEncoder
from .rnn_models import LSTMDecoder
from .transformers import SelfAttentionTransformerDecoder, TransformerDecoder


class _Captioner(RTModule):
    def save(self, ckpt_path):
        """Saves the model's state and hyperparameters (caption_models.py:76-81)."""
        torch.save({'model': self.state_dict(), 'hp': self._hp}, ckpt_path)

    @classmethod
    def _from_pretrained(cls, ckpt_path):
        ckpt = torch.load(ckpt_path, map_location='cpu')
        model = cls(**ckpt['hp'])
        model.load_state_dict(ckpt['model'])
        return model

    def perplexity(self, images, captions, lengths, labels=None, pad_index=0):
        """`perplexity(self(images, captions[:, :-1], lengths[, labels])[:, :T], captions, lengths)` of the reference's
        evaluation loop (experiments/trainer.py:66-81, experiments/metrics.py:4-9) with the log-softmax and target gather
        fused into the classifier contraction: the [bs, T, V] logits (9.6 GB for config 3) are never materialised."""
        enc = self.encoder(images, labels) if labels is not None else self.encoder(images)
        lp = self._token_logprob(enc, captions[:, :-1], lengths, captions)
        T = lp.shape[1]
        tg = captions[:, :T].to(lp.device)
        lp = lp / lengths.to(lp.device).unsqueeze(1)           # divide by lengths BEFORE masking (Q27)
        lp = lp.masked_fill(tg == pad_index, 0.)
        return (-lp.sum(dim=-1)).exp().mean()

    def _gen_kwargs(self, max_len, temperature, beam_size, top_k, eos_index, noise, seed, image_base):
        return dict(max_len=max_len, temperature=temperature, beam_size=beam_size, top_k=top_k, eos_index=eos_index,
                    noise=noise, seed=seed, image_base=image_base)


class CaptioningLSTM(_Captioner):
    def __init__(self, num_tokens, emb_dim=256, hidden_size=512, num_layers=2, enc_dropout=0.3, dec_dropout=0.1):
        super().__init__()
        self.encoder = ImageEncoder(emb_dim=emb_dim, dropout=enc_dropout)
        self.decoder = LSTMDecoder(num_tokens=num_tokens, emb_dim=emb_dim, hidden_size=hidden_size,
                                   num_layers=num_layers, dropout=dec_dropout)
        self._hp = {'num_tokens': num_tokens, 'emb_dim': emb_dim, 'hidden_size': hidden_size,
                    'num_layers': num_layers, 'enc_dropout': enc_dropout, 'dec_dropout': dec_dropout}

    def forward(self, images, captions, lengths=None):
        return self.decoder(self.encoder(images), captions, lengths)

    def _token_logprob(self, enc, captions_in, lengths, targets):
        return self.decoder.token_logprob(enc, captions_in, lengths, targets)

    def generate(self, image, caption=None, max_len=25, temperature=1.0, beam_size=10, top_k=50, eos_index=3,
                 *, noise=None, seed=None, image_base=0):
        image_emb = self.encoder(image).unsqueeze(1)
        return self.decoder.generate(image_emb, caption=caption,
                                     **self._gen_kwargs(max_len, temperature, beam_size, top_k, eos_index, noise, seed,
                                                        image_base))

    @staticmethod
    def from_pretrained(ckpt_path):
        return CaptioningLSTM._from_pretrained(ckpt_path)


class CaptioningLSTMWithLabels(_Captioner):
    def __init__(self, num_tokens, emb_dim=256, hidden_size=512, num_layers=2, enc_dropout=0.3, dec_dropout=0.1):
        super().__init__()
        self.encoder = ImageLabelEncoder(num_tokens=num_tokens, emb_dim=emb_dim, dropout=enc_dropout)
        self.decoder = LSTMDecoder(num_tokens=num_tokens, emb_dim=emb_dim, hidden_size=hidden_size,
                                   num_layers=num_layers, dropout=dec_dropout,
                                   embedding=self.encoder.label_encoder.embedding)     # tied (Q22)
        self._hp = {'num_tokens': num_tokens, 'emb_dim': emb_dim, 'hidden_size': hidden_size,
                    'num_layers': num_layers, 'enc_dropout': enc_dropout, 'dec_dropout': dec_dropout}

    def forward(self, images, captions, lengths, labels):
        return self.decoder(self.encoder(images=images, labels=labels), captions, lengths)

    def _token_logprob(self, enc, captions_in, lengths, targets):
        return self.decoder.token_logprob(enc, captions_in, lengths, targets)

    def generate(self, image, label, caption=None, max_len=25, temperature=1.0, beam_size=10, top_k=50, eos_index=3,
                 *, noise=None, seed=None, image_base=0):
        image_emb = self.encoder(image, label).unsqueeze(1)
        return self.decoder.generate(image_emb, caption=caption,
                                     **self._gen_kwargs(max_len, temperature, beam_size, top_k, eos_index, noise, seed,
                                                        image_base))

    @staticmethod
    def from_pretrained(ckpt_path):
        return CaptioningLSTMWithLabels._from_pretrained(ckpt_path)


class CaptioningTransformerBase(_Captioner):
    def __init__(self, num_tokens, hid_dim=512, n_layers=6, n_heads=8, pf_dim=2048, enc_dropout=0.3, dec_dropout=0.1,
                 pad_index=0, max_len=128):
        super().__init__()
        self.encoder = ImageEncoder(emb_dim=hid_dim, dropout=enc_dropout, spatial_features=False)
        self.decoder = SelfAttentionTransformerDecoder(num_tokens=num_tokens, hid_dim=hid_dim, n_layers=n_layers,
                                                       n_heads=n_heads, pf_dim=pf_dim, dropout=dec_dropout,
                                                       pad_index=pad_index, max_len=max_len)
        self._hp = {'num_tokens': num_tokens, 'hid_dim': hid_dim, 'n_layers': n_layers, 'n_heads': n_heads,
                    'pf_dim': pf_dim, 'enc_dropout': enc_dropout, 'dec_dropout': dec_dropout, 'pad_index': pad_index,
                    'max_len': max_len}

    def forward(self, images, captions, lengths=None):
        return self.decoder(captions, start_emb=self.encoder(images))

    def _token_logprob(self, enc, captions_in, lengths, targets):
        return self.decoder.token_logprob(captions_in, None, enc, targets)

    def generate(self, image, caption=None, max_len=25, temperature=1.0, beam_size=10, top_k=50, eos_index=3,
                 *, noise=None, seed=None, image_base=0):
        return self.decoder.generate(self.encoder(image), caption=caption,
                                     **self._gen_kwargs(max_len, temperature, beam_size, top_k, eos_index, noise, seed,
                                                        image_base))

    @staticmethod
    def from_pretrained(ckpt_path):
        return CaptioningTransformerBase._from_pretrained(ckpt_path)


class CaptioningTransformer(_Captioner):
    def __init__(self, num_tokens, hid_dim=512, n_layers=6, n_heads=8, pf_dim=2048, enc_dropout=0.3, dec_dropout=0.1,
                 pad_index=0, max_len=128):
        super().__init__()
        self.encoder = ImageEncoder(emb_dim=hid_dim, dropout=enc_dropout, spatial_features=True)
        self.decoder = TransformerDecoder(num_tokens=num_tokens, hid_dim=hid_dim, n_layers=n_layers, n_heads=n_heads,
                                          pf_dim=pf_dim, dropout=dec_dropout, pad_index=pad_index, max_len=max_len)
        self._hp = {'num_tokens': num_tokens, 'hid_dim': hid_dim, 'n_layers': n_layers, 'n_heads': n_heads,
                    'pf_dim': pf_dim, 'enc_dropout': enc_dropout, 'dec_dropout': dec_dropout, 'pad_index': pad_index,
                    'max_len': max_len}

    def forward(self, images, captions, lengths=None):
        image_emb, image_spatial_emb = self.encoder(images)
        return self.decoder(captions, enc_out=image_spatial_emb, start_emb=image_emb)

    def _token_logprob(self, enc, captions_in, lengths, targets):
        return self.decoder.token_logprob(captions_in, enc[1], enc[0], targets)

    def generate(self, image, caption=None, max_len=25, temperature=1.0, beam_size=10, top_k=50, eos_index=3,
                 *, noise=None, seed=None, image_base=0):
        image_emb, image_spatial_emb = self.encoder(image)
        return self.decoder.generate(image_emb, image_spatial_emb, caption=caption,
                                     **self._gen_kwargs(max_len, temperature, beam_size, top_k, eos_index, noise, seed,
                                                        image_base))

    @staticmethod
    def from_pretrained(ckpt_path):
        return CaptioningTransformer._from_pretrained(ckpt_path)
