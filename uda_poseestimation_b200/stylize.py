"""Forward-only style transfer — drop-in for how the trainers use ``lib/models/Style_net.py::Net``
(``train_human.py:275-276,350-351,355-356``)::

    x_s = style_net(x_s, x_t, _a)[2]
    x_s = torch.maximum(torch.minimum(x_s.permute(0,2,3,1), recover_max), recover_min).permute(0,3,1,2)

The reference's ``Net.forward`` (Style_net.py:163-177) also re-encodes the stylised image and computes a
content MSE and four Gram-matrix style losses — all discarded by the callers, who keep only element ``[2]``.
``StyleTransfer`` keeps the constructor and the ``(loss_c, loss_s, g_t)`` return shape (the two losses are
``None``) and runs only what ``g_t`` needs: the two relu4_1 encodes (cuDNN, one batched pass), the fused
AdaIN + alpha mix (``udape_adain_mix``: one launch instead of ~19 eager passes) and the decoder (cuDNN);
``stylize`` adds the per-channel clamp (``udape_channel_clamp``) and returns a contiguous NCHW image.
The convolutions stay in PyTorch / cuDNN (north star); there is no CPU path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .adain import adain_mix, channel_clamp

__all__ = ["StyleTransfer"]


class StyleTransfer(nn.Module):
    def __init__(self, encoder, decoder):
        super().__init__()
        enc_layers = list(encoder.children())
        self.enc_1 = nn.Sequential(*enc_layers[:4])      # input -> relu1_1      (Style_net.py:123-126)
        self.enc_2 = nn.Sequential(*enc_layers[4:11])    # relu1_1 -> relu2_1
        self.enc_3 = nn.Sequential(*enc_layers[11:18])   # relu2_1 -> relu3_1
        self.enc_4 = nn.Sequential(*enc_layers[18:31])   # relu3_1 -> relu4_1
        self.decoder = decoder
        for name in ("enc_1", "enc_2", "enc_3", "enc_4"):   # fix the encoder (:131-134)
            for param in getattr(self, name).parameters():
                param.requires_grad = False

    def encode(self, input):
        """relu4_1 features (Style_net.py:145-148)."""
        for i in range(4):
            input = getattr(self, f"enc_{i + 1}")(input)
        return input

    def forward(self, content, style, alpha=1.0):
        """``(None, None, g_t)`` — ``g_t`` as Style_net.py:163-170; the discarded losses are not computed."""
        assert 0 <= alpha <= 1
        if content.shape == style.shape:
            # one batched pass through the frozen encoder for both images
            feats = self.encode(torch.cat([content, style], dim=0))
            content_feat, style_feat = feats[: content.shape[0]], feats[content.shape[0]:]
        else:
            content_feat, style_feat = self.encode(content), self.encode(style)
        with torch.no_grad():
            t = adain_mix(content_feat.detach(), style_feat.detach(), float(alpha))
        return None, None, self.decoder(t)

    def stylize(self, content, style, alpha, recover_min, recover_max):
        """The trainers' two lines in one call: stylised image clamped per channel to the normalised pixel
        range (``recover_min`` / ``recover_max``, train_human.py:32-33), contiguous NCHW."""
        with torch.no_grad():
            g_t = self.forward(content, style, alpha)[2]
            return channel_clamp(g_t, recover_min, recover_max, out=g_t if g_t.is_contiguous() else None)
