"""Host logic of the batched caller-side API on CPU: bucketing, padding, order and trimming
of infer_batch, with a stand-in module that has the reference's .infer() signature (no compute path is exercised)."""
import types

import numpy as np
import torch

from bert_vits2_b200.infer_api import infer_batch, pad_items


class _FakeNet(torch.nn.Module):
    """Emits, per utterance, 2 frames per phone; sample value = phone id of the frame's token (+ 0.5 where the BERT row 0
    is positive) so that order, padding and trimming are all visible in the output."""

    def __init__(self):
        super().__init__()
        self.p = torch.nn.Parameter(torch.zeros(1))
        self.cfg = types.SimpleNamespace(hop=4)
        self.calls = []

    def infer(self, x, x_lengths, sid, tone, language, bert, ja_bert, en_bert, **kw):
        B, T = x.shape
        self.calls.append((B, T, sorted(x_lengths.tolist())))
        assert bert.shape == (B, 8, T) and tone.shape == (B, T) and sid.shape == (B,)
        F = 2 * T
        tok = torch.arange(F) // 2
        val = x[:, tok].float() + 0.5 * (bert[:, 0, tok] > 0).float()
        o = val.repeat_interleave(self.cfg.hop, dim=1).unsqueeze(1)           # [B,1,F*hop]
        y_mask = (torch.arange(F)[None, :] < (2 * x_lengths)[:, None]).float().unsqueeze(1)
        return o, None, y_mask, None


def _item(t, seed):
    g = torch.Generator().manual_seed(seed)
    ph = torch.randint(1, 100, (t,), generator=g)
    return (torch.randn(8, t, generator=g), torch.randn(8, t, generator=g), torch.randn(8, t, generator=g), ph,
            torch.randint(0, 5, (t,), generator=g), torch.zeros(t, dtype=torch.int64))


def test_pad_items_shapes_and_zero_padding():
    items = [_item(5, 0), _item(9, 1), _item(7, 2)]
    d = pad_items(items, "cpu")
    assert d["x"].shape == (3, 9) and d["bert"].shape == (3, 8, 9) and d["x_lengths"].tolist() == [5, 9, 7]
    assert (d["x"][0, 5:] == 0).all() and (d["bert"][0, :, 5:] == 0).all() and torch.equal(d["x"][1], items[1][3])


def test_infer_batch_order_bucketing_and_trimming():
    lens = [11, 3, 7, 3, 12, 6, 1]
    items = [_item(t, 10 + i) for i, t in enumerate(lens)]
    net = _FakeNet()
    outs = infer_batch(net, items, sid=3, batch_size=3)
    assert len(outs) == len(items) and len(net.calls) == 3
    # buckets hold utterances of similar length (sorted dealing), every utterance exactly once
    assert sorted(sum((c[2] for c in net.calls), [])) == sorted(lens)
    assert all(max(c[2]) == c[1] for c in net.calls)
    for it, t, o in zip(items, lens, outs):
        assert isinstance(o, np.ndarray) and o.dtype == np.float32 and o.shape == (2 * t * 4,)
        expect = (it[3].float() + 0.5 * (it[0][0] > 0).float()).repeat_interleave(8).numpy()
        assert np.array_equal(o, expect)
