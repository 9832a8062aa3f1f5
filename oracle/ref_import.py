"""Import the UNMODIFIED reference (fishaudio/Bert-VITS2 @ /root/reference) in this container.

TEST INFRASTRUCTURE ONLY.  Used by tests/golden/make_golden.py and oracle/validate_against_reference.py
to pin the oracle restatement (oracle/vits2_oracle.py) against the real reference.  /root/reference does
not exist on the GPU box, so nothing imported by `-m gpu` tests, smoke() or bench.py may import this.

Recipe: `import models` executes `from text import symbols, num_tones, num_languages`
(models.py:15) and text/__init__.py:62-63 runs check_bert_models() at import (needs config.yml and the
network).  We pre-register a stub `text` package that only exposes text/symbols.py, loaded by path.
No reference file is modified or copied.
"""
import importlib.util
import os
import sys
import types

REF = os.environ.get("BV2_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "models.py"))


def import_reference():
    """Returns (models, utils, commons, hps) from the unmodified reference."""
    if not available():
        raise RuntimeError(f"reference not present at {REF}")
    if "models" in sys.modules and getattr(sys.modules["models"], "__bv2_ref__", False):
        m = sys.modules["models"]
        return m, sys.modules["utils"], sys.modules["commons"], m.__bv2_hps__
    spec = importlib.util.spec_from_file_location("text.symbols", os.path.join(REF, "text", "symbols.py"))
    symmod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(symmod)
    text = types.ModuleType("text")
    text.__path__ = [os.path.join(REF, "text")]
    for k in dir(symmod):
        if not k.startswith("__"):
            setattr(text, k, getattr(symmod, k))
    sys.modules["text"] = text
    sys.modules["text.symbols"] = symmod
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import models  # noqa: E402  (the reference's models.py)
    import utils  # noqa: E402
    import commons  # noqa: E402
    hps = utils.get_hparams_from_file(os.path.join(REF, "configs", "config.json"))
    models.__bv2_ref__ = True
    models.__bv2_hps__ = hps
    return models, utils, commons, hps


def build_reference_net(use_transformer_flow=True, **model_overrides):
    """SynthesizerTrn exactly as infer.get_net_g builds it (infer.py:95-101), on CPU, eval()."""
    models, utils, commons, hps = import_reference()
    from text.symbols import symbols
    kw = dict(hps.model)
    if not use_transformer_flow:
        kw["use_transformer_flow"] = False
    kw.update(model_overrides)
    net = models.SynthesizerTrn(
        len(symbols),
        hps.data.filter_length // 2 + 1,
        hps.train.segment_size // hps.data.hop_length,
        n_speakers=hps.data.n_speakers,
        **kw,
    )
    net.eval()
    return net, hps
