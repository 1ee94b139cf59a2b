"""Generates tests/golden/ftcn_golden.npz by running the UNMODIFIED reference FTCN-TT plugin
(`i3d_temporal_var_fix_dropout_tt_cfg`, setting/ftcn_tt.yaml) from /root/reference in the build container and
checks oracle/ftcn_oracle.py against it.  Run (own process: the reference config is a frozen singleton):
    python tests/golden/make_golden_ftcn.py

Environment shim (not a change to the reference): the plugin copies every name of nn.Conv3d's signature off the
existing modules (i3d_temporal_var_fix_dropout_tt_cfg.py:198,238); torch >= 1.9 added `device` and `dtype`, which
are not attributes, so it only constructs under the README's torch 1.8 — or with the two class attributes below.
"""
import os
import sys

import numpy as np
import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import afb200  # noqa: E402,F401
from afb200 import network, synthetic  # noqa: E402
from oracle import ftcn_oracle, ref_loader  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
STAGE_SAMPLES = 4096
N_CLIPS = 3


def stage_sample_index(numel, n=STAGE_SAMPLES):
    return (np.arange(n, dtype=np.int64) * 2654435761 + 12345) % numel


def main():
    nn.Conv3d.device = None
    nn.Conv3d.dtype = None
    clf = ref_loader.reference_classifier("ftcn_tt.yaml")
    ref_keys = {k: tuple(v.shape) for k, v in clf.network.state_dict().items()}
    assert ref_keys == network.reference_key_set("ftcn_tt"), "FTCNTTParams does not reproduce the reference schema"
    sd = synthetic.synthetic_state_dict(0, "ftcn_tt")
    print("loaded synthetic weights into the reference network:", clf.network.load_state_dict(sd, strict=True))
    u8 = np.stack([synthetic.synthetic_clip_u8(i) for i in range(N_CLIPS)])
    x = synthetic.normalise_clip(u8)
    stage_out, tokens, cls = {}, [], []
    hooks = []
    for name in ("s1", "s2", "s3", "s4"):
        hooks.append(getattr(clf.network.resnet, name).register_forward_hook(
            lambda m, i, o, name=name: stage_out.setdefault(name, []).append(o[0].detach())))
    tt = clf.network.resnet.head.time_T
    hooks.append(tt.register_forward_hook(lambda m, i, o: tokens.append(i[0].detach().clone())))
    hooks.append(tt.mlp_head[1].register_forward_hook(lambda m, i, o: cls.append(i[0].detach())))
    logits = []
    with torch.no_grad():
        for i in range(N_CLIPS):
            logits.append(clf(x[i:i + 1])["final_output"])
    for h in hooks:
        h.remove()
    logits = torch.cat(logits).numpy()
    print("reference logits", logits.ravel())
    o_logits, o_stages = ftcn_oracle.forward(sd, x, return_stages=True)
    d = np.abs(o_logits.numpy() - logits).max()
    print("oracle vs reference  max|dlogit| = %.3e" % d)
    assert d <= 2e-5, d
    out = {"logits": logits.astype(np.float32), "weights_seed": np.int64(0), "clip_indices": np.arange(N_CLIPS)}
    for si, name in enumerate(("s1", "s2", "s3", "s4")):
        ref = torch.cat(stage_out[name]).numpy()
        orc = o_stages[si].numpy()
        rel = np.abs(ref - orc).max() / max(np.abs(ref).max(), 1e-9)
        print("  stage %s shape %s absmax %.3f mean|x| %.4f  oracle rel err %.2e" % (
            name, ref.shape, np.abs(ref).max(), np.abs(ref).mean(), rel))
        assert rel <= 1e-5
        idx = stage_sample_index(ref.size)
        out[name + "_samples"] = ref.ravel()[idx].astype(np.float32)
        out[name + "_shape"] = np.array(ref.shape)
    ref_tok = torch.cat(tokens).numpy()
    # the hook sees the input AFTER TimeTransformer.forward's in-place `x += pos_embedding` on the concatenated copy,
    # i.e. the original token tensor is untouched: compare directly
    d = np.abs(o_stages[4].numpy() - ref_tok).max()
    print("  tokens %s  oracle max abs err %.2e" % (ref_tok.shape, d))
    assert d <= 1e-5, d
    ref_cls = torch.cat(cls).numpy()
    d = np.abs(o_stages[5].numpy() - ref_cls).max()
    print("  cls (input of mlp_head.1) %s  oracle max abs err %.2e" % (ref_cls.shape, d))
    assert d <= 2e-5, d
    out["tokens"] = ref_tok.astype(np.float32)
    out["cls"] = ref_cls.astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "ftcn_golden.npz"), **out)
    print("wrote ftcn_golden.npz")


if __name__ == "__main__":
    main()
