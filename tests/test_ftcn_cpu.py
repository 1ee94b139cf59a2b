"""CPU tests of the FTCN-TT plugin path: the oracle against the golden vectors generated
from the unmodified reference plugin, the reference state_dict schema, and the folded weights / C-ABI structures."""
import os

import numpy as np
import torch

import afb200
from afb200 import arch, network, synthetic
from afb200.weights import FoldedWeights
from oracle import ftcn_oracle
from tests.helpers import stage_sample_index

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _golden():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "ftcn_golden.npz")))


def test_ftcn_oracle_matches_reference_golden():
    g = _golden()
    sd = synthetic.synthetic_state_dict(int(g["weights_seed"]), "ftcn_tt")
    idx = [1]
    u8 = np.stack([synthetic.synthetic_clip_u8(i) for i in idx])
    logits, stages = ftcn_oracle.forward(sd, synthetic.normalise_clip(u8), return_stages=True)
    assert np.abs(logits.numpy() - g["logits"][idx]).max() <= 2e-5
    assert np.abs(stages[4].numpy() - g["tokens"][idx]).max() <= 2e-5
    assert np.abs(stages[5].numpy() - g["cls"][idx]).max() <= 5e-5
    for si, name in enumerate(("s1", "s2", "s3", "s4")):
        shape = g[name + "_shape"]
        per_clip = int(np.prod(shape[1:]))
        gi = stage_sample_index(int(np.prod(shape)))
        clip_of = gi // per_clip
        ours = stages[si].numpy().reshape(len(idx), -1)
        for j, c in enumerate(idx):
            sel = clip_of == c
            assert sel.sum() > 100
            want = g[name + "_samples"][sel]
            assert np.abs(ours[j][gi[sel] - c * per_clip] - want).max() <= 1e-5 * max(1.0, np.abs(want).max()), name


def test_ftcn_schema_and_arch_table():
    keys = network.reference_key_set("ftcn_tt")
    assert len(keys) == 275                      # what the reference plugin's state_dict holds (make_golden_ftcn.py)
    sd = synthetic.synthetic_state_dict(0, "ftcn_tt")
    assert set(sd) == set(keys)
    assert all(tuple(sd[k].shape) == shp for k, shp in keys.items())
    network.FTCNTTParams().load_state_dict(sd, strict=True)
    specs = arch.all_conv_specs("ftcn_tt")
    assert len(specs) == 1 + 3 * 13 + 3          # stem, 13 blocks x (a,b,c), 3 projection shortcuts
    assert all(sp.kernel[1:] == (1, 1) and sp.stride == (1, 1, 1) for sp in specs)
    pooled = [sp.name for sp in specs if sp.pool2]
    assert pooled == ["resnet.s1.pathway0_stem.conv", "resnet.s3.pathway0_res0.branch1", "resnet.s3.pathway0_res0.branch2.b",
                      "resnet.s4.pathway0_res0.branch1", "resnet.s4.pathway0_res0.branch2.b"]
    assert all(sp.bn.endswith(".0") == sp.pool2 for sp in specs)


def test_ftcn_folded_weights_structures():
    sd = synthetic.synthetic_state_dict(0, "ftcn_tt")
    fw = FoldedWeights(sd, 32, 224, "ftcn_tt")
    st = fw.struct
    assert st.n_convs == 43 and st.n_blocks == 13 and st.stem_pool2 == 1 and st.feature_dim == 1024
    assert [fw.blocks[i].spatial_pool for i in range(13)] == [0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0, 0]
    assert [fw.blocks[i].temporal_pool_before for i in range(13)] == [0, 0, 0, 1] + [0] * 9
    h = st.tt_head.contents
    assert (h.dim, h.tokens, h.heads, h.dim_head, h.mlp_dim, h.depth) == (1024, 16, 16, 64, 2048, 1)
    # the i3d variant leaves the FTCN-only fields zero
    fi = FoldedWeights(synthetic.synthetic_state_dict(0), 32, 224).struct
    assert fi.stem_pool2 == 0 and not fi.tt_head and fi.n_blocks == 16


def test_ftcn_classifier_container_loads_reference_schema_checkpoint(tmp_path):
    sd = synthetic.synthetic_state_dict(3, "ftcn_tt")
    p = tmp_path / "ftcn.pth"
    torch.save({"state_dict": {"module." + k: v for k, v in sd.items()}}, p)
    clf = afb200.Classifier(variant="ftcn_tt")
    ok, epoch = clf.load(str(p), epoch=7)
    assert ok and epoch == 7
    own = clf.network.state_dict()
    assert all(torch.equal(own[k], sd[k]) for k in sd)
