"""Drop-in check against the REAL reference tree (build container only; skipped where /root/reference is
absent, e.g. on the GPU box): follow INTEGRATION.md §2 literally — copy plugin/i3d_b200.py into a scratch copy
of the reference's `model/classifier/`, point a setting yaml at it, and let the reference's own PluginLoader,
config and ModelBase build, load and inspect it.  No GPU work: the engine is created lazily on first forward."""
import os
import shutil
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/altfreezing"

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "slowfast")), reason="reference tree not present")


def test_reference_plugin_loader_builds_and_loads_our_plugin(tmp_path):
    work = tmp_path / "altfreezing"
    for sub in ("model", "utils", "slowfast", "setting", "trainer"):
        shutil.copytree(os.path.join(REF, sub), work / sub)
    for f in ("config.py", "root_setting.yaml"):
        shutil.copy(os.path.join(REF, f), work / f)
    pkg = os.path.join(ROOT, "spatiotemporal-deepfake-detection-for-live-video-calls_b200")
    shutil.copy(os.path.join(pkg, "plugin", "i3d_b200.py"), work / "model" / "classifier" / "i3d_b200.py")
    yaml_src = (work / "setting" / "i3d_ori.yaml").read_text()
    assert "classifier_type: i3d_ori" in yaml_src
    (work / "setting" / "i3d_b200.yaml").write_text(yaml_src.replace("classifier_type: i3d_ori", "classifier_type: i3d_b200"))
    script = textwrap.dedent("""
        import sys, os
        sys.path.insert(0, %r)                       # repo root: oracle.ref_loader stubs fvcore/simplejson/termcolor
        from oracle import ref_loader
        ref_loader._install_stubs()
        sys.path.insert(0, %r)
        os.environ["AFB200_ROOT"] = %r
        import torch
        from config import config as cfg
        cfg.init_with_yaml(); cfg.update_with_yaml("i3d_b200.yaml"); cfg.freeze()
        from utils.plugin_loader import PluginLoader
        from model._base import ModelBase
        clf = PluginLoader.get_classifier(cfg.classifier_type)().eval()
        assert isinstance(clf, ModelBase), type(clf)
        import afb200
        assert isinstance(clf._warped_network, afb200.B200Engine)
        keys = set(clf.network.state_dict())
        assert keys == set(afb200.network.reference_key_set()), len(keys)
        # the wrapper's own state_dict() is key-for-key the reference wrapper's (which registers the network twice,
        # model/_base.py:22-23), so checkpoints saved from either load into the other with strict=True
        ref_clf = PluginLoader.get_classifier("i3d_ori")()
        assert list(clf.state_dict().keys()) == list(ref_clf.state_dict().keys())
        clf.load_state_dict(ref_clf.state_dict(), strict=True)
        sd = afb200.synthetic.synthetic_state_dict(0)
        path = os.path.join(%r, "ckpt.pth")
        torch.save({"state_dict": {"module." + k: v for k, v in sd.items()}}, path)
        ok, epoch = clf.load(path)
        assert ok, "reference ModelBase.load failed"
        k = "resnet.s4.pathway0_res2.branch2.b.weight"
        assert torch.equal(clf.network.state_dict()[k], sd[k])
        assert sum(p.numel() for p in clf.parameters()) == 27225921
        try:
            clf(torch.zeros(1, 3, 32, 224, 224))
            raise SystemExit("CPU forward should have failed loudly")
        except RuntimeError as e:
            assert "no CPU fallback" in str(e)
        print("PLUGIN-OK")
    """) % (ROOT, str(work), ROOT, str(tmp_path))
    r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "PLUGIN-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_reference_plugin_loader_builds_and_loads_our_ftcn_tt_plugin(tmp_path):
    """Same drop-in check for the second classifier plugin (FTCN-TT, setting/ftcn_tt.yaml)."""
    work = tmp_path / "altfreezing"
    for sub in ("model", "utils", "slowfast", "setting", "trainer"):
        shutil.copytree(os.path.join(REF, sub), work / sub)
    for f in ("config.py", "root_setting.yaml"):
        shutil.copy(os.path.join(REF, f), work / f)
    pkg = os.path.join(ROOT, "spatiotemporal-deepfake-detection-for-live-video-calls_b200")
    shutil.copy(os.path.join(pkg, "plugin", "ftcn_tt_b200.py"), work / "model" / "classifier" / "ftcn_tt_b200.py")
    yaml_src = (work / "setting" / "ftcn_tt.yaml").read_text()
    old = "classifier_type: i3d_temporal_var_fix_dropout_tt_cfg"
    assert old in yaml_src
    (work / "setting" / "ftcn_tt_b200.yaml").write_text(yaml_src.replace(old, "classifier_type: ftcn_tt_b200"))
    script = textwrap.dedent("""
        import sys, os
        sys.path.insert(0, %r)
        from oracle import ref_loader
        ref_loader._install_stubs()
        sys.path.insert(0, %r)
        os.environ["AFB200_ROOT"] = %r
        import torch
        from torch import nn
        nn.Conv3d.device = None; nn.Conv3d.dtype = None      # torch >= 1.9 shim the reference module needs (plugin docstring)
        from config import config as cfg
        cfg.init_with_yaml(); cfg.update_with_yaml("ftcn_tt_b200.yaml"); cfg.freeze()
        from utils.plugin_loader import PluginLoader
        from model._base import ModelBase
        clf = PluginLoader.get_classifier(cfg.classifier_type)().eval()
        assert isinstance(clf, ModelBase), type(clf)
        import afb200
        eng = clf._warped_network
        assert isinstance(eng, afb200.B200Engine) and eng.variant == "ftcn_tt"
        keys = {k: tuple(v.shape) for k, v in clf.network.state_dict().items()}
        assert keys == afb200.network.reference_key_set("ftcn_tt"), len(keys)
        sd = afb200.synthetic.synthetic_state_dict(0, "ftcn_tt")
        path = os.path.join(%r, "ckpt.pth")
        torch.save({"state_dict": {"module." + k: v for k, v in sd.items()}}, path)
        ok, epoch = clf.load(path)
        assert ok, "reference ModelBase.load failed"
        k = "resnet.head.time_T.transformer.layers.0.0.fn.fn.to_qkv.weight"
        assert torch.equal(clf.network.state_dict()[k], sd[k])
        # the host-side fold of the loaded reference network gives the C-ABI structures of the FTCN-TT variant
        fw = afb200.weights.FoldedWeights(clf.network.state_dict(), cfg.clip_size, cfg.imsize, "ftcn_tt")
        assert fw.struct.n_convs == 43 and fw.struct.stem_pool2 == 1 and fw.struct.tt_head.contents.dim == 1024
        try:
            clf(torch.zeros(1, 3, 32, 224, 224))
            raise SystemExit("CPU forward should have failed loudly")
        except RuntimeError as e:
            assert "no CPU fallback" in str(e)
        print("PLUGIN-OK")
    """) % (ROOT, str(work), ROOT, str(tmp_path))
    r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "PLUGIN-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
