"""GPU parity tests of the FTCN-TT plugin path against oracle/ftcn_oracle.py and the reference-generated golden
vectors: fp32 engine within 1e-3 on logits, bf16 tensor-core engine within 2e-2 (the north-star gates)."""
import os

import numpy as np
import pytest
import torch

import afb200
from afb200 import synthetic
from oracle import ftcn_oracle
from tests.helpers import stage_sample_index

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


@pytest.fixture(scope="module")
def golden():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "ftcn_golden.npz")))


@pytest.fixture(scope="module")
def sd():
    return synthetic.synthetic_state_dict(0, "ftcn_tt")


@pytest.fixture(scope="module")
def clips():
    return np.stack([synthetic.synthetic_clip_u8(i) for i in range(3)])


@pytest.fixture(scope="module")
def oracle_out(sd, clips):
    x = synthetic.normalise_clip(clips)
    logits, stages = ftcn_oracle.forward(sd, x, return_stages=True)
    return x, logits, stages


def test_ftcn_fp32_matches_oracle_and_golden(dev, sd, oracle_out, golden):
    x, o_logits, o_stages = oracle_out
    eng = afb200.Engine(sd, max_batch=3, precision="fp32", variant="ftcn_tt")
    eng.set_option("keep_stages", 1)
    logits, feats = eng.forward(x.to(dev), return_features=True)
    assert (logits.cpu() - o_logits).abs().max().item() <= 1e-3
    assert np.abs(logits.cpu().numpy() - golden["logits"]).max() <= 1e-3
    assert np.abs(feats.cpu().numpy() - golden["cls"]).max() <= 2e-3        # LayerNorm'd cls token, O(1) values
    for si, name in enumerate(("s1", "s2", "s3", "s4")):
        got = eng.get_stage(si + 1).cpu()
        assert tuple(got.shape) == tuple(golden[name + "_shape"])
        idx = stage_sample_index(got.numel())
        assert np.abs(got.reshape(-1).numpy()[idx] - golden[name + "_samples"]).max() <= 1e-4
        assert (got - o_stages[si]).abs().max().item() <= 1e-4
    _, tokens = eng.forward_frames(x.to(dev))
    assert np.abs(tokens.cpu().numpy() - golden["tokens"]).max() <= 1e-4
    eng.close()


def test_ftcn_bf16_within_tolerance(dev, sd, oracle_out, clips):
    x, o_logits, o_stages = oracle_out
    eng = afb200.Engine(sd, max_batch=3, precision="bf16", variant="ftcn_tt")
    eng.set_option("keep_stages", 1)
    logits = eng.forward(x.to(dev)).cpu()
    assert (logits - o_logits).abs().max().item() <= 2e-2
    for si in range(4):
        got = eng.get_stage(si + 1).cpu()
        rel = ((got - o_stages[si]).norm() / o_stages[si].norm()).item()
        assert rel <= 2e-2, (si, rel)
    assert eng.launch_count >= 40                      # our kernels ran
    eng.close()
    # production schedule (fused temporal pool etc.), u8 clips through the ClassifierSvc-style entry, odd batch
    eng2 = afb200.Engine(sd, max_batch=2, precision="bf16", variant="ftcn_tt")
    lg, sc = eng2.infer_u8(torch.from_numpy(clips).to(dev))
    assert (lg.cpu() - o_logits.view(-1)).abs().max().item() <= 2e-2
    assert torch.allclose(sc.cpu(), torch.sigmoid(lg.cpu()), atol=1e-6)
    eng2.close()


def test_ftcn_classifier_plugin_interface(dev, sd, oracle_out):
    x, o_logits, _ = oracle_out
    clf = afb200.Classifier(precision="fp32", max_batch=2, variant="ftcn_tt").to(dev).eval()
    clf.load_state_dict_tolerant(sd)
    seen = []
    proj = clf.network.resnet.head.time_T.mlp_head
    hook = getattr(proj, "1").register_forward_hook(lambda m, i, o: seen.append(i[0].detach().cpu()))
    out = clf(x[:2].to(dev))
    hook.remove()
    assert set(out) == {"final_output"} and tuple(out["final_output"].shape) == (2, 1)
    assert (out["final_output"].cpu() - o_logits[:2]).abs().max().item() <= 1e-3
    assert len(seen) == 1 and tuple(seen[0].shape) == (2, 1024)           # the hook saw the last Linear's input
    out2 = clf(x[:2].to(dev))                                            # fused head path (no hook)
    assert (out2["final_output"] - out["final_output"]).abs().max().item() <= 1e-5


def test_ftcn_tensor_core_stem_elementwise(dev, sd, clips):
    """Stage s1 of the bf16 engine (ftcn_stem_umma_kernel: conv k[5,1,1] + BN + MaxPool(1,2,2) + ReLU + MaxPool 3x3/2 on the
    tensor cores) element by element against fp32 torch on the same bf16-rounded clip and folded weights: one bf16 ulp."""
    import torch.nn.functional as F
    eng = afb200.Engine(sd, max_batch=3, precision="bf16", variant="ftcn_tt")
    eng.set_option("keep_stages", 1)
    x = synthetic.normalise_clip(clips)
    eng.forward(x.to(dev))
    got = eng.get_stage(1).cpu()                                            # fp32 NCTHW [3,64,32,56,56]
    w, b = (torch.from_numpy(a) for a in afb200.fold_conv_bn(sd, afb200.arch.stem_spec_for("ftcn_tt")))
    xr = x.to(torch.bfloat16).float()
    y = F.conv3d(xr, w.to(torch.bfloat16).float(), b, 1, (2, 0, 0))
    y = F.relu(F.max_pool3d(y, (1, 2, 2), (1, 2, 2)))
    want = F.max_pool3d(y, (1, 3, 3), (1, 2, 2), (0, 1, 1)).to(torch.bfloat16).float()
    assert got.shape == want.shape
    diff = (got - want).abs()
    tol = 2.0 ** -7 * max(1.0, want.abs().max().item()) + 1e-3
    assert diff.max().item() <= tol, (diff.max().item(), tol)
    assert (diff > 2.0 ** -8 * want.abs().clamp_min(1.0) + 1e-3).float().mean().item() < 1e-3
    # tile borders of the pooled 8x16 tiles, clip-end frames, image border
    assert diff[:, :, :, ::8].max().item() <= tol and diff[:, :, :, :, ::4].max().item() <= tol
    assert diff[:, :, 0].max().item() <= tol and diff[:, :, -1].max().item() <= tol
    assert want[:, :, :, 8, 4].abs().max().item() > 0.05
    eng.close()
