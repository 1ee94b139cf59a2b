"""GPU tests of the TF32 tensor-core form of the fp32 path (csrc/conv_tf32.cu, engine precision "tf32").

BASELINE.json's north_star names an "fp32/TF32 path"; the reference runs its fp32 model through cuDNN with torch's
default allow_tf32 (altfreezing/demo.py:317-324).  Storage, accumulation, bias, residual and ReLU are fp32; the tensor
core reads 10 mantissa bits of each operand, so a product carries a relative error of up to 2^-10 and a K-term dot
product about |x||w| sqrt(K) 2^-11: the tolerances below are that bound, written out per test.  The exact fp32 engine
(FFMA, <= 1e-3 on logits) stays the strict parity path and is tested in test_gpu_parity.py."""
import numpy as np
import pytest
import torch

import afb200
from afb200 import synthetic
from oracle import i3d_oracle
from tests.test_gpu_parity import CONV_CASES, _conv_ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    afb200.lib()                     # fails loudly if libafb200.so is missing
    return torch.device("cuda", 0)


@pytest.fixture(scope="module")
def oracle_out(state_dict):
    clips = np.stack([synthetic.synthetic_clip_u8(i) for i in range(4)])
    x = synthetic.normalise_clip(clips)
    logits, stages = i3d_oracle.forward(state_dict, x, return_stages=True)
    return x, logits, stages


def _tf32_trunc(t):
    """fp32 -> the 10 mantissa bits a TF32 operand keeps (truncation; the hardware may round instead: both are within
    the tolerance used below)."""
    return (t.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("case", CONV_CASES)
def test_tf32_conv_kernel_vs_torch_fp32(dev, case):
    cin, cout, k, s, p, B, T, H, W = case
    g = torch.Generator().manual_seed(cin * 11 + cout)
    x = torch.randn(B, T, H, W, cin, generator=g).to(dev)
    w = torch.randn(cout, cin, *k, generator=g) * (2.0 / (cin * k[0] * k[1] * k[2])) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    K = cin * k[0] * k[1] * k[2]
    for res_on in (False, True):
        y0 = _conv_ref(x, w, b, s, p, False, None)
        res = torch.randn(y0.shape, generator=g).to(dev) if res_on else None
        want = _conv_ref(x, w, b, s, p, True, res)
        got = afb200.conv_ndhwc(x, w, b, s, p, True, res, impl=0, tf32=True).cpu()
        exact = afb200.conv_ndhwc(x, w, b, s, p, True, res, impl=1).cpu()          # FFMA kernel
        assert (exact - want).abs().max().item() <= 1e-4
        # |x| ~ 1, |w| ~ sqrt(2/K): a K-term dot product of TF32-rounded factors is off by ~ sqrt(2) * 2^-10 (rms 3 sigma)
        tol = 6.0 * 2.0 ** -10 * max(1.0, want.abs().max().item())
        assert (got - want).abs().max().item() <= tol, ((got - want).abs().max().item(), tol, K)
        rel = ((got - want).norm() / want.norm()).item()
        assert rel <= 1.5e-3, rel
        # against fp32 maths on TF32-truncated operands the kernel is exact up to accumulation order / rounding mode
        want_t = _conv_ref(_tf32_trunc(x), _tf32_trunc(w), b, s, p, True, res)
        assert ((got - want_t).norm() / want_t.norm()).item() <= 1e-3


def test_tf32_engine_matches_oracle(dev, state_dict, oracle_out):
    """Whole network in precision "tf32": all 53 convs on the tensor cores (the stem through its own form of the kernel,
    straight from the padded fp32 clip), fp32 everywhere else.  Gates: stage rel-L2 <= 2e-3 (53 TF32 GEMMs deep), logits
    within 5e-3 of the fp32 oracle, same decisions; and the tensor-core kernel really ran (no SIMT conv launch)."""
    x, o_logits, o_stages = oracle_out
    eng = afb200.Engine(state_dict, max_batch=4, precision="tf32")
    eng.set_option("keep_stages", 1)
    eng.set_option("reset_stats", 1)
    eng.set_option("profile_events", 1)
    logits = eng.forward(x.to(dev)).cpu()
    torch.cuda.synchronize()
    eng.set_option("profile_events", 0)
    err = (logits - o_logits).abs().max().item()
    assert err <= 5e-3, err
    assert torch.equal(logits > float(o_logits.median()), o_logits > float(o_logits.median()))
    for si in range(5):
        got = eng.get_stage(si + 1).cpu()
        rel = ((got - o_stages[si]).norm() / o_stages[si].norm()).item()
        assert rel <= 2e-3, (si, rel)
    assert eng.get_stat("conv_simt_launches") == 0
    assert eng.get_stat("conv_umma_launches") == 53           # every conv of the I3D
    # and it is faster than the exact engine by a wide margin (FFMA vs tensor cores)
    eng32 = afb200.Engine(state_dict, max_batch=4, precision="fp32")
    for e in (eng, eng32):
        e.forward(x.to(dev))
    torch.cuda.synchronize()
    t = []
    for e in (eng, eng32):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        e.forward(x.to(dev))
        e1.record()
        torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    assert t[0] < 0.5 * t[1], t
    eng.close()
    eng32.close()
