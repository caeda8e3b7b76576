"""Parameter update against golden vectors of the UNMODIFIED reference: oracle/_ref/nts_ref_driver
`adam` mode drives core/NtsScheduler.hpp's `Parameter` exactly as toolkits/GCN.hpp:209-215 does
(all_reduce_to_gradient -> learn..._Adam -> next) for 8 steps and dumps W, M, V after every step
(tests/golden/adam/adam_ref.npz, oracle/make_golden.py --adam).  Pins: the update arithmetic, the bias-correction
schedule folded into alpha by next(), and the reference's `int decay_rate` quirk (0.97 truncates to 0: the weights
freeze after DECAY_EPOCH).  The host mirror (torch ops) runs everywhere; the fused kernel nts_adam_update on the GPU."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
Z = np.load(os.path.join(ROOT, "tests", "golden", "adam", "adam_ref.npz"))


def _run(device):
    from neutronstarlite_b200.toolkits import Parameter
    w, h, steps = (int(x) for x in Z["meta"])
    alpha, b1, b2, eps, wd, decay_rate, decay_epoch = (float(x) for x in Z["hyper"])
    p = Parameter(w, h, alpha, b1, b2, eps, wd, device=device)
    p.W = torch.from_numpy(Z["W0"].copy()).to(device).requires_grad_(True)
    p.set_decay(decay_rate, decay_epoch)
    out = []
    for s in range(steps):
        p.all_reduce_to_gradient(torch.from_numpy(Z["grads"][s].copy()).to(device))
        p.learn_with_decay_Adam()
        p.next()
        out.append((p.W.detach().cpu().numpy().copy(), p.M.cpu().numpy().copy(), p.V.cpu().numpy().copy()))
    return out


def _check(out):
    for s, (W, M, V) in enumerate(out):
        np.testing.assert_allclose(M, Z["M"][s], rtol=2e-6, atol=1e-12, err_msg="M step %d" % s)
        np.testing.assert_allclose(V, Z["V"][s], rtol=2e-6, atol=1e-15, err_msg="V step %d" % s)
        np.testing.assert_allclose(W, Z["W"][s], rtol=1e-5, atol=1e-7, err_msg="W step %d" % s)
    # the quirk: decay_rate is an int in the reference, alpha collapses to 0 at the first decay epoch
    assert np.array_equal(Z["W"][5], Z["W"][6]) and np.array_equal(out[5][0], out[6][0])
    assert not np.array_equal(out[0][0], out[1][0])


def test_parameter_host_mirror_matches_reference():
    _check(_run(torch.device("cpu")))


@pytest.mark.gpu
def test_fused_adam_kernel_matches_reference():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from neutronstarlite_b200 import _lib
    n0 = _lib.load().nts_kernel_launch_count()
    out = _run(torch.device("cuda:0"))
    assert _lib.load().nts_kernel_launch_count() - n0 == int(Z["meta"][2])   # one kernel per step
    _check(out)
