"""tcgen05 building blocks (descriptors, TMEM operand routes, TMEM load shapes) against numpy, one CTA each."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def bf16_round(x):
    return torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()


@pytest.mark.parametrize("mode", range(16))
@pytest.mark.parametrize("N,K", [(208, 208), (80, 64), (48, 128), (16, 16)])
def test_tc_probe(mode, N, K):
    if (mode & 3) == 3:
        pytest.skip("unused A-source code")
    import os
    import eco_dqn_b200
    import eco_dqn_b200._lib as _lib
    eco_dqn_b200.lib()                  # the probe library links against the product library (test-only kernel, not shipped in it)
    L = C.CDLL(os.path.join(os.path.dirname(_lib.LIB_PATH), "libecodqn_b200_probe.so"))
    fn = L.eco_tc_probe
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    rng = np.random.default_rng(mode * 1000 + N + K)
    A = bf16_round(rng.standard_normal((128, K)).astype(np.float32))
    B = bf16_round(rng.standard_normal((N, K)).astype(np.float32))
    Ad, Bd = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    Dd = torch.full((128, N), float("nan"), device="cuda")
    _lib.check(fn(Ad.data_ptr(), Bd.data_ptr(), Dd.data_ptr(), N, K, mode, None))
    torch.cuda.synchronize()
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    got = Dd.cpu().numpy()
    assert np.allclose(got, ref, rtol=1e-4, atol=1e-3), "mode %d N %d K %d max err %g" % (
        mode, N, K, float(np.nanmax(np.abs(got - ref))))
