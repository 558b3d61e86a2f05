"""GPU parity of every kernel behind the C ABI against plain PyTorch fp32 / numpy / scipy (pytest -m gpu)."""
import pytest

import gpu_kernel_checks as K


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(K.CHECKS))
def test_kernel(name):
    import torch
    res = K.CHECKS[name]()
    torch.cuda.synchronize()
    assert res is not None
