"""tcgen05 / TMEM plumbing self-test: one 128 x N x K bf16 GEMM through the same descriptor + TMEM-load
helpers the bf16 net kernel uses, against a plain PyTorch fp32 matmul of the same bf16 inputs."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,N", [(16, 16), (64, 32), (32, 64), (288, 32), (400, 48)])
def test_umma_gemm_matches_torch(K, N):
    from chinesecheckersagent_b200.engine import Engine
    eng = Engine(0)
    g = torch.Generator(device="cuda").manual_seed(K * 1000 + N)
    A = torch.randn((128, K), device="cuda", generator=g).to(torch.bfloat16)
    Bt = torch.randn((N, K), device="cuda", generator=g).to(torch.bfloat16)
    D = torch.zeros((128, N), device="cuda", dtype=torch.float32)
    eng.call("ccx_debug_umma_gemm", ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(Bt.data_ptr()), K, N, ctypes.c_void_p(D.data_ptr()))
    torch.cuda.synchronize()
    ref = A.float() @ Bt.float().t()
    assert torch.allclose(D, ref, atol=1e-2, rtol=1e-3), (D - ref).abs().max().item()
    eng.close()


@pytest.mark.parametrize("shift", [0, 1, 5, 7, 8, 13])
@pytest.mark.parametrize("K,N", [(32, 32), (64, 64)])
def test_umma_row_shifted_descriptor(K, N, shift):
    """The 3x3 conv addresses its A operand through descriptors whose start is offset by whole rows (16 B) in a
    row-contiguous layout; the result must equal the GEMM on the shifted row window."""
    from chinesecheckersagent_b200.engine import Engine
    eng = Engine(0)
    rows = 144
    g = torch.Generator(device="cuda").manual_seed(K * 1000 + N + shift)
    A = torch.randn((rows, K), device="cuda", generator=g).to(torch.bfloat16)
    Bt = torch.randn((N, K), device="cuda", generator=g).to(torch.bfloat16)
    D = torch.zeros((128, N), device="cuda", dtype=torch.float32)
    eng.call("ccx_debug_umma_gemm_rows", ctypes.c_void_p(A.data_ptr()), rows, shift, ctypes.c_void_p(Bt.data_ptr()), K, N,
             ctypes.c_void_p(D.data_ptr()))
    torch.cuda.synchronize()
    ref = A[shift:shift + 128].float() @ Bt.float().t()
    assert torch.allclose(D, ref, atol=1e-2, rtol=1e-3), (D - ref).abs().max().item()
    eng.close()


@pytest.mark.parametrize("N", [32, 64])
def test_umma_a_operand_from_tmem(N):
    """1x1 convs read their A operand from TMEM (packed 16-bit pairs, element 2j in the low half of column j)."""
    from chinesecheckersagent_b200.engine import Engine
    eng = Engine(0)
    g = torch.Generator(device="cuda").manual_seed(N)
    A = torch.randn((128, 64), device="cuda", generator=g).to(torch.bfloat16)
    Bt = torch.randn((N, 64), device="cuda", generator=g).to(torch.bfloat16)
    D = torch.zeros((128, N), device="cuda", dtype=torch.float32)
    eng.call("ccx_debug_umma_gemm_ts", ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(Bt.data_ptr()), N, ctypes.c_void_p(D.data_ptr()))
    torch.cuda.synchronize()
    ref = A.float() @ Bt.float().t()
    assert torch.allclose(D, ref, atol=1e-2, rtol=1e-3), (D - ref).abs().max().item()
    eng.close()
