"""GPU parity tests, one hot-path kernel at a time, through the C ABI's single-kernel entry points
(include/netcuda.h, netcuda_op_*).  The CPU oracle is the checker; nothing here is a fallback.

Tolerances (written where they are used):
  * integer paths: bit-exact;
  * fp32-accumulating tensor-core GEMM with fp32 output, operands pre-rounded to the operand type on
    both sides: 1e-4 of max|ref| (accumulation-order noise only);
  * bf16 outputs: 2^-8 relative rounding on top -> 6e-3 of max|ref|.
"""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _bf16_round(torch, a):
    return torch.from_numpy(a).to(torch.bfloat16).float().numpy()


def _tf32_trunc(a):
    return (a.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def _max_rel(a, ref):
    return float(np.abs(a.astype(np.float64) - ref.astype(np.float64)).max() / max(np.abs(ref).max(), 1e-30))


GEMM_SHAPES = [
    # (M, N, K) -- K in elements
    (128, 128, 64),     # one tile, one k-block (bf16)
    (128, 256, 64),     # BN = 256 path
    (256, 384, 256),    # several tiles and k-blocks
    (300, 200, 136),    # ragged in every dimension (K = 136: partial k-block, TMA zero fill)
    (64, 10, 64),       # C1 last layer shape: tiny N
    (64, 128, 784),     # C1 first layer
    (1000, 1000, 192),  # more tiles than a wave of 148 would need at BN=256? no: exercises n tail
    (20000, 768, 768),  # > 148 tiles: persistent loop + both accumulator stages + phase wrap
]


@pytest.mark.parametrize("m,n,k", GEMM_SHAPES)
@pytest.mark.parametrize("out_bf16", [False, True])
@pytest.mark.parametrize("pitch", ["odd", "aligned"])
def test_gemm_bf16_vs_oracle(netcuda, oracle, torch_cuda, m, n, k, out_bf16, pitch):
    """pitch "aligned": rows of the output are multiples of 16 bytes -> smem slab + TMA-store epilogue;
    pitch "odd": direct-store epilogue (what e.g. a 10-wide fp32 output layer gets)."""
    torch = torch_cuda
    rng = np.random.default_rng(m * 7 + n * 3 + k)
    a = _bf16_round(torch, rng.uniform(-1, 1, (m, k)).astype(np.float32))
    w = _bf16_round(torch, rng.uniform(-1, 1, (n, k)).astype(np.float32))
    bias = rng.uniform(-1, 1, n).astype(np.float32)
    ldk = (k + 7) // 8 * 8
    da = torch.zeros((m, ldk), dtype=torch.bfloat16, device="cuda")
    dw = torch.zeros((n, ldk), dtype=torch.bfloat16, device="cuda")
    da[:, :k] = torch.from_numpy(a).cuda().to(torch.bfloat16)
    dw[:, :k] = torch.from_numpy(w).cuda().to(torch.bfloat16)
    db = torch.from_numpy(bias).cuda()
    if pitch == "odd":
        ldc = n + 2 if out_bf16 else n + 3  # the epilogue must not assume alignment of rows
    else:
        ldc = (n + 8 + 7) // 8 * 8  # 16-byte multiple for both types, still wider than N
    out = torch.full((m, ldc), -77.0, dtype=torch.bfloat16 if out_bf16 else torch.float32, device="cuda")
    netcuda.op_gemm(da, dw, db, out, netcuda.PREC_BF16, netcuda.OUT_BF16 if out_bf16 else netcuda.OUT_F32,
                    m=m, n=n, k=k, lda=ldk, ldw=ldk, ldc=ldc)
    torch.cuda.synchronize()
    got = out.float().cpu().numpy()
    want = oracle.linear(a, w, bias)
    assert _max_rel(got[:, :n], want) <= (6e-3 if out_bf16 else 1e-4)
    assert (got[:, n:] == -77.0).all()  # nothing written past N


@pytest.mark.parametrize("epi", ["relu", "gelu", "residual"])
def test_gemm_bf16_epilogues(netcuda, oracle, torch_cuda, epi):
    torch = torch_cuda
    m, n, k = 333, 320, 192
    rng = np.random.default_rng(5)
    a = _bf16_round(torch, rng.standard_normal((m, k)).astype(np.float32))
    w = _bf16_round(torch, (rng.standard_normal((n, k)) * 0.1).astype(np.float32))
    bias = rng.standard_normal(n).astype(np.float32)
    lin = oracle.linear(a, w, bias)
    da, dw, db = (torch.from_numpy(v).cuda() for v in (a, w, bias))
    da, dw = da.to(torch.bfloat16), dw.to(torch.bfloat16)
    if epi == "residual":
        res = rng.standard_normal((m, n)).astype(np.float32)
        out = torch.from_numpy(res).cuda()
        netcuda.op_gemm(da, dw, db, out, netcuda.PREC_BF16, netcuda.OUT_F32, epilogue=netcuda.EPI_RESIDUAL)
        want = res + lin
        tol = 1e-4
    else:
        out = torch.empty((m, n), dtype=torch.bfloat16, device="cuda")
        netcuda.op_gemm(da, dw, db, out, netcuda.PREC_BF16, netcuda.OUT_BF16,
                        epilogue=netcuda.EPI_RELU if epi == "relu" else netcuda.EPI_GELU)
        want = np.maximum(lin, 0) if epi == "relu" else oracle.gelu(lin)
        tol = 6e-3
    torch.cuda.synchronize()
    assert _max_rel(out.float().cpu().numpy(), want) <= tol


@pytest.mark.parametrize("m,n,k", [(128, 128, 32), (64, 128, 784), (300, 200, 100), (4096, 512, 512)])
def test_gemm_tf32_vs_oracle(netcuda, oracle, torch_cuda, m, n, k):
    torch = torch_cuda
    rng = np.random.default_rng(m + n + k)
    # operands pre-truncated to tf32 (10 mantissa bits): products are exact in fp32 on both sides
    a = _tf32_trunc(rng.uniform(-1, 1, (m, k)).astype(np.float32))
    w = _tf32_trunc(rng.uniform(-1, 1, (n, k)).astype(np.float32))
    bias = rng.uniform(-1, 1, n).astype(np.float32)
    da, dw, db = (torch.from_numpy(v).cuda() for v in (a, w, bias))
    out = torch.empty((m, n), dtype=torch.float32, device="cuda")
    netcuda.op_gemm(da, dw, db, out, netcuda.PREC_TF32, netcuda.OUT_F32, epilogue=netcuda.EPI_RELU)
    torch.cuda.synchronize()
    want = np.maximum(oracle.linear(a, w, bias), 0)
    assert _max_rel(out.cpu().numpy(), want) <= 1e-4


# (M, N, K, row pitch of A and W, element offset of both base pointers).  The latency-oriented ordered kernel serves every problem with
# fewer 64 x 64 tiles than SMs: CH = 1 / 2 / 4 neurons per thread by grid size, 16-byte cp.async when pointers and pitches allow it
# (with a zero-filled partial piece when K is not a multiple of four), 4-byte cp.async otherwise.
FP32_ORDERED_SHAPES = [
    (1, 128, 784, 784, 0),     # config C1, first layer, the reference's one sample per call
    (64, 128, 784, 784, 0),    # ... at batch 64
    (64, 10, 64, 64, 0),       # C1 last layer: a partly empty neuron group
    (37, 33, 61, 61, 0),       # nothing aligned: 4-byte copies, ragged k tail
    (5, 7, 3, 3, 0),           # K smaller than one float4
    (64, 128, 784, 784, 1),    # base pointers 4 bytes off a 16-byte boundary: 4-byte copies
    (33, 50, 70, 72, 0),       # K % 4 = 2 under an aligned pitch: partial 16-byte piece
    (200, 40, 300, 300, 0),    # four sample tiles, the last one ragged; five k chunks (ring wraps)
    (64, 1000, 100, 100, 0),   # CH = 2
    (64, 4090, 136, 136, 0),   # CH = 4, ragged neuron tail
    (130, 1000, 1030, 1032, 0),  # CH = 4 with three sample tiles, 17 k chunks
]


@pytest.mark.parametrize("m,n,k,ld,off", FP32_ORDERED_SHAPES)
@pytest.mark.parametrize("relu", [0, 1])
def test_gemm_fp32_ordered_small_problem_kernel_bit_equal(netcuda, oracle, torch_cuda, m, n, k, ld, off, relu):
    """NETCUDA_PREC_FP32 on problems that do not fill the GPU: the latency-oriented kernel (variant 0) against the oracle's
    ordered fmaf chain and against the 64 x 64 tile kernel (variant 1), bit for bit."""
    torch = torch_cuda
    rng = np.random.default_rng(m * 7 + n * 3 + k + off)
    a = rng.uniform(-1, 1, (m, k)).astype(np.float32)
    w = rng.uniform(-1, 1, (n, k)).astype(np.float32)
    bias = rng.uniform(-1, 1, n).astype(np.float32)
    da_buf = torch.full((m * ld + 8,), float("nan"), dtype=torch.float32, device="cuda")
    dw_buf = torch.full((n * ld + 8,), float("nan"), dtype=torch.float32, device="cuda")
    da = da_buf[off:off + m * ld].view(m, ld)
    dw = dw_buf[off:off + n * ld].view(n, ld)
    da[:, :k] = torch.from_numpy(a).cuda()
    dw[:, :k] = torch.from_numpy(w).cuda()
    db = torch.from_numpy(bias).cuda()
    epi = netcuda.EPI_RELU if relu else netcuda.EPI_NONE
    want = oracle.linear(a, w, bias)
    if relu:
        want = np.maximum(want, 0)
    outs = []
    for variant in (0, 1):
        out = torch.full((m, n + 3), -7.0, dtype=torch.float32, device="cuda")  # padded pitch: columns past N stay untouched
        netcuda.op_gemm(da, dw, db, out, netcuda.PREC_FP32, netcuda.OUT_F32, epilogue=epi, variant=variant, m=m, n=n, k=k, lda=ld, ldw=ld)
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        assert (got[:, n:] == -7.0).all()
        outs.append(got[:, :n])
    np.testing.assert_array_equal(outs[0], want)
    np.testing.assert_array_equal(outs[1], want)


@pytest.mark.parametrize("m,n,k", [(128, 128, 128), (1, 4096, 4096), (77, 300, 200), (513, 4096, 4096), (16384, 256, 4096)])
def test_gemm_int8_bit_exact(netcuda, torch_cuda, m, n, k):
    torch = torch_cuda
    rng = np.random.default_rng(m * 3 + n + k)
    a = rng.integers(-128, 128, (m, k), dtype=np.int8)
    w = rng.integers(-128, 128, (n, k), dtype=np.int8)
    bias = rng.integers(-(1 << 14), 1 << 14, n, dtype=np.int32)
    ldk = (k + 15) // 16 * 16
    da = torch.zeros((m, ldk), dtype=torch.int8, device="cuda")
    dw = torch.zeros((n, ldk), dtype=torch.int8, device="cuda")
    da[:, :k] = torch.from_numpy(a).cuda()
    dw[:, :k] = torch.from_numpy(w).cuda()
    db = torch.from_numpy(bias).cuda()
    # exact integer reference on the GPU in int64-free form: fp64 matmul is exact for |acc| < 2^53
    acc = (da[:, :k].double() @ dw[:, :k].double().T).cpu().numpy().astype(np.int64) + bias.astype(np.int64)
    out32 = torch.empty((m, n), dtype=torch.int32, device="cuda")
    netcuda.op_gemm(da, dw, db, out32, netcuda.PREC_INT8, netcuda.OUT_S32, m=m, n=n, k=k, lda=ldk, ldw=ldk)
    out8 = torch.empty((m, n), dtype=torch.int8, device="cuda")
    netcuda.op_gemm(da, dw, db, out8, netcuda.PREC_INT8, netcuda.OUT_S8, epilogue=netcuda.EPI_RELU, m=m, n=n, k=k, lda=ldk, ldw=ldk)
    out8s = torch.empty((m, n), dtype=torch.int8, device="cuda")
    netcuda.op_gemm(da, dw, db, out8s, netcuda.PREC_INT8, netcuda.OUT_S8, m=m, n=n, k=k, lda=ldk, ldw=ldk)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out32.cpu().numpy().astype(np.int64), acc)
    np.testing.assert_array_equal(out8.cpu().numpy().astype(np.int64), np.minimum(127, np.maximum(acc, 0) >> 7))
    np.testing.assert_array_equal(out8s.cpu().numpy().astype(np.int64), np.clip(acc >> 7, -128, 127))


def test_gemm_tensor_core_matches_cuda_core_variant_at_full_size(netcuda, torch_cuda):
    """ViT-B fc1 at a full pass (M = 64 images * 197 tokens): tcgen05 kernel vs the CUDA-core kernel with
    the same operand rounding -- a size the CPU oracle would need minutes for."""
    torch = torch_cuda
    m, n, k = 64 * 197, 3072, 768
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randn((m, k), generator=g, device="cuda").to(torch.bfloat16)
    w = (torch.randn((n, k), generator=g, device="cuda") * 0.05).to(torch.bfloat16)
    b = torch.randn(n, generator=g, device="cuda")
    o0 = torch.empty((m, n), dtype=torch.float32, device="cuda")
    o1 = torch.empty_like(o0)
    netcuda.op_gemm(a, w, b, o0, netcuda.PREC_BF16, netcuda.OUT_F32, variant=0)
    netcuda.op_gemm(a, w, b, o1, netcuda.PREC_BF16, netcuda.OUT_F32, variant=1)
    torch.cuda.synchronize()
    err = ((o0 - o1).abs().max() / o1.abs().max()).item()
    assert err <= 1e-4, err
    # CTA-pair kernel (default for M > 128) vs one CTA per tile: same K order, same epilogue -> identical bits
    o2 = torch.empty_like(o0)
    netcuda.op_gemm(a, w, b, o2, netcuda.PREC_BF16, netcuda.OUT_F32, variant=2)
    torch.cuda.synchronize()
    assert torch.equal(o0, o2)
    # GELU epilogue: 8 epilogue warps (default) vs 16 (variant 3) vs one CTA per tile (variant 2): identical bits
    g16 = torch.empty((m, n), dtype=torch.bfloat16, device="cuda")
    g8, g1 = torch.empty_like(g16), torch.empty_like(g16)
    netcuda.op_gemm(a, w, b, g16, netcuda.PREC_BF16, netcuda.OUT_BF16, epilogue=netcuda.EPI_GELU, variant=0)
    netcuda.op_gemm(a, w, b, g8, netcuda.PREC_BF16, netcuda.OUT_BF16, epilogue=netcuda.EPI_GELU, variant=3)
    netcuda.op_gemm(a, w, b, g1, netcuda.PREC_BF16, netcuda.OUT_BF16, epilogue=netcuda.EPI_GELU, variant=2)
    torch.cuda.synchronize()
    assert torch.equal(g16, g8) and torch.equal(g16, g1)
    want = torch.nn.functional.gelu(o1)  # exact-erf GELU of the CUDA-core result
    assert ((g16.float() - want).abs().max() / want.abs().max()).item() <= 6e-3


@pytest.mark.parametrize("n,k,out_bf16,epi", [(576, 192, True, "none"), (192, 192, False, "residual"), (768, 192, True, "gelu"), (192, 768, False, "residual"),
                                              (200, 136, False, "none")])
def test_gemm_short_k_sixteen_epilogue_warps(netcuda, torch_cuda, n, k, out_bf16, epi):
    """GEMMs that are all epilogue (K <= 256 or a single column of tiles: the linear layers of ViT-Tiny) run with 16 epilogue warps.
    Against the CUDA-core kernel with the same operand rounding (1e-4 / bf16 rounding), and bit for bit against the 8-warp
    configurations (variant 5: one slab per warp, variant 4: two)."""
    torch = torch_cuda
    m = 64 * 197  # 50 row blocks of 256: the CTA-pair path
    g = torch.Generator(device="cuda").manual_seed(n + k)
    a = torch.randn((m, k), generator=g, device="cuda").to(torch.bfloat16)
    w = (torch.randn((n, k), generator=g, device="cuda") * 0.05).to(torch.bfloat16)
    b = torch.randn(n, generator=g, device="cuda")
    base = torch.randn((m, n), generator=g, device="cuda")
    e = {"none": netcuda.EPI_NONE, "residual": netcuda.EPI_RESIDUAL, "gelu": netcuda.EPI_GELU}[epi]
    outs = {}
    for variant in (0, 3, 4, 5, 1):
        out = base.clone() if not out_bf16 else torch.empty((m, n), dtype=torch.bfloat16, device="cuda")
        netcuda.op_gemm(a, w, b, out, netcuda.PREC_BF16, netcuda.OUT_BF16 if out_bf16 else netcuda.OUT_F32, epilogue=e, variant=variant)
        torch.cuda.synchronize()
        outs[variant] = out.float()
    for variant in (3, 4, 5):
        assert torch.equal(outs[0], outs[variant]), variant
    err = ((outs[0] - outs[1]).abs().max() / outs[1].abs().max()).item()
    assert err <= (6e-3 if out_bf16 else 1e-4), err


@pytest.mark.parametrize("n,k,epi", [(576, 192, "none"), (192, 192, "residual"), (768, 192, "gelu"), (192, 256, "residual")])
def test_gemm_tf32_short_k_sixteen_epilogue_warps(netcuda, torch_cuda, n, k, epi):
    """The same for tf32 operands with fp32 outputs (the linear layers of a TF32 ViT-Tiny): 16 epilogue warps (default at K <= 256)
    against 8 (variant 5) bit for bit, and against the CUDA-core kernel with tf32-truncated operands."""
    torch = torch_cuda
    m = 64 * 197
    g = torch.Generator(device="cuda").manual_seed(n * 3 + k)
    a = torch.randn((m, k), generator=g, device="cuda")
    w = torch.randn((n, k), generator=g, device="cuda") * 0.05
    b = torch.randn(n, generator=g, device="cuda")
    base = torch.randn((m, n), generator=g, device="cuda")
    e = {"none": netcuda.EPI_NONE, "residual": netcuda.EPI_RESIDUAL, "gelu": netcuda.EPI_GELU}[epi]
    outs = {}
    for variant in (0, 3, 5, 1):
        out = base.clone()
        netcuda.op_gemm(a, w, b, out, netcuda.PREC_TF32, netcuda.OUT_F32, epilogue=e, variant=variant)
        torch.cuda.synchronize()
        outs[variant] = out
    assert torch.equal(outs[0], outs[3]) and torch.equal(outs[0], outs[5])
    err = ((outs[0] - outs[1]).abs().max() / outs[1].abs().max()).item()
    assert err <= 1e-4, err


@pytest.mark.parametrize("rows,dim", [(1, 192), (197, 192), (1001, 64), (50, 128), (1000, 768), (333, 1024), (64, 4096)])
def test_layernorm_vs_oracle(netcuda, oracle, torch_cuda, rows, dim):
    torch = torch_cuda
    rng = np.random.default_rng(rows + dim)
    x = (rng.standard_normal((rows, dim)) * 2 + 0.5).astype(np.float32)
    g = rng.standard_normal(dim).astype(np.float32)
    b = rng.standard_normal(dim).astype(np.float32)
    dx, dg, db = (torch.from_numpy(v).cuda() for v in (x, g, b))
    y = torch.empty((rows, dim), dtype=torch.bfloat16, device="cuda")
    netcuda.op_layernorm(dx, dg, db, y)
    torch.cuda.synchronize()
    want = oracle.layernorm(x, g, b)
    # bf16 output rounding (2^-9 relative) dominates
    assert _max_rel(y.float().cpu().numpy(), want) <= 5e-3


def test_layernorm_strided_rows(netcuda, oracle, torch_cuda):
    """Final LayerNorm reads only the class-token rows: row pitch = tokens * dim."""
    torch = torch_cuda
    rng = np.random.default_rng(9)
    B, T, D = 5, 7, 256
    x = rng.standard_normal((B * T, D)).astype(np.float32)
    g, b = rng.standard_normal(D).astype(np.float32), rng.standard_normal(D).astype(np.float32)
    dx, dg, db = (torch.from_numpy(v).cuda() for v in (x, g, b))
    y = torch.empty((B, D), dtype=torch.bfloat16, device="cuda")
    netcuda.op_layernorm(dx, dg, db, y, rows=B, dim=D, ldx=T * D, ldy=D)
    torch.cuda.synchronize()
    want = oracle.layernorm(x[::T].copy(), g, b)
    assert _max_rel(y.float().cpu().numpy(), want) <= 5e-3


ATTENTION_SHAPES = [(2, 197, 3), (1, 5, 1), (3, 37, 2), (1, 577, 2), (2, 64, 12), (4, 16, 1), (1, 113, 1),
                    (1, 128, 1), (2, 129, 2), (2, 256, 1), (3, 200, 2), (1, 257, 1), (40, 197, 12),
                    (3, 577, 4), (2, 300, 1), (1, 1025, 2), (2, 640, 3), (5, 513, 1)]  # > 256 tokens: key-blocked kernel


@pytest.mark.parametrize("batch,tokens,heads", ATTENTION_SHAPES)
def test_attention_vs_oracle(netcuda, oracle, torch_cuda, batch, tokens, heads):
    torch = torch_cuda
    rng = np.random.default_rng(tokens * 5 + heads)
    qkv = _bf16_round(torch, rng.standard_normal((batch * tokens, 3 * heads * 64)).astype(np.float32))
    dq = torch.from_numpy(qkv).cuda().to(torch.bfloat16)
    out = torch.full((batch * tokens, heads * 64), 55.0, dtype=torch.bfloat16, device="cuda")
    netcuda.op_attention(dq, out, batch, tokens, heads)
    torch.cuda.synchronize()
    want = oracle.attention(qkv, batch, tokens, heads)
    got = out.float().cpu().numpy()
    # P is rounded to bf16 before P.V and the output is bf16: 1e-2 of max|ref| (north_star tolerance)
    assert _max_rel(got, want) <= 1e-2


@pytest.mark.parametrize("tokens", [197, 577])
def test_attention_late_peaks(netcuda, oracle, torch_cuda, tokens):
    """Rows whose maximum arrives late -- keys 100 and 180 beat everything before them by 8.7 and 9 binades (and, in the key-blocked
    kernel, key 400 of a later block does so again) while the first 32 keys still carry ~8 % of the weight before the second peak:
    any shortcut around the exact row maximum, or a wrong rescale of the running output across key blocks, shows up far above the
    tolerance."""
    torch = torch_cuda
    rng = np.random.default_rng(99)
    batch, heads = 3, 2
    qkv = (rng.standard_normal((batch, tokens, 3, heads, 64)) * 0.25).astype(np.float32)
    u = np.full(64, 0.125, np.float32)  # unit vector; all products below are exact in bf16
    qkv[:, :, 0, 0, :] = 8.0 * u                       # every query of head 0
    qkv[:, :32, 1, 0, :] = 3.0 * u                     # raw score 24 for the first chunk of keys
    qkv[:, 100, 1, 0, :] = 9.0 * u                     # 72: +48 raw = +8.7 binades -> first raise
    qkv[0, 180, 1, 0, :] = 15.25 * u                   # image 0 only: 122 = +9 binades -> second raise
    qkv[:, :32, 2, 0, :] = 1.0                         # values that tell the three groups apart
    qkv[:, 100, 2, 0, :] = -1.0
    qkv[:, 180, 2, 0, :] = 3.0
    if tokens > 256:  # key-blocked kernel (online softmax across key blocks): image 2 gets a new maximum in a LATER key block
        qkv[2, 400, 1, 0, :] = 15.25 * u
        qkv[2, 400, 2, 0, :] = 5.0
    qkv = _bf16_round(torch, qkv.reshape(batch * tokens, 3 * heads * 64))
    dq = torch.from_numpy(qkv).cuda().to(torch.bfloat16)
    out = torch.full((batch * tokens, heads * 64), 55.0, dtype=torch.bfloat16, device="cuda")
    netcuda.op_attention(dq, out, batch, tokens, heads)
    torch.cuda.synchronize()
    want = oracle.attention(qkv, batch, tokens, heads)
    got = out.float().cpu().numpy()
    assert -0.9 < want[tokens + 5, 0] < -0.75  # image 1: mostly key 100 (-1), visibly pulled towards the first chunk's +1
    assert want[5, 0] > 2.9                    # image 0: key 180 (3) after the second raise
    assert _max_rel(got, want) <= 1e-2


def test_attention_tcgen05_matches_mma_sync_kernel(netcuda, torch_cuda, monkeypatch):
    """Both attention kernels on a full ViT-B pass worth of heads (256 images x 12 heads x 197 tokens)."""
    torch = torch_cuda
    batch, tokens, heads = 256, 197, 12
    g = torch.Generator(device="cuda").manual_seed(11)
    qkv = torch.randn((batch * tokens, 3 * heads * 64), generator=g, device="cuda").to(torch.bfloat16)
    o0 = torch.empty((batch * tokens, heads * 64), dtype=torch.bfloat16, device="cuda")
    o1 = torch.empty_like(o0)
    netcuda.op_attention(qkv, o0, batch, tokens, heads)
    monkeypatch.setenv("NETCUDA_ATTENTION_VARIANT", "1")
    netcuda.op_attention(qkv, o1, batch, tokens, heads)
    torch.cuda.synchronize()
    err = ((o0.float() - o1.float()).abs().max() / o1.float().abs().max()).item()
    assert err <= 1e-2, err


@pytest.mark.parametrize("kernel", [0, 1, 2, 4, 14, 24, 34, 104])
@pytest.mark.parametrize("batch,tokens,heads", [(40, 197, 12), (3, 37, 2), (2, 129, 2), (2, 256, 1), (1, 128, 1), (5, 200, 3)])
def test_attention_kernel_variants_vs_oracle(netcuda, oracle, torch_cuda, kernel, batch, tokens, heads):
    """Every build variant of the short-sequence tcgen05 kernels (12-warp kernel: polling / per-tile MMA issuers, exp2 turn-taking;
    16-softmax-warp kernel: key halves per row block, row sums from the N = 80 P.V MMA, single / double-buffered loads, both key
    splits, with and without tile turn-taking), bf16 and fp32 outputs, against the oracle; several items per CTA (40 x 12 heads over
    148 SMs), one and two query tiles, ragged last tile, sequences with and without keys for the second half."""
    torch = torch_cuda
    rng = np.random.default_rng(tokens * 7 + heads + kernel)
    qkv = _bf16_round(torch, (rng.standard_normal((batch * tokens, 3 * heads * 64)) * 1.5).astype(np.float32))
    dq = torch.from_numpy(qkv).cuda().to(torch.bfloat16)
    want = oracle.attention(qkv, batch, tokens, heads)
    for f32 in (False, True):
        out = torch.full((batch * tokens, heads * 64), 55.0, dtype=torch.float32 if f32 else torch.bfloat16, device="cuda")
        netcuda.op_attention_ex(dq, out, batch, tokens, heads, kernel=kernel, out_f32=f32)
        torch.cuda.synchronize()
        assert _max_rel(out.float().cpu().numpy(), want) <= 1e-2, (kernel, f32)


@pytest.mark.parametrize("batch,tokens,heads", [(2, 577, 3), (1, 300, 1), (2, 197, 2)])
def test_attention_fp32_output(netcuda, oracle, torch_cuda, batch, tokens, heads):
    """fp32 output rows (two 32-column TMA store boxes per warp) of the key-blocked kernel, the short-sequence kernel and the
    mma.sync cross-check kernel: the bf16 output of the same kernel, before its final rounding."""
    torch = torch_cuda
    rng = np.random.default_rng(tokens)
    qkv = torch.from_numpy(rng.standard_normal((batch * tokens, 3 * heads * 64)).astype(np.float32)).cuda().to(torch.bfloat16)
    for kernel in (-1, netcuda.ATT_KERNEL_MMA_SYNC):
        o16 = torch.full((batch * tokens, heads * 64), 7.0, dtype=torch.bfloat16, device="cuda")
        o32 = torch.full((batch * tokens, heads * 64), 7.0, dtype=torch.float32, device="cuda")
        netcuda.op_attention_ex(qkv, o16, batch, tokens, heads, kernel=kernel, out_f32=False)
        netcuda.op_attention_ex(qkv, o32, batch, tokens, heads, kernel=kernel, out_f32=True)
        torch.cuda.synchronize()
        assert torch.equal(o32.to(torch.bfloat16), o16), kernel
    want = oracle.attention(qkv.float().cpu().numpy(), batch, tokens, heads)
    assert _max_rel(o32.cpu().numpy(), want) <= 1e-2


def test_patchify_exact(netcuda, torch_cuda):
    torch = torch_cuda
    rng = np.random.default_rng(4)
    B, S, P = 3, 64, 16
    img = rng.uniform(-1, 1, (B, 3, S, S)).astype(np.float32)
    g = S // P
    dimg = torch.from_numpy(img).cuda()
    out = torch.empty((B * g * g, 3 * P * P), dtype=torch.bfloat16, device="cuda")
    netcuda.op_patchify(dimg, out, B, S, P)
    torch.cuda.synchronize()
    # column = c*P*P + py*P + px: the flattening order of a conv weight [D][3][P][P] (torchvision conv_proj)
    want = img.reshape(B, 3, g, P, g, P).transpose(0, 2, 4, 1, 3, 5).reshape(B * g * g, 3 * P * P)
    want = torch.from_numpy(want).to(torch.bfloat16)
    assert torch.equal(out.cpu(), want)
