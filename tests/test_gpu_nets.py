"""GPU parity tests of whole nets: the CUDA path driven through the C ABI (netcuda.Net) and through the
C++ class behind net::net_abstract* (netcuda.HostNet), checked against the CPU oracle and the committed
golden fixtures.

Bars, error = max |out - ref| relative to max |ref| per sample, worst sample:
  * NETCUDA_PREC_FP32: bit-equal to the oracle (same fmaf sequence);
  * TF32 (default for nets built from net::net_data) and BF16 ViT at the BASELINE shapes (ViT-Ti/B/L):
    <= 1e-2 and identical arg-max -- the north_star tolerance;
  * BF16 on narrow nets (the 128-wide golden ViT, MLPs with the reference's unscaled +-1 weights): the
    operand rounding alone (2^-9 per weight and per activation) costs more than 1e-2 on the worst sample --
    1.14e-2 on the golden ViT, reproduced to 7 digits by a CPU model of the rounding points
    (tests/bf16_pipeline_model.py).  There the bar is split: <= 2e-3 against that rounding model (what the
    kernels control) and <= 2e-2 against the fp32 reference (what bf16 operands cost);
  * INT8: bit-exact at every batch size.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, c1_net, rel_err

pytestmark = pytest.mark.gpu


# ---- config C1: 784-128-64-10, batch 64 -----------------------------------------------------------

def test_c1_golden_fp32_bit_equal(netcuda, oracle, torch_cuda):
    """The golden fixture holds outputs of the reference's own host runtime (oracle/_ref)."""
    g = np.load(os.path.join(GOLDEN, "mlp_c1.npz"))
    npl, n_ins, w, b = c1_net(oracle)
    net = netcuda.Net.mlp(npl, n_ins, precision=netcuda.PREC_FP32)
    net.upload_mlp(w, b)
    got = net.forward(g["x"])
    np.testing.assert_array_equal(got, g["y"])
    assert net.last_forward_us > 0
    net.close()


@pytest.mark.parametrize("prec", ["tf32", "bf16"])
def test_c1_tensor_core_precisions(netcuda, oracle, torch_cuda, prec):
    g = np.load(os.path.join(GOLDEN, "mlp_c1.npz"))
    npl, n_ins, w, b = c1_net(oracle)
    net = netcuda.Net.mlp(npl, n_ins, precision=netcuda.PRECISIONS[prec])
    net.upload_mlp(w, b)
    got = net.forward(g["x"])
    net.close()
    # tf32: north_star tolerance.  bf16: +-1 weights over fan-in 784 leave 1.4e-2 of operand-rounding noise.
    assert rel_err(got, g["y"]) <= (1e-2 if prec == "tf32" else 2e-2)
    np.testing.assert_array_equal(got.argmax(1), g["y"].argmax(1))


@pytest.mark.parametrize("act", [0, 1, 2])
def test_mlp_activation_modes_and_ragged_shapes(netcuda, oracle, torch_cuda, act):
    rng = np.random.default_rng(21 + act)
    npl, n_ins = [33, 17, 9, 4], 61  # nothing is a multiple of a tile or of 16 bytes
    n_params = sum(a * b for a, b in zip([n_ins] + npl[:-1], npl))
    w = rng.uniform(-1, 1, n_params).astype(np.float32)
    b = rng.uniform(-1, 1, sum(npl)).astype(np.float32)
    x = rng.uniform(-1, 1, (37, n_ins)).astype(np.float32)
    want = oracle.mlp_forward(x, w, b, npl, n_ins, act)
    # act 2 (purely linear) lets hidden negatives cancel in the output, which magnifies relative rounding noise
    for prec, tol in (("fp32", 0.0), ("tf32", 1e-2), ("bf16", 5e-2 if act == 2 else 2e-2)):
        net = netcuda.Net.mlp(npl, n_ins, precision=netcuda.PRECISIONS[prec], activation=act)
        net.upload_mlp(w, b)
        got = net.forward(x)
        net.close()
        if tol == 0.0:
            np.testing.assert_array_equal(got, want)
        else:
            assert rel_err(got, want) <= tol, prec


def test_mlp_edge_batches(netcuda, oracle, torch_cuda):
    """Empty batch, batch 1 (the reference's contract), and a batch that needs several internal passes."""
    npl, n_ins, w, b = c1_net(oracle)
    net = netcuda.Net.mlp(npl, n_ins, precision=netcuda.PREC_FP32, max_batch=50)
    net.upload_mlp(w, b)
    assert net.forward(np.zeros((0, n_ins), np.float32)).shape == (0, 10)
    rng = np.random.default_rng(8)
    for batch in (1, 49, 50, 51, 173):
        x = rng.uniform(-1, 1, (batch, n_ins)).astype(np.float32)
        np.testing.assert_array_equal(net.forward(x), oracle.mlp_forward(x, w, b, npl, n_ins))
    net.close()
    with pytest.raises(netcuda.NetcudaError):
        fresh = netcuda.Net.mlp(npl, n_ins)
        fresh.forward(np.zeros((1, n_ins), np.float32))  # forward before upload must fail loudly


def test_forward_device_and_pinned_paths(netcuda, oracle, torch_cuda):
    torch = torch_cuda
    npl, n_ins, w, b = c1_net(oracle)
    net = netcuda.Net.mlp(npl, n_ins, precision=netcuda.PREC_FP32)
    net.upload_mlp(w, b)
    x = np.random.default_rng(3).uniform(-1, 1, (64, n_ins)).astype(np.float32)
    want = oracle.mlp_forward(x, w, b, npl, n_ins)
    dx = torch.from_numpy(x).cuda()
    dy = torch.empty((64, 10), dtype=torch.float32, device="cuda")
    net.forward_device(dx, dy, 64, torch.cuda.current_stream())
    torch.cuda.synchronize()
    np.testing.assert_array_equal(dy.cpu().numpy(), want)
    px = torch.from_numpy(x).pin_memory()
    py = torch.empty((64, 10), dtype=torch.float32).pin_memory()
    net.forward_into(px, py)  # page-locked buffers are DMA'd in place
    np.testing.assert_array_equal(py.numpy(), want)
    net.close()


# ---- the C++ class through net::net_abstract* -------------------------------------------------------

def test_cpp_class_drop_in(netcuda, oracle, torch_cuda):
    npl, n_ins, w, b = c1_net(oracle)
    g = np.load(os.path.join(GOLDEN, "mlp_c1.npz"))
    net = netcuda.HostNet.mlp(npl, n_ins, w, b, precision=netcuda.PREC_FP32)
    # one sample per call: exactly the reference's contract (src/netFPGA.cpp:266-289)
    for i in (0, 17, 63):
        np.testing.assert_array_equal(net.launch_forward(g["x"][i])[0], g["y"][i])
    # batched extension: B * n_ins in, B * n_out out
    np.testing.assert_array_equal(net.launch_forward(g["x"]), g["y"])
    assert net.forward_us() > 0
    with pytest.raises(ValueError):
        net.launch_forward(np.zeros(n_ins + 1, np.float32))  # the reference would read out of bounds here
    # get_net_data is the exact inverse of the constructor's flatten (the reference's is broken, App. A)
    w2, b2, n_ins2, n_layers2 = net.get_net_data(w.size, b.size)
    np.testing.assert_array_equal(w2, w)
    np.testing.assert_array_equal(b2, b)
    assert (n_ins2, n_layers2) == (n_ins, 3)
    assert net.check_stubs() == 0
    net.close()


def test_cpp_class_random_init_matches_reference_rule(netcuda, oracle, torch_cuda):
    """random=true draws float(rand()%200-100)/100, params first then biases (src/netFPGA.cpp:82-88)."""
    npl, n_ins = [128, 64, 10], 784
    net = netcuda.HostNet.mlp(npl, n_ins, random=True, seed=1, precision=netcuda.PREC_FP32)
    w, b = oracle.rand_init(1, 109184, 202)
    w2, b2, _, _ = net.get_net_data(w.size, b.size)
    np.testing.assert_array_equal(w2, w)
    np.testing.assert_array_equal(b2, b)
    net.close()


def test_example_consumer_application(netcuda, oracle, torch_cuda, tmp_path):
    """examples/drop_in_app.cpp -- a C++ host application in the shape of the reference's consumer (backend header, backend class,
    then only net::net_abstract*): built with g++ -std=gnu++14, run with NETCUDA_PRECISION=fp32, its printed outputs are the CPU
    oracle's bits for the reference's own random initialisation (srand(1); src/netFPGA.cpp:82-88)."""
    import subprocess
    from test_boundary import build_example_app
    exe = build_example_app(str(tmp_path))
    r = subprocess.run([exe], capture_output=True, text=True, env=dict(os.environ, NETCUDA_PRECISION="fp32"), timeout=120)
    assert r.returncode == 0, r.stderr
    line = next(l for l in r.stdout.splitlines() if l.startswith("outputs"))
    got = np.array([float(v) for v in line.split()[1:]], dtype=np.float32)
    npl, n_ins = [128, 64, 10], 784
    w, b = oracle.rand_init(1, 784 * 128 + 128 * 64 + 64 * 10, sum(npl))
    i = np.arange(n_ins, dtype=np.uint64)
    x = (((i * 2654435761) & 0xFFFFFFFF) >> 8 & 0xFFFF).astype(np.float32) / np.float32(32768.0) - np.float32(1.0)
    want = oracle.mlp_forward(x[None], w, b, npl, n_ins)[0]
    np.testing.assert_array_equal(got, want)
    assert "one sample per call" in r.stdout and "64 samples per call" in r.stdout


def test_cpp_class_shards_over_gpus(netcuda, oracle, torch_cuda, monkeypatch):
    """net_cuda_options::n_devices / NETCUDA_DEVICES: one net_cuda object drives several GPUs -- weights replicated, every batched
    forward cut into contiguous slices, one host thread per GPU.  Samples are independent, so the outputs are the single-GPU bits
    (MLP fp32 path: the oracle's bits; ViT: the same kernels on another device)."""
    if torch_cuda.cuda.device_count() < 2:
        if int(os.environ.get("NETCUDA_REQUIRE_GPUS", "0")) >= 2:
            pytest.fail("NETCUDA_REQUIRE_GPUS >= 2 but fewer than two CUDA devices are visible")
        pytest.skip("needs two GPUs (NETCUDA_REQUIRE_GPUS=2 turns this skip into a failure)")
    npl, n_ins, w, b = c1_net(oracle)
    x = np.random.default_rng(2).uniform(-1, 1, (301, n_ins)).astype(np.float32)  # ragged slices: 151 + 150
    monkeypatch.setenv("NETCUDA_DEVICES", "2")
    net2 = netcuda.HostNet.mlp(npl, n_ins, w, b, precision=netcuda.PREC_FP32)
    np.testing.assert_array_equal(net2.launch_forward(x), oracle.mlp_forward(x, w, b, npl, n_ins))
    np.testing.assert_array_equal(net2.launch_forward(x[:1]), oracle.mlp_forward(x[:1], w, b, npl, n_ins))  # one sample: one GPU
    assert net2.forward_us() > 0
    net2.close()
    g = np.load(os.path.join(GOLDEN, "vit_small.npz"))
    cfg = dict(zip(("image_size", "patch_size", "dim", "depth", "heads", "mlp_dim", "n_classes"), (int(v) for v in g["cfg"])))
    imgs = np.random.default_rng(3).uniform(-1, 1, (37, 3 * cfg["image_size"] ** 2)).astype(np.float32)
    vit2 = netcuda.HostNet.vit(cfg, g["flat"], max_batch=8)
    got2 = vit2.launch_forward(imgs)
    vit2.close()
    monkeypatch.setenv("NETCUDA_DEVICES", "1")
    vit1 = netcuda.HostNet.vit(cfg, g["flat"], max_batch=8)
    np.testing.assert_array_equal(got2, vit1.launch_forward(imgs))
    vit1.close()


def test_page_locked_caller_buffers(netcuda, torch_cuda, monkeypatch):
    """netcuda_host_register: a caller-owned buffer that is page-locked is DMA'd from in place (no staging copy) -- same bits as the
    staged path, registering twice is not an error, the buffer can be unregistered and used again.  And the class does it for its
    callers under net_cuda_options::pin_inputs / NETCUDA_PIN_INPUTS: the vector a launch_forward call reads is page-locked the first
    time it is seen, later calls reuse it, a vector that dies is forgotten first (release_inputs)."""
    g = np.load(os.path.join(GOLDEN, "vit_small.npz"))
    cfg = dict(zip(("image_size", "patch_size", "dim", "depth", "heads", "mlp_dim", "n_classes"), (int(v) for v in g["cfg"])))
    n_in = 3 * cfg["image_size"] ** 2
    batch = max(37, (3 << 20) // (4 * n_in) + 1)  # > 1 MiB of input: below that the class does not bother
    imgs = np.random.default_rng(4).uniform(-1, 1, (batch, n_in)).astype(np.float32)
    net = netcuda.Net.vit(cfg, max_batch=8)
    net.upload_vit(g["flat"])
    want = net.forward(imgs)
    netcuda.host_register(imgs)
    netcuda.host_register(imgs)
    np.testing.assert_array_equal(net.forward(imgs), want)
    np.testing.assert_array_equal(net.forward(imgs[5:]), want[5:])  # a pointer inside a registered range
    netcuda.host_unregister(imgs)
    np.testing.assert_array_equal(net.forward(imgs), want)
    net.close()
    monkeypatch.setenv("NETCUDA_PIN_INPUTS", "1")
    host = netcuda.HostNet.vit(cfg, g["flat"], max_batch=8)
    s, got = host.time_launch_forward(imgs, reps=3)  # one vector, five calls: page-locked by the first
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(host.launch_forward(imgs), want)  # a vector per call: registered, used, forgotten
    np.testing.assert_array_equal(host.launch_forward(imgs[::-1].copy()), want[::-1])
    np.testing.assert_array_equal(host.launch_forward(imgs[:3]), want[:3])
    host.close()


def test_cpp_class_move_and_copy(netcuda, oracle, torch_cuda):
    npl, n_ins, w, b = c1_net(oracle)
    net = netcuda.HostNet.mlp(npl, n_ins, w, b, precision=-1)  # the reference-shaped 3-argument constructor
    x = np.random.default_rng(1).uniform(-1, 1, n_ins).astype(np.float32)
    y = net.launch_forward(x)[0]
    assert rel_err(y[None], oracle.mlp_forward(x[None], w, b, npl, n_ins)) <= 1e-2  # default precision is TF32
    assert net.check_move_copy(x, y) == 0
    net.close()


# ---- config C5: INT8 wide MLP, bit-exact batch sweep ----------------------------------------------------

def _int8_net(rng, npl, n_ins):
    n_params = sum(a * b for a, b in zip([n_ins] + npl[:-1], npl))
    # weights scaled so that hidden activations neither saturate nor die (about N(0, 20/128) per layer)
    wq = np.clip(np.rint(rng.standard_normal(n_params) * 128.0 / np.sqrt(n_ins) * 1.4), -128, 127).astype(np.int8)
    bq = rng.integers(-2000, 2000, sum(npl), dtype=np.int32)
    return wq, bq


def _assert_same_ints(got, want, what):
    """assert_array_equal with the mismatch pattern (which rows / columns) in the message: it tells a stale tile from a stale row."""
    bad = got != want
    if bad.any():
        rows, cols = np.unique(np.nonzero(bad)[0]), np.unique(np.nonzero(bad)[1])
        raise AssertionError(f"{what}: {int(bad.sum())} of {bad.size} differ; rows {rows.tolist()[:40]} ({rows.size} rows), columns "
                             f"{cols.tolist()[:40]} ({cols.size} columns), max |diff| {int(np.abs(got.astype(np.int64) - want)[bad].max())}")


def test_int8_small_bit_exact_all_activations(netcuda, oracle, torch_cuda):
    rng = np.random.default_rng(31)
    npl, n_ins = [64, 50, 32], 48
    wq, bq = _int8_net(rng, npl, n_ins)
    xq = rng.integers(-128, 128, (19, n_ins), dtype=np.int8)
    for act in (0, 1, 2):
        net = netcuda.Net.mlp(npl, n_ins, precision=netcuda.PREC_INT8, activation=act)
        net.upload_mlp_i8(wq, bq)
        np.testing.assert_array_equal(net.forward_i8(xq), oracle.mlp_forward_i8(xq, wq, bq, npl, n_ins, act))
        net.close()


@pytest.mark.parametrize("batch", [1, 77, 128, 129])
def test_int8_split_k_small_batch_bit_exact(netcuda, oracle, torch_cuda, batch):
    """Up to 128 samples a long-K int8 layer is split along K over the whole GPU (int32 partial sums through TMA reduce-adds,
    then a finalize kernel); 129 samples take the single-pass kernels.  Same integers either way, for every activation mode."""
    rng = np.random.default_rng(40 + batch)
    npl, n_ins = [1000, 2176, 12], 2176  # 17 and 8 k-blocks of 128 bytes; fan-outs that are not tile multiples
    n_params = sum(a * b for a, b in zip([n_ins] + npl[:-1], npl))
    wq = np.clip(np.rint(rng.standard_normal(n_params) * 4), -128, 127).astype(np.int8)
    bq = rng.integers(-3000, 3000, sum(npl), dtype=np.int32)
    xq = rng.integers(-128, 128, (batch, n_ins), dtype=np.int8)
    for act in (0, 1, 2):
        net = netcuda.Net.mlp(npl, n_ins, precision=netcuda.PREC_INT8, activation=act)
        net.upload_mlp_i8(wq, bq)
        want = oracle.mlp_forward_i8(xq, wq, bq, npl, n_ins, act)
        np.testing.assert_array_equal(net.forward_i8(xq), want)
        np.testing.assert_array_equal(net.forward_i8(xq), want)  # the workspace is left zeroed for the next call
        net.set_gemm_variant(2)  # single-pass kernels only
        np.testing.assert_array_equal(net.forward_i8(xq), want)
        net.close()


@pytest.mark.parametrize("npl,n_ins", [([272, 48, 10], 1040), ([64, 32], 4080), ([4096, 304, 4096], 4096), ([16], 48)])
def test_int8_weight_streaming_kernel_bit_exact(netcuda, oracle, torch_cuda, npl, n_ins, monkeypatch):
    """Up to 16 samples (32 with NETCUDA_MLP_STREAM_SPLIT=32: from 17 on the tcgen05 streaming kernel is faster) an INT8 net whose
    fan-ins are multiples of 16 runs as ONE persistent weight-streaming kernel (mlp_stream.cu:
    K split over 8 warps, mma.sync int8, per-CTA output slices; hidden activations exchanged as tagged words up to 4 samples, behind a
    grid barrier above -- and at every batch size with NETCUDA_MLP_STREAM_LL=0).  Same integers as the oracle and as the
    split-K GEMM path, for batches 1..32, ragged output slices (10 or 304 neurons over 148 CTAs), fan-ins that are not a multiple
    of the 32-byte MMA step or leave warps without work, all three activation modes, and across repeated launches (the barrier
    counters reset themselves, the tag epoch advances)."""
    rng = np.random.default_rng(77)
    wq, bq = _int8_net(rng, npl, n_ins)
    for rnd, act in enumerate((0, 1, 2, 0)):
        if rnd == 3:  # last round: the grid barrier at 1..4 samples too, and this kernel up to 32 samples (default hand-over: 16)
            monkeypatch.setenv("NETCUDA_MLP_STREAM_LL", "0")
            monkeypatch.setenv("NETCUDA_MLP_STREAM_SPLIT", "32")
        net = netcuda.Net.mlp(npl, n_ins, precision=netcuda.PREC_INT8, activation=act, max_batch=64)
        net.upload_mlp_i8(wq, bq)
        for batch in (1, 2, 3, 4, 5, 8, 11, 16, 17, 25, 32, 33, 2, 1):
            xq = rng.integers(-128, 128, (batch, n_ins), dtype=np.int8)
            want = oracle.mlp_forward_i8(xq, wq, bq, npl, n_ins, act)
            _assert_same_ints(net.forward_i8(xq), want, f"round {rnd} act {act} batch {batch} first call")
            _assert_same_ints(net.forward_i8(xq), want, f"round {rnd} act {act} batch {batch} second call")
            net.profile_enable(True)  # the kernel really ran (its label shows up in the per-kernel profile)
            net.forward_i8(xq)
            assert ("mlp_stream" in net.profile_read()) == (batch <= (32 if rnd == 3 else 16))
            net.profile_enable(False)
        net.close()
    monkeypatch.delenv("NETCUDA_MLP_STREAM_LL")
    monkeypatch.delenv("NETCUDA_MLP_STREAM_SPLIT")
    monkeypatch.setenv("NETCUDA_MLP_STREAM", "0")  # the split-K GEMM path on the same net: identical integers
    net = netcuda.Net.mlp(npl, n_ins, precision=netcuda.PREC_INT8, max_batch=64)
    net.upload_mlp_i8(wq, bq)
    xq = rng.integers(-128, 128, (7, n_ins), dtype=np.int8)
    net.profile_enable(True)
    got = net.forward_i8(xq)
    assert "mlp_stream" not in net.profile_read()
    net.close()
    np.testing.assert_array_equal(got, oracle.mlp_forward_i8(xq, wq, bq, npl, n_ins))


def _umma_kernels(prof):
    """Which of the two tcgen05 streaming kernels a profile shows: 'cluster' (split-K CTA clusters), 'single', or None."""
    cluster, single = "mlp_umma_stream_cluster" in prof, "mlp_umma_stream" in prof
    assert not (cluster and single)
    return "cluster" if cluster else "single" if single else None


@pytest.mark.parametrize("npl,n_ins", [([272, 48, 10], 1040), ([64, 32], 4080), ([4096, 304, 4096], 4096), ([16], 48), ([4096] * 3, 4096),
                                       ([272, 208, 10], 1040), ([9600, 160], 256), ([400, 1008, 10], 1040), ([9600, 160], 512)])
@pytest.mark.parametrize("pair", [1, 4, 3, 0])
def test_int8_tcgen05_streaming_kernel_bit_exact(netcuda, oracle, torch_cuda, npl, n_ins, pair, monkeypatch):
    """17..128 samples of an INT8 net with 16-byte-aligned fan-ins run as ONE persistent tcgen05 kernel (mlp_umma_stream.cu: weights
    and activations through TMA rings, two MMA-issuing threads that split the K range of a tile -- kind::i8 MMAs of 128 samples x 32
    neurons into an accumulator each in tensor memory -- grid barrier between layers; NETCUDA_MLP_UMMA_MIN moves the hand-over from
    the mma.sync kernel).  Same integers as the oracle for every batch up to 128 (rows past the batch are TMA zero fill), ragged
    neuron tiles (10, 48, 304 neurons), fan-ins that are not a multiple of the 128-byte k-block or leave the second issuer without
    a weight group, all activation modes, repeated launches (the barrier counters reset themselves), and -- with the hand-over point
    moved to zero -- for the small batches the mma.sync kernel normally serves.

    Nets whose every fan-in exceeds one k-block (128 bytes) run as split-K CTA CLUSTERS by default (mlp_i8_umma_cluster_kernel: two
    CTAs per 64-neuron tile -- four per 128-neuron tile when every fan-in has at least four k-blocks -- each with its part of K, partial
    sums exchanged through distributed shared memory as st.async stores counted on the receiver's barrier): K ranges of unequal
    length (1040 = 9 k-blocks, 272 = 3, 400 = 4 with a ragged last one), tiles in which some CTAs have no neuron to finalise (272,
    208, 400, 1008, 10 outputs), more tiles than clusters (9600 neurons).  pair = 1 (the default): clusters of four up to 88 samples where the net
    allows them, pairs with four MMA-issuing threads otherwise; pair = 4: clusters of four at every batch; pair = 3: pairs, two issuers;
    pair = 0 (NETCUDA_MLP_UMMA_PAIR=0): the single-CTA kernel."""
    rng = np.random.default_rng(78)
    wq, bq = _int8_net(rng, npl, n_ins)
    pair_capable = all(f > 128 for f in [n_ins] + npl[:-1])
    if pair != 1 and not pair_capable: pytest.skip("this net runs on the single-CTA kernel anyway (covered by pair = 1)")
    monkeypatch.setenv("NETCUDA_MLP_UMMA_PAIR", str(pair))
    want_kernel = "cluster" if pair and pair_capable else "single"
    for act in (0, 1, 2):
        net = netcuda.Net.mlp(npl, n_ins, precision=netcuda.PREC_INT8, activation=act, max_batch=160)
        net.upload_mlp_i8(wq, bq)
        for batch in (17, 24, 33, 47, 64, 100, 127, 128, 129):
            xq = rng.integers(-128, 128, (batch, n_ins), dtype=np.int8)
            want = oracle.mlp_forward_i8(xq, wq, bq, npl, n_ins, act)
            _assert_same_ints(net.forward_i8(xq), want, f"act {act} batch {batch} first call")
            _assert_same_ints(net.forward_i8(xq), want, f"act {act} batch {batch} second call")
            net.profile_enable(True)
            net.forward_i8(xq)
            assert _umma_kernels(net.profile_read()) == (want_kernel if 17 <= batch <= 128 else None)
            net.profile_enable(False)
        net.close()
    monkeypatch.setenv("NETCUDA_MLP_STREAM_SPLIT", "0")  # the tcgen05 kernel for every batch up to 128
    monkeypatch.setenv("NETCUDA_MLP_UMMA_MIN", "1")
    net = netcuda.Net.mlp(npl, n_ins, precision=netcuda.PREC_INT8, max_batch=64)
    net.upload_mlp_i8(wq, bq)
    for batch in (1, 7, 32, 33, 47):
        xq = rng.integers(-128, 128, (batch, n_ins), dtype=np.int8)
        net.profile_enable(True)
        got = net.forward_i8(xq)
        assert _umma_kernels(net.profile_read()) == want_kernel
        net.profile_enable(False)
        np.testing.assert_array_equal(got, oracle.mlp_forward_i8(xq, wq, bq, npl, n_ins))
    # float API on the same kernel: quantise -> stream -> dequantise
    x = rng.uniform(-1, 1, (40, n_ins)).astype(np.float32)
    acc = oracle.mlp_forward_i8(oracle.quantize_q17(x), wq, bq, npl, n_ins)
    np.testing.assert_array_equal(net.forward(x), acc.astype(np.float32) * np.float32(1.0 / 16384.0))
    net.close()


def test_int8_tcgen05_streaming_narrow_layer_does_not_run_ahead(netcuda, oracle, torch_cuda):
    """Regression (tools/umma_stream_stress.py): 272 -> 48 -> 10 neurons is 9, 2 and 1 output tiles on a 9-CTA grid.  The CTAs without a
    tile in the 48-neuron layer wait for nothing there; with one barrier arrival per CTA and layer their early arrivals stood in for
    CTAs still storing the 272-neuron layer, and the first call after an idle gap read stale bytes of its ragged last tile (2 % of the
    calls).  The barrier now counts finished tiles.  Fresh nets, an oracle run (the idle gap) before every first call."""
    npl, n_ins = [272, 48, 10], 1040
    for rep in range(40):
        rng = np.random.default_rng(500 + rep)
        wq, bq = _int8_net(rng, npl, n_ins)
        net = netcuda.Net.mlp(npl, n_ins, precision=netcuda.PREC_INT8, max_batch=160)
        net.upload_mlp_i8(wq, bq)
        for batch in (100, 127, 128):
            xq = rng.integers(-128, 128, (batch, n_ins), dtype=np.int8)
            want = oracle.mlp_forward_i8(xq, wq, bq, npl, n_ins)
            _assert_same_ints(net.forward_i8(xq), want, f"net {rep} batch {batch}")
        net.close()


def test_int8_float_api_quantises_like_oracle(netcuda, oracle, torch_cuda):
    rng = np.random.default_rng(32)
    npl, n_ins = [96, 40], 72
    n_params = n_ins * 96 + 96 * 40
    w = rng.uniform(-1, 1, n_params).astype(np.float32) * 0.3
    b = rng.uniform(-1, 1, sum(npl)).astype(np.float32) * 0.1
    x = rng.uniform(-1, 1, (11, n_ins)).astype(np.float32)
    net = netcuda.Net.mlp(npl, n_ins, precision=netcuda.PREC_INT8)
    net.upload_mlp(w, b)
    got = net.forward(x)
    net.close()
    acc = oracle.mlp_forward_i8(oracle.quantize_q17(x), oracle.quantize_q17(w), oracle.quantize_bias(b), npl, n_ins)
    np.testing.assert_array_equal(got, acc.astype(np.float32) * np.float32(1.0 / 16384.0))


@pytest.mark.parametrize("batch", [1, 2, 16, 128, 1024, 16384])
def test_c5_int8_wide_mlp_bit_exact(netcuda, oracle, torch_cuda, batch):
    """8 x 4096 INT8.  The oracle checks up to 128 samples per batch (a few seconds of CPU); the rest of a
    large batch is covered by a size-independent property: rows are independent, so forwarding the batch
    must equal forwarding its permutation un-permuted, and duplicated inputs must give duplicated outputs."""
    rng = np.random.default_rng(50)
    npl, n_ins = [4096] * 8, 4096
    wq, bq = _int8_net(rng, npl, n_ins)
    net = netcuda.Net.mlp(npl, n_ins, precision=netcuda.PREC_INT8, max_batch=16384)
    net.upload_mlp_i8(wq, bq)
    xq = rng.integers(-128, 128, (batch, n_ins), dtype=np.int8)
    got = net.forward_i8(xq)
    probe = np.unique(np.concatenate([np.arange(min(batch, 64)), rng.integers(0, batch, 64)]))
    want = oracle.mlp_forward_i8(xq[probe], wq, bq, npl, n_ins)
    np.testing.assert_array_equal(got[probe], want)
    assert np.abs(want).max() > 1000  # the net is alive (not all-zero activations)
    if batch > 128:
        perm = rng.permutation(batch)
        np.testing.assert_array_equal(net.forward_i8(xq[perm])[np.argsort(perm)], got)
    net.close()


# ---- ViT ----------------------------------------------------------------------------------------------

def _golden_vit():
    g = np.load(os.path.join(GOLDEN, "vit_small.npz"))
    cfg = dict(zip(("image_size", "patch_size", "dim", "depth", "heads", "mlp_dim", "n_classes"), [int(v) for v in g["cfg"]]))
    return g, cfg


def test_vit_golden_torchvision_fixture(netcuda, torch_cuda):
    """Logits of torchvision's VisionTransformer (tests/golden/make_golden.py) on the same weights/inputs."""
    from bf16_pipeline_model import vit_forward_bf16_model

    g, cfg = _golden_vit()
    net = netcuda.Net.vit(cfg)
    net.upload_vit(g["flat"])
    got = net.forward(g["images"].reshape(len(g["images"]), -1))
    net.close()
    model = vit_forward_bf16_model(cfg, g["flat"], g["images"])
    assert rel_err(got, model) <= 2e-3        # kernel arithmetic vs the same roundings on the CPU
    assert rel_err(got, g["logits"]) <= 2e-2  # bf16 operand budget on a 128-wide net (1.14e-2 measured and modelled)
    np.testing.assert_array_equal(got.argmax(1), g["logits"].argmax(1))


def test_vit_cpp_class(netcuda, torch_cuda):
    g, cfg = _golden_vit()
    net = netcuda.HostNet.vit(cfg, g["flat"])
    got = net.launch_forward(g["images"])
    net.close()
    assert rel_err(got, g["logits"]) <= 2e-2
    np.testing.assert_array_equal(got.argmax(1), g["logits"].argmax(1))


@pytest.mark.parametrize("name,batch,max_batch", [("vit_tiny_16_224", 5, 2), ("vit_base_16_224", 5, 3)])
def test_vit_real_shapes_vs_oracle(netcuda, oracle, torch_cuda, name, batch, max_batch):
    """197-token configurations at FULL depth (12 blocks): ViT-Tiny (config C2) and ViT-B/16-224 (config C3, the headline) against
    the CPU oracle, with max_batch forcing several internal passes, the last one partial.  The rounding of 12 blocks compounds in
    the fp32 residual stream; the bar stays the north_star's: <= 1e-2 of max |logit| per image and identical top-1."""
    cfg = dict(netcuda.VIT_PRESETS[name])
    assert cfg["depth"] == 12
    flat = netcuda.vit_random_params(cfg, seed=2)
    x = np.random.default_rng(6).uniform(-1, 1, (batch, 3 * cfg["image_size"] ** 2)).astype(np.float32)
    want = oracle.vit_forward(cfg, flat, x)
    net = netcuda.Net.vit(cfg, max_batch=max_batch)
    net.upload_vit(flat)
    got = net.forward(x)
    assert rel_err(got, want) <= 1e-2
    np.testing.assert_array_equal(got.argmax(1), want.argmax(1))
    # tcgen05 kernels == CUDA-core GEMM + mma.sync attention with the same operand types (differences: accumulation
    # order, and the flash kernel rounds P relative to a running max) -- both inside the bf16 budget
    net.set_gemm_variant(1)
    got_ref = net.forward(x)
    net.close()
    assert rel_err(got, got_ref) <= 1e-2
    np.testing.assert_array_equal(got.argmax(1), got_ref.argmax(1))


@pytest.mark.parametrize("depth,batch", [(1, 2), (24, 2)])
def test_vit_large_sequence_577(netcuda, oracle, torch_cuda, depth, batch):
    """ViT-L/16-384 (config C4: 577 tokens, 16 heads, D = 1024) at depth 1 (fast) and at its FULL depth of 24 blocks, one image per
    internal pass, against the CPU oracle (about a minute of host time at depth 24): <= 1e-2 and identical top-1."""
    cfg = dict(netcuda.VIT_PRESETS["vit_large_16_384"], depth=depth)
    flat = netcuda.vit_random_params(cfg, seed=4)
    x = np.random.default_rng(7).uniform(-1, 1, (batch, 3 * 384 * 384)).astype(np.float32)
    want = oracle.vit_forward(cfg, flat, x)
    net = netcuda.Net.vit(cfg, max_batch=1 if depth > 1 else 2)
    net.upload_vit(flat)
    got = net.forward(x)
    net.close()
    assert rel_err(got, want) <= 1e-2
    np.testing.assert_array_equal(got.argmax(1), want.argmax(1))


@pytest.mark.parametrize("name,batch,depth", [("vit_tiny_16_224", 5, 12), ("vit_base_16_224", 3, 12)])
def test_vit_tf32_vs_oracle(netcuda, oracle, torch_cuda, name, batch, depth):
    """NETCUDA_PREC_TF32 vision transformers (config C2: "FP32 reference vs TF32/BF16 netCUDA"; the reference's DATA_TYPE is float,
    def/defines.h:10): fp32 weights and activations, kind::tf32 MMAs in every linear layer, bf16 operands in the attention core
    only.  Same bar as bf16 -- <= 1e-2 of max |logit| and identical top-1 -- and measurably closer to the fp32 oracle than bf16."""
    cfg = dict(netcuda.VIT_PRESETS[name], depth=depth)
    flat = netcuda.vit_random_params(cfg, seed=2)
    x = np.random.default_rng(6).uniform(-1, 1, (batch, 3 * cfg["image_size"] ** 2)).astype(np.float32)
    want = oracle.vit_forward(cfg, flat, x)
    errs = {}
    for prec in ("tf32", "bf16"):
        net = netcuda.Net.vit(cfg, max_batch=2, precision=netcuda.PRECISIONS[prec])
        net.upload_vit(flat)
        got = net.forward(x)
        errs[prec] = rel_err(got, want)
        np.testing.assert_array_equal(got.argmax(1), want.argmax(1))
        if prec == "tf32":
            net.set_gemm_variant(1)  # CUDA-core GEMMs with tf32-truncated operands + mma.sync attention writing fp32
            assert rel_err(net.forward(x), want) <= 1e-2
        net.close()
    assert errs["tf32"] <= 1e-2 and errs["bf16"] <= 1e-2
    assert errs["tf32"] < errs["bf16"], errs  # the linear layers keep 3 more mantissa bits per operand


def test_vit_tf32_cpp_class_and_u8(netcuda, torch_cuda):
    """The C++ class with net_cuda_options::precision = PREC_TF32 on the golden ViT: u8 frames and float images agree bit for bit,
    and the 128-wide fixture that costs bf16 1.14e-2 of operand rounding stays well inside 1e-2."""
    g, cfg = _golden_vit()
    h = netcuda.HostNet.vit(cfg, g["flat"], precision=netcuda.PREC_TF32)
    got = h.launch_forward(g["images"])
    assert rel_err(got, g["logits"]) <= 1e-2
    np.testing.assert_array_equal(got.argmax(1), g["logits"].argmax(1))
    frame = np.random.default_rng(4).integers(0, 256, (32, 32, 3), dtype=np.uint8)
    want = h.launch_forward(_frames_to_float(frame[None], (0.5,) * 3, (0.5,) * 3).ravel()).ravel()
    np.testing.assert_array_equal(h.launch_forward_frame(frame, 32, 32), want)
    h.close()


def test_vit_batch_independence_at_full_batch(netcuda, torch_cuda):
    """Size-independent property at a bench-sized batch: logits of image i do not depend on its neighbours."""
    torch = torch_cuda
    cfg = netcuda.VIT_PRESETS["vit_tiny_16_224"]
    flat = netcuda.vit_random_params(cfg, seed=1)
    net = netcuda.Net.vit(cfg, max_batch=256)
    net.upload_vit(flat)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.rand((256, net.n_in), generator=g, device="cuda") * 2 - 1
    y = torch.empty((256, 1000), device="cuda")
    net.forward_device(x, y, 256, torch.cuda.current_stream())
    perm = torch.randperm(256, device="cuda")
    y2 = torch.empty_like(y)
    net.forward_device(x[perm].contiguous(), y2, 256, torch.cuda.current_stream())
    torch.cuda.synchronize()
    net.close()
    assert torch.equal(y2, y[perm])


# ---- runtime plumbing -------------------------------------------------------------------------------------

def test_forward_device_runs_on_the_stream_it_is_given(netcuda, oracle, torch_cuda):
    """Regression test: torch's default stream has the raw handle 0, which the C ABI reads as "the handle's own stream".
    The binding must pass it as cudaStreamLegacy, otherwise work and the caller's events / reads live on different streams."""
    torch = torch_cuda
    npl, n_ins, w, b = c1_net(oracle)
    net = netcuda.Net.mlp(npl, n_ins, precision=netcuda.PREC_FP32, max_batch=4096)
    net.upload_mlp(w, b)
    x_src = torch.rand((4096, n_ins), device="cuda") * 2 - 1
    want = oracle.mlp_forward(x_src.cpu().numpy(), w, b, npl, n_ins)
    for stream in (torch.cuda.current_stream(), torch.cuda.Stream()):
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            x = torch.zeros_like(x_src)
            y = torch.zeros((4096, 10), device="cuda")
            spin = torch.empty(1 << 26, device="cuda")
            for _ in range(20):
                spin.normal_()    # keeps `stream` busy ...
            x.copy_(x_src)        # ... so the input appears late: a forward on any other stream would read zeros
            net.forward_device(x, y, 4096, stream)
            got = y.cpu().numpy()  # stream-ordered read-back on the same stream
        np.testing.assert_array_equal(got, want)
    net.close()


def test_per_kernel_profile_api(netcuda, torch_cuda):
    torch = torch_cuda
    cfg = dict(netcuda.VIT_PRESETS["vit_tiny_16_224"], depth=2)
    net = netcuda.Net.vit(cfg, max_batch=8)
    net.upload_vit(netcuda.vit_random_params(cfg, seed=3))
    x = torch.rand((8, net.n_in), device="cuda")
    y = torch.empty((8, 1000), device="cuda")
    s = torch.cuda.Stream()
    l0 = net.launches
    net.forward_device(x, y, 8, s)
    s.synchronize()
    per_pass = net.launches - l0
    assert per_pass == 3 + 2 * 7 + 2  # patchify, patch-embed, cls rows; 7 kernels per block; final LayerNorm, head
    assert net.profile_read() == {}   # nothing is recorded while profiling is off
    net.profile_enable(True)
    net.forward_device(x, y, 8, s)
    prof = net.profile_read()
    net.profile_enable(False)
    assert sum(v["launches"] for v in prof.values()) == per_pass
    assert prof["layernorm"]["launches"] == 4 and prof["attention"]["launches"] == 2
    assert all(v["ms"] > 0 for v in prof.values())
    D, F, T = cfg["dim"], cfg["mlp_dim"], 197
    assert prof["fc1"]["flops"] == 2 * (2.0 * 8 * T * D * F)
    flops = sum(v["flops"] for v in prof.values())
    assert abs(flops - 8 * net.flops_per_sample) / flops < 1e-9  # the labels add up to the advertised FLOPs per sample
    assert net.profile_read() == {}
    net.close()


def test_small_mlp_pass_graph_replay(netcuda, oracle, torch_cuda):
    """Small passes of deeper MLPs are replayed from a CUDA graph while the caller presents the same buffers; new buffers, a new
    batch size or the default (uncapturable) stream must all still give the oracle's integers."""
    torch = torch_cuda
    rng = np.random.default_rng(77)
    npl, n_ins = [256, 192, 128, 96, 64], 320
    n_params = sum(a * b for a, b in zip([n_ins] + npl[:-1], npl))
    wq = np.clip(np.rint(rng.standard_normal(n_params) * 6), -128, 127).astype(np.int8)
    bq = rng.integers(-3000, 3000, sum(npl), dtype=np.int32)
    net = netcuda.Net.mlp(npl, n_ins, precision=netcuda.PREC_INT8)
    net.upload_mlp_i8(wq, bq)
    side = torch.cuda.Stream()
    bufs = [(torch.from_numpy(rng.integers(-128, 128, (b, n_ins), dtype=np.int8)).cuda(), torch.empty((b, 64), dtype=torch.int32, device="cuda"))
            for b in (33, 33, 200)]
    torch.cuda.synchronize()
    for stream in (side, torch.cuda.current_stream()):
        for rep in range(3):
            for x, y in bufs:  # alternating buffers: capture, replace, capture again ...
                with torch.cuda.stream(stream):
                    y.zero_()
                    net.forward_device_i8(x, y, x.shape[0], stream)
                    net.forward_device_i8(x, y, x.shape[0], stream)  # ... and an immediate replay
                    got = y.cpu().numpy()
                np.testing.assert_array_equal(got, oracle.mlp_forward_i8(x.cpu().numpy(), wq, bq, npl, n_ins))
    x, y = bufs[0]
    x.copy_(torch.from_numpy(rng.integers(-128, 128, (33, n_ins), dtype=np.int8)).cuda())  # same buffer, new contents
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        net.forward_device_i8(x, y, 33, side)
        got = y.cpu().numpy()
    np.testing.assert_array_equal(got, oracle.mlp_forward_i8(x.cpu().numpy(), wq, bq, npl, n_ins))
    net.close()


def test_small_vit_pass_graph_replay(netcuda, torch_cuda):
    """ViT passes of a few samples are replayed from a CUDA graph once the same (buffers, batch) combination shows up twice in a row
    (single-sample latency is the host's launch rate otherwise).  Plain launch, capture and replays return the same bits; new
    contents in the same buffers, other batch sizes, other buffers and the host-buffer API (which cycles through its staging
    slots) all stay correct."""
    torch = torch_cuda
    g = np.load(os.path.join(GOLDEN, "vit_small.npz"))
    cfg = dict(zip(("image_size", "patch_size", "dim", "depth", "heads", "mlp_dim", "n_classes"), (int(v) for v in g["cfg"])))
    net = netcuda.Net.vit(cfg, max_batch=8)
    net.upload_vit(g["flat"])
    rng = np.random.default_rng(8)
    stream = torch.cuda.Stream()
    ref = {}
    with torch.cuda.stream(stream):
        bufs = {n: (torch.empty((n, net.n_in), device="cuda"), torch.empty((n, net.n_out), device="cuda")) for n in (1, 3, 8)}
        for rep in range(5):  # rep 0: plain, rep 1: capture, rep >= 2: replay -- per batch size, with fresh contents every time
            for n, (x, y) in bufs.items():
                hx = rng.uniform(-1, 1, (n, net.n_in)).astype(np.float32)
                x.copy_(torch.from_numpy(hx).cuda())
                l0 = net.launches
                net.forward_device(x, y, n, stream)
                stream.synchronize()
                assert net.launches > l0  # (the launch counter keeps counting kernels inside a replayed graph)
                if rep == 0:
                    ref[n] = (hx, y.cpu().numpy().copy())
                if rep == 4:  # the first contents again: the replay must reproduce the plain-launch bits
                    x.copy_(torch.from_numpy(ref[n][0]).cuda())
                    net.forward_device(x, y, n, stream)
                    stream.synchronize()
                    np.testing.assert_array_equal(y.cpu().numpy(), ref[n][1])
    # one sample per call through the host API, as the reference's launch_forward does it
    x1 = ref[1][0]
    for _ in range(10):
        np.testing.assert_array_equal(net.forward(x1), ref[1][1])
    net.close()


def test_async_submit_wait(netcuda, oracle, torch_cuda):
    """netcuda_submit / netcuda_wait: calls in flight give exactly what the blocking call gives -- waited out of order,
    with more submits than ring slots, with pinned and pageable buffers, with batches of one pass and of several."""
    torch = torch_cuda
    rng = np.random.default_rng(21)
    npl, n_ins = [96, 48, 12], 200
    w = rng.uniform(-1, 1, 200 * 96 + 96 * 48 + 48 * 12).astype(np.float32)
    b = rng.uniform(-1, 1, sum(npl)).astype(np.float32)
    net = netcuda.Net.mlp(npl, n_ins, precision=netcuda.PREC_FP32, max_batch=64)
    net.upload_mlp(w, b)
    batches = [1, 64, 65, 200, 7, 130, 64, 3, 500]
    xs = [rng.uniform(-1, 1, (n, n_ins)).astype(np.float32) for n in batches]
    want = [oracle.mlp_forward(x, w, b, npl, n_ins) for x in xs]
    # pinned buffers, all nine in flight at once (ring of 4: the older ones are retired by later submits)
    px = [torch.from_numpy(x).pin_memory() for x in xs]
    py = [torch.empty((x.shape[0], 12), dtype=torch.float32).pin_memory() for x in xs]
    tickets = [net.submit(a, o) for a, o in zip(px, py)]
    assert tickets == list(range(tickets[0], tickets[0] + len(xs)))
    for i in (8, 2, 5, 0, 1, 3, 4, 6, 7):
        net.wait(tickets[i])
        assert net.query(tickets[i])
        np.testing.assert_array_equal(py[i].numpy(), want[i])
    net.wait(tickets[4])  # waiting twice is harmless
    # pageable numpy buffers
    outs = [np.empty((x.shape[0], 12), dtype=np.float32) for x in xs]
    tickets = [net.submit(x, o) for x, o in zip(xs[:4], outs[:4])]
    for t, o, y in zip(tickets, outs, want):
        net.wait(t)
        np.testing.assert_array_equal(o, y)
    # a blocking call between asynchronous ones
    t = net.submit(px[3], py[3].zero_())
    np.testing.assert_array_equal(net.forward(xs[5]), want[5])
    net.wait(t)
    np.testing.assert_array_equal(py[3].numpy(), want[3])
    with pytest.raises(netcuda.NetcudaError):
        net.wait(10 ** 9)
    with pytest.raises(netcuda.NetcudaError):
        net.wait(0)
    assert net.last_forward_us > 0
    net.close()


def test_async_vit_pipeline_matches_blocking(netcuda, torch_cuda):
    """Two ViT calls in flight (the bench's e2e loop): bit-identical logits to the blocking call."""
    torch = torch_cuda
    g = np.load(os.path.join(GOLDEN, "vit_small.npz"))
    cfg = dict(zip(("image_size", "patch_size", "dim", "depth", "heads", "mlp_dim", "n_classes"), (int(v) for v in g["cfg"])))
    net = netcuda.Net.vit(cfg, max_batch=8)
    net.upload_vit(g["flat"])
    rng = np.random.default_rng(5)
    xs = [torch.from_numpy(rng.uniform(-1, 1, (n, net.n_in)).astype(np.float32)).pin_memory() for n in (20, 8, 33)]
    want = [net.forward(x.numpy()) for x in xs]
    ys = [torch.empty((x.shape[0], net.n_out)).pin_memory() for x in xs]
    prev = None
    for rep in range(3):
        for x, y in zip(xs, ys):
            t = net.submit(x, y)
            if prev is not None:
                net.wait(prev)
            prev = t
    net.wait(prev)
    for y, wnt in zip(ys, want):
        np.testing.assert_array_equal(y.numpy(), wnt)
    net.close()


def test_host_call_quarter_first_chunk_is_bit_equal(netcuda, torch_cuda):
    """A large ViT batch handed to an idle GPU starts with a quarter-size first chunk (its H2D copy is the only exposed one); chunking
    never changes a logit: blocking call, two calls in flight and the device-resident path agree bit for bit."""
    torch = torch_cuda
    cfg = dict(image_size=32, patch_size=16, dim=128, depth=1, heads=2, mlp_dim=256, n_classes=10)
    net = netcuda.Net.vit(cfg, max_batch=512)
    net.upload_vit(netcuda.vit_random_params(cfg, seed=3))
    rng = np.random.default_rng(6)
    x = torch.from_numpy(rng.uniform(-1, 1, (600, net.n_in)).astype(np.float32)).pin_memory()  # chunks of 128 + 472 (idle) or 512 + 88
    dx, dy = x.cuda(), torch.empty((600, net.n_out), device="cuda")
    net.forward_device(dx, dy, 600)
    torch.cuda.synchronize()
    want = dy.cpu().numpy()
    np.testing.assert_array_equal(net.forward(x.numpy()), want)
    ys = [torch.empty((600, net.n_out)).pin_memory() for _ in range(2)]
    t0 = net.submit(x, ys[0])
    t1 = net.submit(x, ys[1])  # the GPU is busy with t0: full-size first chunk
    net.wait(t0), net.wait(t1)
    np.testing.assert_array_equal(ys[0].numpy(), want)
    np.testing.assert_array_equal(ys[1].numpy(), want)
    net.close()


def _frames_to_float(frames, mean, std):
    """What patchify_u8_kernel computes, in numpy fp32 with the same operation order, as the CHW float images of the float path."""
    u = frames.astype(np.float32) * np.float32(1.0 / 255.0)
    inv = (np.float32(1.0) / np.asarray(std, dtype=np.float32)).astype(np.float32)
    v = (u - np.asarray(mean, dtype=np.float32)) * inv
    return np.ascontiguousarray(v.astype(np.float32).transpose(0, 3, 1, 2))


def test_vit_u8_frames_equal_float_path(netcuda, torch_cuda):
    """u8 HWC frames (the reference's image carrier, def/defines.h:31-38) through netcuda_forward_u8 give bit for bit the logits of
    the float path on the identically normalised images: the fused u8 patchify rounds exactly like host fp32 arithmetic."""
    torch = torch_cuda
    g = np.load(os.path.join(GOLDEN, "vit_small.npz"))
    cfg = dict(zip(("image_size", "patch_size", "dim", "depth", "heads", "mlp_dim", "n_classes"), (int(v) for v in g["cfg"])))
    net = netcuda.Net.vit(cfg, max_batch=16)
    net.upload_vit(g["flat"])
    rng = np.random.default_rng(9)
    frames = rng.integers(0, 256, (37, 32, 32, 3), dtype=np.uint8)  # 37 frames: three passes, the last one ragged
    frames[0] = 0; frames[1] = 255
    want = net.forward(_frames_to_float(frames, (0.5, 0.5, 0.5), (0.5, 0.5, 0.5)).reshape(37, -1))
    np.testing.assert_array_equal(net.forward_u8(frames), want)
    # device-resident and asynchronous forms
    d_f = torch.from_numpy(frames).cuda(); d_y = torch.empty((37, net.n_out), device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        net.forward_device_u8(d_f, d_y, 37, s)
    s.synchronize()
    np.testing.assert_array_equal(d_y.cpu().numpy(), want)
    pf = torch.from_numpy(frames).pin_memory(); py = torch.empty((37, net.n_out)).pin_memory()
    net.wait(net.submit_u8(pf, py))
    np.testing.assert_array_equal(py.numpy(), want)
    # ImageNet statistics
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    net.set_u8_normalization(mean, std)
    want2 = net.forward(_frames_to_float(frames, mean, std).reshape(37, -1))
    np.testing.assert_array_equal(net.forward_u8(frames), want2)
    assert np.abs(want2 - want).max() > 1e-3
    with pytest.raises(netcuda.NetcudaError):
        net.set_u8_normalization(mean, (0.2, 0.0, 0.2))
    net.close()
    mlp = netcuda.Net.mlp([4], 3072, precision=netcuda.PREC_FP32)
    with pytest.raises(netcuda.NetcudaError, match="ViT"):
        mlp.forward_u8(frames[:1])
    mlp.close()


def test_u8_normalization_change_invalidates_captured_graphs(netcuda, torch_cuda):
    """Regression: mean / inv_std are by-value kernel arguments, so a captured pass has them baked in.  With
    batch <= max_batch the same (staging slot, output) pair repeats and the pass is replayed from a CUDA graph; a later
    netcuda_set_u8_normalization must drop those graphs or the new statistics are silently ignored."""
    torch = torch_cuda
    g, cfg = _golden_vit()
    net = netcuda.Net.vit(cfg, max_batch=16)
    net.upload_vit(g["flat"])
    frames = np.random.default_rng(10).integers(0, 256, (8, 32, 32, 3), dtype=np.uint8)
    want1 = net.forward(_frames_to_float(frames, (0.5,) * 3, (0.5,) * 3).reshape(8, -1))
    d_f = torch.from_numpy(frames).cuda()
    d_y = torch.empty((8, net.n_out), device="cuda")
    s = torch.cuda.Stream()
    for _ in range(6):  # plain, capture, replays
        np.testing.assert_array_equal(net.forward_u8(frames), want1)
        net.forward_device_u8(d_f, d_y, 8, s)
        s.synchronize()
        np.testing.assert_array_equal(d_y.cpu().numpy(), want1)
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    net.set_u8_normalization(mean, std)
    want2 = net.forward(_frames_to_float(frames, mean, std).reshape(8, -1))
    assert np.abs(want2 - want1).max() > 1e-3
    for _ in range(6):
        np.testing.assert_array_equal(net.forward_u8(frames), want2)
        net.forward_device_u8(d_f, d_y, 8, s)
        s.synchronize()
        np.testing.assert_array_equal(d_y.cpu().numpy(), want2)
    net.close()


def test_class_launch_forward_image_set(netcuda, torch_cuda):
    """cuda::net_cuda::launch_forward(const net::image_set&): one frame in the reference's carrier type."""
    g = np.load(os.path.join(GOLDEN, "vit_small.npz"))
    cfg = dict(zip(("image_size", "patch_size", "dim", "depth", "heads", "mlp_dim", "n_classes"), (int(v) for v in g["cfg"])))
    h = netcuda.HostNet.vit(cfg, g["flat"])
    frame = np.random.default_rng(4).integers(0, 256, (32, 32, 3), dtype=np.uint8)
    want = h.launch_forward(_frames_to_float(frame[None], (0.5,) * 3, (0.5,) * 3).ravel()).ravel()
    np.testing.assert_array_equal(h.launch_forward_frame(frame, 32, 32), want)
    np.testing.assert_array_equal(h.launch_forward_frame(frame), want)  # original_h / original_w left at 0
    with pytest.raises(RuntimeError, match="image_size"):
        h.launch_forward_frame(frame[:16])
    with pytest.raises(RuntimeError, match="original_h"):
        h.launch_forward_frame(frame, 64, 32)
    h.close()
