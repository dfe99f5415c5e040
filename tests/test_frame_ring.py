"""The image side channel: net_abstract::filter_image / get_filtered_image (reference: src/netFPGA.cpp:292-365).

The reference's device stage `image_process` is absent (no source, no bitstream); what its host code pins is the ring contract:
one byte per pixel in and out, original_h * original_w bytes per frame (:314-315, :441-442), 24 slots (BATCH_SIZE, :12), FIFO order
(:322, :355), a full ring drops the frame (:333), an empty ring returns a 1080 x 1920 header without pixels (:340-342, :359).  The
filter itself is a builder decision (3 x 3 binomial smoothing in integers, oracle/oracle_image.c), hence bit-exact.
"""
import numpy as np
import pytest


def _numpy_filter(img):
    """Independent restatement in numpy (the C oracle is checked against it on CPU)."""
    p = np.pad(img.astype(np.int32), 1, mode="edge")
    acc = np.zeros(img.shape, np.int32)
    for dy, wy in zip(range(3), (1, 2, 1)):
        for dx, wx in zip(range(3), (1, 2, 1)):
            acc += wy * wx * p[dy:dy + img.shape[0], dx:dx + img.shape[1]]
    return ((acc + 8) >> 4).astype(np.uint8)


def test_oracle_filter_matches_numpy_restatement(oracle):
    rng = np.random.default_rng(3)
    for h, w in ((1, 1), (1, 7), (5, 1), (3, 3), (17, 31), (64, 64), (1080, 1920)):
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        np.testing.assert_array_equal(oracle.filter3x3(img), _numpy_filter(img))
    flat = np.full((9, 13), 200, np.uint8)
    np.testing.assert_array_equal(oracle.filter3x3(flat), flat)  # weights sum to 16: constants are fixed points
    impulse = np.zeros((5, 5), np.uint8)
    impulse[2, 2] = 160
    want = np.zeros((5, 5), np.uint8)
    want[1:4, 1:4] = np.array([[10, 20, 10], [20, 40, 20], [10, 20, 10]])
    np.testing.assert_array_equal(oracle.filter3x3(impulse), want)


@pytest.mark.gpu
def test_filter_kernel_bit_exact(netcuda, oracle, torch_cuda):
    torch = torch_cuda
    rng = np.random.default_rng(5)
    for h, w in ((1, 1), (2, 3), (7, 5), (33, 1021), (64, 64), (1080, 1920), (333, 4099)):
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        d_in = torch.from_numpy(img).cuda()
        d_out = torch.full((h, w), 77, dtype=torch.uint8, device="cuda")
        netcuda.op_filter3x3(d_in, d_out, h, w, stream=torch.cuda.current_stream())
        torch.cuda.synchronize()
        np.testing.assert_array_equal(d_out.cpu().numpy(), oracle.filter3x3(img))


@pytest.mark.gpu
def test_ring_fifo_full_and_empty(netcuda, oracle, torch_cuda):
    rng = np.random.default_rng(6)
    ring = netcuda.FrameRing(max_pixels=120 * 160, depth=4)
    assert ring.in_flight == 0
    with pytest.raises(netcuda.NetcudaError) as e:
        ring.pop()
    assert e.value.code == netcuda.ERR_RING_EMPTY  # "PILA VACIA"
    frames = [rng.integers(0, 256, (h, w), dtype=np.uint8) for h, w in ((120, 160), (60, 80), (1, 19), (77, 3), (120, 160), (9, 9))]
    for f in frames[:4]:
        ring.push(f)
    assert ring.in_flight == 4
    with pytest.raises(netcuda.NetcudaError) as e:
        ring.push(frames[4])  # every slot in flight: the frame is refused ("PILA LLENA")
    assert e.value.code == netcuda.ERR_RING_FULL and ring.dropped == 1
    np.testing.assert_array_equal(ring.pop(), oracle.filter3x3(frames[0]))  # oldest first
    ring.push(frames[4])  # the freed slot is reused (write index wraps)
    for f in frames[1:5]:
        np.testing.assert_array_equal(ring.pop(), oracle.filter3x3(f))
    assert ring.in_flight == 0
    with pytest.raises(netcuda.NetcudaError):
        ring.push(np.zeros((121, 160), np.uint8))  # larger than the ring's slots
    # many laps around the ring, interleaved pushes and pops
    want = []
    for i in range(50):
        f = rng.integers(0, 256, (int(rng.integers(1, 121)), int(rng.integers(1, 161))), dtype=np.uint8)
        ring.push(f)
        want.append(oracle.filter3x3(f))
        if i % 3 != 0:
            np.testing.assert_array_equal(ring.pop(), want.pop(0))
        if ring.in_flight == 4:
            np.testing.assert_array_equal(ring.pop(), want.pop(0))
    while want:
        np.testing.assert_array_equal(ring.pop(), want.pop(0))
    ring.close()


@pytest.mark.gpu
def test_class_filter_image_ring(netcuda, oracle, torch_cuda):
    """cuda::net_cuda::filter_image / get_filtered_image through net::net_abstract*: 24 frames in flight, the 25th is dropped like
    the reference drops it, results come back in order; an empty ring answers with the reference's 1080 x 1920 header."""
    npl, n_ins = [4, 2], 3
    net = netcuda.HostNet.mlp(npl, n_ins, np.zeros(20, np.float32), np.zeros(6, np.float32), precision=netcuda.PREC_FP32)
    pixels, dims = net.get_filtered_image()
    assert pixels is None and dims == (1080, 1920)
    rng = np.random.default_rng(7)
    frames = [rng.integers(0, 256, (1080, 1920), dtype=np.uint8) for _ in range(3)] + \
             [rng.integers(0, 256, (48, 64), dtype=np.uint8) for _ in range(22)]
    for f in frames:
        net.filter_image(f)  # the 25th is dropped silently (src/netFPGA.cpp:331-334)
    for f in frames[:24]:
        pixels, dims = net.get_filtered_image()
        assert dims == f.shape
        np.testing.assert_array_equal(pixels, oracle.filter3x3(f))
    pixels, dims = net.get_filtered_image()
    assert pixels is None and dims == (1080, 1920)
    net.filter_image(frames[24])  # room again
    pixels, _ = net.get_filtered_image()
    np.testing.assert_array_equal(pixels, oracle.filter3x3(frames[24]))
    assert net.check_stubs() == 0
    net.close()
