"""Weight files (include/netcuda.h "weight files"): the flat layout of src/netFPGA.cpp:91-106 on disk.

CPU tests: the file functions are pure host code -- round trips, header validation, corruption and truncation are rejected with a
message instead of undefined behaviour, and the checkpoint converters reproduce the committed torchvision fixture.
GPU tests: a net created from a file computes exactly what the net created from the same arrays computes, through the C ABI and
through cuda::net_cuda::save / load.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, c1_net

VIT_KEYS = ("image_size", "patch_size", "dim", "depth", "heads", "mlp_dim", "n_classes")


def test_mlp_file_round_trip(netcuda, oracle, tmp_path):
    npl, n_ins, w, b = c1_net(oracle)
    path = tmp_path / "c1.ncw"
    netcuda.file_write_mlp(path, npl, n_ins, w, b, activation=netcuda.ACT_RELU_ALL)
    info = netcuda.file_info(path)
    assert info["kind"] == netcuda.KIND_MLP and info["dtype"] == netcuda.FILE_F32 and info["precision"] == netcuda.PREC_FP32
    assert info["npl"] == npl and info["n_ins"] == n_ins and info["activation"] == netcuda.ACT_RELU_ALL
    assert info["n_weights"] == w.size and info["n_biases"] == b.size
    # header 96 + n_p_l padded to 16 + payload sections padded to 16 (the documented layout)
    assert os.path.getsize(path) == 96 + 16 + (w.nbytes + 15) // 16 * 16 + (b.nbytes + 15) // 16 * 16
    w2, b2 = netcuda.file_read(path)
    assert np.array_equal(w2, w) and np.array_equal(b2, b)
    raw = open(path, "rb").read()
    assert raw[:8] == b"NETCUDAW" and np.frombuffer(raw[96:96 + 12], dtype=np.int32).tolist() == npl
    assert np.array_equal(np.frombuffer(raw[112:112 + w.nbytes], dtype=np.float32), w)  # the reference's flat order, verbatim


def test_q17_file_round_trip(netcuda, tmp_path):
    rng = np.random.default_rng(3)
    npl, n_ins = [24, 17, 5], 33  # ragged sizes: the padding rules must hold for any byte count
    wq = rng.integers(-128, 128, 33 * 24 + 24 * 17 + 17 * 5, dtype=np.int8)
    bq = rng.integers(-(1 << 20), 1 << 20, sum(npl), dtype=np.int32)
    path = tmp_path / "q.ncw"
    netcuda.file_write_mlp_i8(path, npl, n_ins, wq, bq)
    info = netcuda.file_info(path)
    assert info["dtype"] == netcuda.FILE_Q17 and info["precision"] == netcuda.PREC_INT8 and info["npl"] == npl
    w2, b2 = netcuda.file_read(path)
    assert w2.dtype == np.int8 and b2.dtype == np.int32 and np.array_equal(w2, wq) and np.array_equal(b2, bq)


def test_vit_file_round_trip(netcuda, tmp_path):
    g = np.load(os.path.join(GOLDEN, "vit_small.npz"))
    cfg = dict(zip(VIT_KEYS, (int(v) for v in g["cfg"])))
    path = tmp_path / "vit.ncw"
    netcuda.file_write_vit(path, cfg, g["flat"])
    info = netcuda.file_info(path)
    assert info["kind"] == netcuda.KIND_VIT and info["cfg"] == cfg and info["n_biases"] == 0 and info["precision"] == netcuda.PREC_BF16
    flat, b = netcuda.file_read(path)
    assert b.size == 0 and np.array_equal(flat, g["flat"])
    with pytest.raises(netcuda.NetcudaError, match="floats"):
        netcuda.file_write_vit(tmp_path / "bad.ncw", cfg, g["flat"][:-1])


def test_damaged_files_are_rejected(netcuda, oracle, tmp_path):
    npl, n_ins, w, b = c1_net(oracle)
    path = tmp_path / "c1.ncw"
    netcuda.file_write_mlp(path, npl, n_ins, w, b)
    raw = bytearray(open(path, "rb").read())

    def variant(name, data):
        p = tmp_path / name
        open(p, "wb").write(bytes(data))
        return p

    flipped = bytearray(raw); flipped[5000] ^= 0x40
    with pytest.raises(netcuda.NetcudaError, match="checksum"):
        netcuda.file_read(variant("flip.ncw", flipped))
    netcuda.file_info(variant("flip2.ncw", flipped))  # the header alone is still fine
    act = bytearray(raw); act[20] ^= 1  # header field `activation` (offset 20): sizes still agree, only the checksum can tell
    netcuda.file_info(variant("act.ncw", act))
    with pytest.raises(netcuda.NetcudaError, match="checksum"):
        netcuda.file_read(variant("act2.ncw", act))
    with pytest.raises(netcuda.NetcudaError, match="bytes on disk"):
        netcuda.file_info(variant("short.ncw", raw[:-16]))
    with pytest.raises(netcuda.NetcudaError, match="bytes on disk"):
        netcuda.file_info(variant("long.ncw", raw + b"\0" * 16))
    with pytest.raises(netcuda.NetcudaError, match="magic"):
        netcuda.file_info(variant("magic.ncw", b"NOTAFILE" + raw[8:]))
    with pytest.raises(netcuda.NetcudaError, match="header"):
        netcuda.file_info(variant("tiny.ncw", raw[:40]))
    ver = bytearray(raw); ver[8] = 9
    with pytest.raises(netcuda.NetcudaError, match="version"):
        netcuda.file_info(variant("ver.ncw", ver))
    lay = bytearray(raw); lay[96] ^= 1  # n_p_l[0] no longer matches the element counts
    with pytest.raises(netcuda.NetcudaError, match="layer table"):
        netcuda.file_info(variant("lay.ncw", lay))
    with pytest.raises(netcuda.NetcudaError, match="cannot open"):
        netcuda.file_info(tmp_path / "missing.ncw")
    with pytest.raises(netcuda.NetcudaError):
        netcuda.file_write_mlp(tmp_path / "x.ncw", [4, 0, 2], 3, np.zeros(20, np.float32), np.zeros(6, np.float32))


def test_create_from_file_fails_loudly_without_gpu(netcuda, oracle, tmp_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    npl, n_ins, w, b = c1_net(oracle)
    path = tmp_path / "c1.ncw"
    netcuda.file_write_mlp(path, npl, n_ins, w, b)
    with pytest.raises(netcuda.NetcudaError) as e:
        netcuda.Net.from_file(path)
    assert e.value.code == netcuda.ERR_NO_DEVICE


def test_torchvision_converter_matches_fixture():
    """netcuda.convert reproduces the flat vector of the committed fixture from a live torchvision model
    (tests/golden/make_golden.py built the fixture with its own flattening code)."""
    pytest.importorskip("torchvision")
    from golden.make_golden import make_torchvision_vit
    from netcuda import convert

    g = np.load(os.path.join(GOLDEN, "vit_small.npz"))
    cfg = dict(zip(VIT_KEYS, (int(v) for v in g["cfg"])))
    sd = make_torchvision_vit(cfg).state_dict()
    assert convert.cfg_from_torchvision(sd) == cfg
    assert np.array_equal(convert.flat_from_torchvision(sd), g["flat"])


def test_hf_converter_matches_oracle(oracle):
    """A randomly initialised transformers.ViTForImageClassification, flattened by netcuda.convert and run through the CPU oracle,
    gives the logits the HF model computes (fp32): pins the q|k|v stacking and every tensor's place in the flat vector."""
    transformers = pytest.importorskip("transformers")
    import torch
    from netcuda import convert

    torch.manual_seed(5)
    hf_cfg = transformers.ViTConfig(image_size=32, patch_size=16, hidden_size=128, num_hidden_layers=2, num_attention_heads=2,
                                    intermediate_size=256, num_labels=7, layer_norm_eps=1e-6, hidden_act="gelu", qkv_bias=True,
                                    hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    model = transformers.ViTForImageClassification(hf_cfg).eval()
    with torch.no_grad():
        for p in model.parameters():  # HF zero-initialises biases / cls: make every tensor count
            if p.abs().max() == 0:
                p.copy_(torch.randn_like(p) * 0.02)
    sd = model.state_dict()
    cfg = convert.cfg_from_hf(sd)
    assert cfg == dict(image_size=32, patch_size=16, dim=128, depth=2, heads=2, mlp_dim=256, n_classes=7)
    flat = convert.flat_from_hf(sd)
    x = np.random.default_rng(0).uniform(-1, 1, (3, 3, 32, 32)).astype(np.float32)
    with torch.no_grad():
        want = model(pixel_values=torch.from_numpy(x)).logits.numpy()
    got = oracle.vit_forward(cfg, flat, x)
    assert np.abs(got - want).max() <= 2e-5 * max(1.0, np.abs(want).max())


# ---- GPU ------------------------------------------------------------------------------------------------------------

@pytest.mark.gpu
def test_net_from_file_equals_net_from_arrays(netcuda, oracle, torch_cuda, tmp_path):
    npl, n_ins, w, b = c1_net(oracle)
    x = np.random.default_rng(1234).uniform(-1, 1, (64, n_ins)).astype(np.float32)
    path = tmp_path / "c1.ncw"
    netcuda.file_write_mlp(path, npl, n_ins, w, b)
    for prec in (netcuda.PREC_FP32, netcuda.PREC_TF32, netcuda.PREC_BF16, netcuda.PREC_INT8):
        a = netcuda.Net.mlp(npl, n_ins, precision=prec); a.upload_mlp(w, b)
        f = netcuda.Net.from_file(path, precision=prec)
        assert f.n_in == n_ins and f.n_out == npl[-1]
        assert np.array_equal(a.forward(x), f.forward(x))
        a.close(); f.close()
    f = netcuda.Net.from_file(path)  # natural precision of an fp32 MLP file: FP32 (DATA_TYPE is float; bit-equal to the oracle)
    a = netcuda.Net.mlp(npl, n_ins, precision=netcuda.PREC_FP32); a.upload_mlp(w, b)
    assert np.array_equal(a.forward(x), f.forward(x))
    golden = np.load(os.path.join(GOLDEN, "mlp_c1.npz"))
    f32 = netcuda.Net.from_file(path, precision=netcuda.PREC_FP32)
    assert np.array_equal(f32.forward(golden["x"]), golden["y"])  # file -> GPU reproduces the reference runtime's outputs
    for n in (a, f, f32): n.close()


@pytest.mark.gpu
def test_q17_file_is_bit_exact_and_int8_only(netcuda, oracle, torch_cuda, tmp_path):
    rng = np.random.default_rng(8)
    npl, n_ins = [256, 128, 64], 512
    wq = np.clip(np.rint(rng.standard_normal(512 * 256 + 256 * 128 + 128 * 64) * 10), -128, 127).astype(np.int8)
    bq = rng.integers(-4000, 4000, sum(npl), dtype=np.int32)
    xq = rng.integers(-128, 128, (77, n_ins), dtype=np.int8)
    path = tmp_path / "q.ncw"
    netcuda.file_write_mlp_i8(path, npl, n_ins, wq, bq)
    net = netcuda.Net.from_file(path)
    assert np.array_equal(net.forward_i8(xq), oracle.mlp_forward_i8(xq, wq, bq, npl, n_ins))
    net.close()
    with pytest.raises(netcuda.NetcudaError, match="INT8"):
        netcuda.Net.from_file(path, precision=netcuda.PREC_BF16)


@pytest.mark.gpu
def test_vit_file_and_class_save_load(netcuda, torch_cuda, tmp_path):
    g = np.load(os.path.join(GOLDEN, "vit_small.npz"))
    cfg = dict(zip(VIT_KEYS, (int(v) for v in g["cfg"])))
    x = g["images"].reshape(4, -1)
    a = netcuda.Net.vit(cfg); a.upload_vit(g["flat"])
    want = a.forward(x)
    path = tmp_path / "vit.ncw"
    netcuda.file_write_vit(path, cfg, g["flat"])
    f = netcuda.Net.from_file(path)
    assert np.array_equal(f.forward(x), want)
    # the C++ class: save() writes the same bytes, load() gives an equal net behind net::net_abstract*
    h = netcuda.HostNet.vit(cfg, g["flat"])
    path2 = tmp_path / "vit2.ncw"
    h.save(path2)
    assert open(path, "rb").read() == open(path2, "rb").read()
    l = netcuda.HostNet.load(path2)
    assert np.array_equal(l.launch_forward(x.ravel()).reshape(4, -1), want)
    for n in (a, f, h, l): n.close()


@pytest.mark.gpu
def test_class_save_load_mlp(netcuda, oracle, torch_cuda, tmp_path):
    npl, n_ins, w, b = c1_net(oracle)
    x = np.random.default_rng(2).uniform(-1, 1, (5, n_ins)).astype(np.float32)
    h = netcuda.HostNet.mlp(npl, n_ins, w, b, precision=netcuda.PREC_FP32, activation=netcuda.ACT_RELU_ALL)
    path = tmp_path / "m.ncw"
    h.save(path)
    assert netcuda.file_info(path)["activation"] == netcuda.ACT_RELU_ALL
    l = netcuda.HostNet.load(path, precision=netcuda.PREC_FP32)
    assert np.array_equal(l.launch_forward(x.ravel()), h.launch_forward(x.ravel()))
    w2, b2 = l.get_net_data(w.size, b.size)[:2]
    assert np.array_equal(w2, w) and np.array_equal(b2, b)
    h.close(); l.close()
