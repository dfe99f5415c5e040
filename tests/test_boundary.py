"""CPU tests of the drop-in boundary: interface headers, C ABI exports, loud failure without a GPU."""
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import REFERENCE_TREE, ROOT

INCLUDE = os.path.join(ROOT, "include")

PROBE = r"""
#include <netAbstract.h>
#include <cstddef>
#include <cstdio>
struct probe_net : net::net_abstract {
    net::net_data get_net_data() override;
    std::vector<DATA_TYPE> launch_forward(const std::vector<DATA_TYPE> &inputs) override;
    void init_gradient(const net::net_sets &sets) override;
    std::vector<DATA_TYPE> launch_gradient(size_t iterations, DATA_TYPE error_threshold, DATA_TYPE multiplier) override;
    void print_inner_vals() override;
    signed long get_gradient_performance() override;
    signed long get_forward_performance() override;
    void filter_image(const net::image_set &set) override;
    net::image_set get_filtered_image() override;
};
net::net_data probe_net::get_net_data() { return net::net_data(); }
std::vector<DATA_TYPE> probe_net::launch_forward(const std::vector<DATA_TYPE> &i) { return i; }
void probe_net::init_gradient(const net::net_sets &) {}
std::vector<DATA_TYPE> probe_net::launch_gradient(size_t n, DATA_TYPE, DATA_TYPE) { return std::vector<DATA_TYPE>(n); }
void probe_net::print_inner_vals() {}
signed long probe_net::get_gradient_performance() { return 7; }
signed long probe_net::get_forward_performance() { return 8; }
void probe_net::filter_image(const net::image_set &) {}
net::image_set probe_net::get_filtered_image() { return net::image_set(); }
void takes_data(const net::net_data &, const net::net_sets &, const net::image_set &) {}
int main() {
    printf("net_data %zu %zu %zu %zu %zu %zu %zu\n", sizeof(net::net_data), offsetof(net::net_data, n_ins), offsetof(net::net_data, n_layers),
           offsetof(net::net_data, n_p_l), offsetof(net::net_data, params), offsetof(net::net_data, bias), offsetof(net::net_data, activations));
    printf("net_sets %zu %zu %zu\n", sizeof(net::net_sets), offsetof(net::net_sets, set_ins), offsetof(net::net_sets, set_outs));
    printf("image_set %zu %zu %zu %zu %zu %zu\n", sizeof(net::image_set), offsetof(net::image_set, resized_image_data),
           offsetof(net::image_set, original_x_pos), offsetof(net::image_set, original_y_pos), offsetof(net::image_set, original_h),
           offsetof(net::image_set, original_w));
    printf("abstract %zu data_type %zu range %g %g\n", sizeof(net::net_abstract), sizeof(DATA_TYPE), (double)net::MAX_RANGE, (double)net::MIN_RANGE);
    // vtable slot order, observed through the Itanium ABI layout: slots 2.. are the 9 virtuals in declaration order
    probe_net p; net::net_abstract *a = &p;
    void **vt = *reinterpret_cast<void ***>(a);
    typedef signed long (*perf_fn)(net::net_abstract *);
    printf("slot7 %ld slot8 %ld\n", reinterpret_cast<perf_fn>(vt[7])(a), reinterpret_cast<perf_fn>(vt[8])(a));
    return 0;
}
"""


def _compile_probe(tmp, tag, include_flags):
    src = os.path.join(tmp, f"probe_{tag}.cpp")
    exe = os.path.join(tmp, f"probe_{tag}")
    open(src, "w").write(PROBE)
    subprocess.run(["/usr/bin/g++", "-std=gnu++14", "-O0", "-w"] + include_flags + [src, "-o", exe], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    syms = subprocess.run(["nm", exe], check=True, capture_output=True, text=True).stdout
    mangled = sorted({line.split()[-1] for line in syms.splitlines() if "probe_net" in line or "takes_data" in line})
    return out, mangled


def test_interface_headers_compile_and_have_expected_layout():
    with tempfile.TemporaryDirectory() as tmp:
        out, mangled = _compile_probe(tmp, "ours", ["-I", INCLUDE])
    assert "slot7 7 slot8 8" in out  # get_gradient_performance / get_forward_performance sit in vtable slots 7 / 8
    assert "net_data 112 0 8 16 40 64 88" in out
    assert "_Z10takes_dataRKN3net8net_dataERKNS_8net_setsERKNS_9image_setE" in mangled


@pytest.mark.skipif(not os.path.isdir(REFERENCE_TREE), reason="reference tree not present")
def test_interface_headers_abi_identical_to_reference():
    ref_flags = ["-include", "stddef.h", "-I", os.path.join(REFERENCE_TREE, "include"), "-I", os.path.join(REFERENCE_TREE, "def")]
    with tempfile.TemporaryDirectory() as tmp:
        ours = _compile_probe(tmp, "ours", ["-I", INCLUDE])
        ref = _compile_probe(tmp, "ref", ref_flags)
    assert ours[0] == ref[0]  # sizes, offsets, constants, vtable slots
    assert ours[1] == ref[1]  # mangled names of every function that mentions the interface types


@pytest.mark.skipif(not os.path.isdir(REFERENCE_TREE), reason="reference tree not present")
def test_virtual_declaration_order_matches_reference():
    def virtuals(path):
        text = re.sub(r"//.*", "", open(path).read())
        return re.findall(r"virtual\s+[^;{]*?(~?\w+)\s*\(", text)

    ours = virtuals(os.path.join(INCLUDE, "netAbstract.h"))
    ref = virtuals(os.path.join(REFERENCE_TREE, "include", "netAbstract.h"))
    assert ours == ref and len(ours) == 10


def test_class_header_is_cxx14_clean():
    # the only language level the reference states is gnu++14 (.vscode/c_cpp_properties.json:13)
    code = "#include <netCUDA.h>\nint main(){ cuda::net_cuda_options o; return o.max_batch; }\n"
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "t.cpp")
        open(src, "w").write(code)
        subprocess.run(["/usr/bin/g++", "-std=gnu++14", "-Wall", "-Werror", "-fsyntax-only", "-I", INCLUDE, src], check=True)
        # and the C ABI header is plain C
        csrc = os.path.join(tmp, "t.c")
        open(csrc, "w").write("#include <netcuda.h>\nint main(void){ netcuda_desc d; (void)d; return NETCUDA_OK; }\n")
        subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", INCLUDE, csrc], check=True)


def test_c_abi_exports_every_declared_symbol(netcuda):
    names = netcuda.declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(netcuda.lib, n)]
    assert not missing, missing
    assert netcuda.lib.netcuda_abi_version() == 1
    # no torch types in the library's dynamic dependencies
    deps = subprocess.run(["ldd", os.path.join(netcuda.LIB_DIR, "libnetcuda.so")], capture_output=True, text=True).stdout
    assert "torch" not in deps and "c10" not in deps


def test_library_contains_blackwell_sass(netcuda):
    """The shipped SASS must contain tcgen05 MMA (UTC*MMA), TMA loads (UTMALDG) and TMEM loads (LDTM)."""
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", os.path.join(netcuda.LIB_DIR, "libnetcuda.so")], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTCIMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic
    # every kernel family the runtime dispatches to is in the shipped library (a stale build would miss the newer ones)
    for kernel in ("gemm_tn_tcgen05_kernel", "attention_tc16_kernel", "attention_tc_long_kernel", "mlp_i8_stream_kernel", "mlp_i8_umma_cluster_kernel",
                   "gemm_fp32_ordered_kernel", "gemm_fp32_ordered_small_kernel", "layernorm_kernel", "layernorm_halfwarp_kernel", "filter3x3_kernel"):
        assert kernel in sass, kernel


def test_vit_param_count_matches_oracle(netcuda, oracle):
    for name, cfg in netcuda.VIT_PRESETS.items():
        assert netcuda.vit_param_count(cfg) == oracle.vit_param_count(cfg), name
    # torchvision's parameter counts for the two published sizes
    assert netcuda.vit_param_count(netcuda.VIT_PRESETS["vit_tiny_16_224"]) == 5717416
    assert netcuda.vit_param_count(netcuda.VIT_PRESETS["vit_base_16_224"]) == 86567656


def test_fails_loudly_without_gpu(netcuda):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert netcuda.device_count() == 0
    with pytest.raises(netcuda.NetcudaError) as e:
        netcuda.Net.mlp([4, 2], 3)
    assert e.value.code == netcuda.ERR_NO_DEVICE and "no CPU fallback" in str(e.value)
    with pytest.raises(RuntimeError):
        netcuda.HostNet.mlp([4, 2], 3, np.zeros(20, np.float32), np.zeros(6, np.float32))


def test_descriptor_validation(netcuda):
    # argument validation happens before any device work, so it is checkable on CPU
    with pytest.raises(netcuda.NetcudaError) as e:
        netcuda.Net.vit(dict(image_size=224, patch_size=16, dim=100, depth=1, heads=2, mlp_dim=64, n_classes=10))
    assert e.value.code == netcuda.ERR_UNSUPPORTED
    with pytest.raises(netcuda.NetcudaError) as e:
        netcuda.Net.mlp([4, 0], 3)
    assert e.value.code == netcuda.ERR_INVALID
    with pytest.raises(netcuda.NetcudaError) as e:
        netcuda.Net(netcuda.Desc(kind=7))
    assert e.value.code == netcuda.ERR_INVALID


EXAMPLE = os.path.join(ROOT, "examples", "drop_in_app.cpp")


def build_example_app(out_dir):
    """Compile the consumer application of examples/ against the public headers and the two shared libraries (gnu++14, like the
    reference's only stated language level)."""
    lib = os.path.join(ROOT, "vit-fpga_b200", "lib")
    exe = os.path.join(out_dir, "drop_in_app")
    r = subprocess.run(["g++", "-std=gnu++14", "-O2", "-Wall", "-Werror", "-I", INCLUDE, EXAMPLE, "-L", lib, "-lnetcuda_host", "-lnetcuda",
                        "-Wl,-rpath," + lib, "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_example_consumer_application_builds_and_fails_loudly_without_gpu(netcuda):
    """A host application written against net::net_abstract only (examples/drop_in_app.cpp: the reference consumer's shape) builds
    against include/ + the two libraries; without a GPU it stops with the library's error instead of computing anything on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible: the GPU suite runs the application")
    with tempfile.TemporaryDirectory() as d:
        exe = build_example_app(d)
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode != 0
        assert "no CUDA device" in r.stderr
