"""Stress of the tcgen05 INT8 streaming kernel (mlp_umma_stream.cu): the call pattern of tests/test_gpu_nets.py
(test_int8_tcgen05_streaming_kernel_bit_exact: a fresh net, growing batch sizes, three calls each) repeated for `seconds`, every result
checked against the oracle.  A mismatch is printed as a pattern (rows / columns) and saved with the inputs of the failing call and of
the call before it under gpurun_out/ for offline analysis.  Diagnostics, not a test."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import netcuda as nc
from oracle import Oracle

o = Oracle()
import torch  # the test suite's processes hold torch's CUDA context and allocator next to the library's
_keep = torch.empty(1 << 20, device="cuda")
seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
# (the first two run on the single-CTA kernel, the others as split-K clusters: pairs for the 304- and 208-byte fan-ins, four CTAs per tile
#  for the last one -- NETCUDA_MLP_UMMA_PAIR selects, see runtime.cu)
nets = (([272, 48, 10], 1040), ([64, 32], 4080), ([4096, 304, 4096], 4096), ([272, 208, 10], 1040), ([400, 1008, 10], 1040), ([2048, 2048, 640], 1024))
t_end = time.time() + seconds
rounds = calls = nfail = 0
while time.time() < t_end:
    npl, n_ins = nets[rounds % len(nets)]
    rng = np.random.default_rng(78 + rounds)
    n_w = sum(a * b for a, b in zip([n_ins] + npl[:-1], npl))
    wq = np.clip(np.rint(rng.standard_normal(n_w) * 128.0 / np.sqrt(n_ins) * 1.4), -128, 127).astype(np.int8)
    bq = rng.integers(-2000, 2000, sum(npl), dtype=np.int32)
    act = rounds % 3
    net = nc.Net.mlp(npl, n_ins, precision=nc.PREC_INT8, activation=act, max_batch=160)
    net.upload_mlp_i8(wq, bq)
    prev = None
    for batch in (17, 33, 47, 64, 88, 89, 100, 127, 128, 129):
        xq = rng.integers(-128, 128, (batch, n_ins), dtype=np.int8)
        want = o.mlp_forward_i8(xq, wq, bq, npl, n_ins, act)
        for call in range(3):
            got = net.forward_i8(xq)
            calls += 1
            bad = got != want
            if bad.any():
                nfail += 1
                rows, cols = np.unique(np.nonzero(bad)[0]), np.unique(np.nonzero(bad)[1])
                print(f"FAIL round {rounds} net {npl} act {act} batch {batch} call {call}: {int(bad.sum())} of {bad.size} differ; rows "
                      f"{rows.tolist()[:48]} ({rows.size}); cols {cols.tolist()[:24]} ({cols.size}); max |diff| "
                      f"{int(np.abs(got.astype(np.int64) - want)[bad].max())}", flush=True)
                if nfail <= 4:
                    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
                    np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"umma_fail_{nfail}.npz"), npl=np.array(npl), n_ins=n_ins, act=act,
                                        wq=wq, bq=bq, xq=xq, got=got, want=want, prev_xq=prev if prev is not None else np.zeros(0, np.int8))
            prev = xq
    net.close()
    rounds += 1
print(f"{rounds} nets, {calls} checked calls, {nfail} mismatching calls", flush=True)
